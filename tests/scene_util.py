"""Seeded random test scenes (test input generation only; plain numpy)."""
import numpy as np

from oracle import oracle as O


def _mat4(translation, scale, rot_y):
    c, s = np.float32(np.cos(rot_y)), np.float32(np.sin(rot_y))
    m = np.zeros((4, 4), np.float32)     # m[col][row]
    m[0] = [scale[0] * c, 0, -scale[0] * s, 0]
    m[1] = [0, scale[1], 0, 0]
    m[2] = [scale[2] * s, 0, scale[2] * c, 0]
    m[3] = [translation[0], translation[1], translation[2], 1]
    return m.reshape(16)


def random_scene(seed, n_tris=200, n_spheres=20, room=True, all_types=True, sort_morton=False):
    """A Cornell-like room (light, floor, back wall) plus random triangles / spheres under a few random models."""
    rng = np.random.default_rng(seed)
    models, mats, tris, sphs = [], [], [], []

    def add_model(t=(0, 0, 0), s=(1, 1, 1), ry=0.0):
        models.append(_mat4(np.float32(t), np.float32(s), np.float32(ry)))
        return len(models) - 1

    def add_mat(albedo, mtype):
        mats.append((np.float32(albedo), mtype))
        return len(mats) - 1

    def add_tri(v0, v1, v2, mat, model):
        tris.append((np.float32(v0), np.float32(v1), np.float32(v2), mat, model))

    ident = add_model()
    if room:
        light = add_mat((15, 15, 15), 0)
        white = add_mat((0.73, 0.73, 0.73), 1)
        red = add_mat((0.65, 0.05, 0.05), 1)
        add_tri((210, 549, 250), (340, 549, 250), (340, 549, 350), light, ident)
        add_tri((210, 549, 250), (340, 549, 350), (210, 549, 350), light, ident)
        add_tri((0, 0, 0), (550, 0, 0), (550, 0, 550), white, ident)
        add_tri((0, 0, 0), (550, 0, 550), (0, 0, 550), white, ident)
        add_tri((0, 0, 550), (550, 0, 550), (550, 550, 550), red, ident)
        add_tri((0, 0, 550), (550, 550, 550), (0, 550, 550), red, ident)
    n_models = 4
    mids = [add_model(rng.uniform(100, 450, 3), rng.uniform(20, 60, 3), rng.uniform(0, 6.28)) for _ in range(n_models)]
    type_choices = [1, 1, 1, 0, 2, 3] if all_types else [1]
    mat_ids = [add_mat(rng.uniform(0.1, 0.9, 3), int(rng.choice(type_choices))) for _ in range(8)]
    for _ in range(n_tris):
        c = rng.uniform(-1, 1, 3)
        add_tri(c + rng.uniform(-0.4, 0.4, 3), c + rng.uniform(-0.4, 0.4, 3), c + rng.uniform(-0.4, 0.4, 3),
                int(rng.choice(mat_ids)), int(rng.choice(mids)))
    for _ in range(n_spheres):
        sphs.append((np.float32(rng.uniform(-1.5, 1.5, 3)), np.float32(rng.uniform(5, 40)), int(rng.choice(mat_ids)),
                     int(rng.choice(mids))))

    M = np.zeros(len(models), O.MODEL); M["m"] = np.stack(models)
    MT = np.zeros(len(mats), O.MATERIAL)
    for i, (a, t) in enumerate(mats):
        MT["albedo"][i, :3] = a; MT["materialType"][i] = t
    T = np.zeros(len(tris), O.TRIANGLE)
    for i, (a, b, c, m, md) in enumerate(tris):
        T["v0"][i, :3] = a; T["v1"][i, :3] = b; T["v2"][i, :3] = c
        T["materialIndex"][i] = m; T["modelIndex"][i] = md
    S = np.zeros(len(sphs), O.SPHERE)
    for i, (c, r, m, md) in enumerate(sphs):
        S["center"][i, :3] = c; S["radius"][i] = r; S["materialIndex"][i] = m; S["modelIndex"][i] = md
    scene = dict(models=M, triangles=T, spheres=S, materials=MT)
    if sort_morton:
        scene = sort_scene_by_reference_morton(scene)
    return scene


def sort_scene_by_reference_morton(scene):
    """D8: re-emit triangles / spheres in the reference's own Morton order (triangles first, then spheres)."""
    tw, sw = O.model_to_world(scene["models"], scene["triangles"], scene["spheres"])
    enc = O.enclosing_aabb(tw, sw)
    codes = O.morton_codes(tw, sw, enc)["code"]
    T = len(tw)
    ot = np.argsort(codes[:T], kind="stable"); os_ = np.argsort(codes[T:], kind="stable")
    out = dict(scene)
    out["triangles"] = scene["triangles"][ot].copy(); out["spheres"] = scene["spheres"][os_].copy()
    return out


def make_ubo(scene, max_depth=8, random_state=12345, vfov=40.0):
    u = np.zeros(1, O.UBO)
    u["camPos"][0, :3] = (275.0, 275.0, -800.0)
    u["camLookAt"][0, :3] = (275.0, 275.0, 0.0)
    u["camUpDir"][0, :3] = (0.0, 1.0, 0.0)
    u["verticalFOV"] = vfov
    u["numTriangles"] = len(scene["triangles"]); u["numSpheres"] = len(scene["spheres"])
    u["numMaterials"] = len(scene["materials"]); u["numLights"] = 20
    u["maxRayTraceDepth"] = max_depth; u["randomState"] = random_state
    return u


def fuzz_scene(seed):
    """Adversarial scene for the nearest-first / t-culled walk (tests/test_gpu_fuzz.py): a cloud of triangles and spheres of mixed
    sizes placed far from the origin (|coordinate| up to ~1e6, where the hit-point slack's 1e-5 R term matters), with slivers
    (angles down to 1e-6), exactly coincident and nested / overlapping spheres, duplicate primitives (equal-t ties) and one big
    light.  Returns (scene, camera positions): one camera outside the scene box and one inside."""
    rng = np.random.default_rng(1000 + seed)
    off_mag = [0.0, 1.0e4, 1.0e5, 1.0e6][seed % 4]
    ext = max(300.0, off_mag * 0.02) * rng.uniform(0.5, 2.0)                 # scene extent
    offset = rng.uniform(-1, 1, 3); offset = offset / np.linalg.norm(offset) * off_mag
    nt, ns = int(rng.integers(40, 400)), int(rng.integers(0, 60))
    mats = [((15, 15, 15), 0), ((0.7, 0.7, 0.7), 1), ((0.6, 0.2, 0.2), 1), ((0.2, 0.6, 0.3), 1), ((0.8, 0.8, 0.2), 2 if seed % 3 == 0 else 1)]
    T = np.zeros(nt + 2, O.TRIANGLE); S = np.zeros(ns, O.SPHERE)
    c = offset + rng.uniform(-0.5, 0.5, (nt, 3)) * ext
    size = ext * 10.0 ** rng.uniform(-3.0, -0.5, (nt, 1))
    v0 = c + rng.normal(size=(nt, 3)) * size; v1 = c + rng.normal(size=(nt, 3)) * size; v2 = c + rng.normal(size=(nt, 3)) * size
    sl = rng.random(nt) < 0.25                                               # slivers: v2 almost on the line v0-v1
    # two seeds of three keep every needle above the 1e-4 rad below which eta_leaf_kernel gives up (eta = inf disables t-culling for
    # the whole scene); every third seed goes down to 1e-6 rad and exercises exactly that fallback
    ang = 10.0 ** (rng.uniform(-6, -2, (nt, 1)) if seed % 3 == 2 else rng.uniform(-2.7, -1, (nt, 1)))
    v2 = np.where(sl[:, None], v0 + (v1 - v0) * rng.uniform(0.2, 0.8, (nt, 1)) + rng.normal(size=(nt, 3)) * size * ang, v2)
    dup = rng.random(nt) < 0.1                                               # duplicates of the previous triangle: equal-t ties
    for k in np.nonzero(dup)[0]:
        if k > 0:
            v0[k], v1[k], v2[k] = v0[k - 1], v1[k - 1], v2[k - 1]
    T["v0"][:nt, :3] = v0; T["v1"][:nt, :3] = v1; T["v2"][:nt, :3] = v2
    T["materialIndex"][:nt] = rng.integers(1, len(mats), nt)
    lo, hi = offset - 0.55 * ext, offset + 0.55 * ext                        # a light quad above the cloud
    T["v0"][nt, :3] = (lo[0], hi[1], lo[2]); T["v1"][nt, :3] = (hi[0], hi[1], lo[2]); T["v2"][nt, :3] = (hi[0], hi[1], hi[2])
    T["v0"][nt + 1, :3] = (lo[0], hi[1], lo[2]); T["v1"][nt + 1, :3] = (hi[0], hi[1], hi[2]); T["v2"][nt + 1, :3] = (lo[0], hi[1], hi[2])
    if ns:
        sc_ = offset + rng.uniform(-0.45, 0.45, (ns, 3)) * ext
        sr = ext * 10.0 ** rng.uniform(-2.5, -0.7, ns)
        for k in range(1, ns):
            m = rng.random()
            if m < 0.2:   sc_[k] = sc_[k - 1]; sr[k] = sr[k - 1] * rng.uniform(0.3, 0.95)      # nested, concentric
            elif m < 0.3: sc_[k] = sc_[k - 1]; sr[k] = sr[k - 1]                                # coincident
            elif m < 0.5: sc_[k] = sc_[k - 1] + rng.normal(size=3) * sr[k - 1] * 0.5            # overlapping
        S["center"][:, :3] = sc_; S["radius"] = sr; S["materialIndex"] = rng.integers(1, len(mats), ns)
    M = np.zeros(1, O.MODEL); M["m"][0] = np.eye(4, dtype=np.float32).reshape(16)
    MT = np.zeros(len(mats), O.MATERIAL)
    for i, (a, t) in enumerate(mats):
        MT["albedo"][i, :3] = a; MT["materialType"][i] = t
    scene = dict(models=M, triangles=T, spheres=S, materials=MT)
    d = rng.normal(size=3); d /= np.linalg.norm(d)
    cams = [(offset + d * ext * 1.6, offset), (offset + rng.uniform(-0.2, 0.2, 3) * ext, offset + d * ext)]
    return scene, cams


def ubo_with_camera(scene, cam, look, max_depth=16, random_state=7, vfov=50.0):
    u = make_ubo(scene, max_depth=max_depth, random_state=random_state, vfov=vfov)
    u["camPos"][0, :3] = np.float32(cam); u["camLookAt"][0, :3] = np.float32(look)
    return u
