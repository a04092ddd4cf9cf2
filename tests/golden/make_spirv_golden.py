"""Generates tests/golden/spirv_*.npz by executing THE REFERENCE'S OWN COMPILED SHADERS.

The reference commits its compute shaders as SPIR-V binaries (`RaytracerGPU_MastersProject/shaders/compiled/*.spv`, built by its
`compile.bat` with glslc).  No Vulkan driver exists in this image, but `oracle/spirv_interp.py` executes those binaries on the CPU,
instruction by instruction, in the dispatch order and with the dispatch sizes of `RaytracerBVH.cpp:734-1050` /
`Raytracer.cpp:394-538` / `LogisticMap.cpp:384`.  The vectors written here are therefore outputs of the reference itself (its
binaries, not a restatement of its sources); the driver-defined built-in arithmetic follows the pins listed in
`oracle/spirv_interp.py` (sin / cos = the oracle's pinned recipe, the single thing borrowed from the oracle).

This script needs `/root/reference` and therefore only runs in the build container:
    python tests/golden/make_spirv_golden.py            # ~2 minutes
The fixtures it writes are committed; `tests/test_spirv_golden.py` checks the C oracle (CPU) and the CUDA path (GPU) against them.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import scene_util as SU  # noqa: E402
from oracle import oracle as O  # noqa: E402

SPV_DIR = os.environ.get("RTB_REFERENCE_SPV", "/root/reference/RaytracerGPU_MastersProject/shaders/compiled")

# name -> scene (tests/scene_util.random_scene arguments, or "dups"), image size, dispatches (= raysPerPixel), programs to run
CASES = {
    # K1..K6 + raytraceBVH.comp (BVH program) + raytrace.comp (non-BVH program) + fragment resolve
    "room": dict(scene=dict(seed=1, n_tris=30, n_spheres=6), W=48, H=40, spp=3, depth=8, random_state=12345,
                 programs=("bvh", "linear", "resolve")),
    # primitives pre-sorted by the reference's Morton code (the layout the big configs use), other seed / depth
    "sorted": dict(scene=dict(seed=2, n_tris=90, n_spheres=14, sort_morton=True), W=40, H=40, spp=2, depth=5, random_state=777,
                   programs=("bvh",)),
    # no room: most rays miss, spheres dominate; odd image size (partial workgroups)
    "spheres": dict(scene=dict(seed=3, n_tris=0, n_spheres=40, room=False), W=37, H=29, spp=2, depth=8, random_state=4242,
                    programs=("bvh", "linear")),
    # duplicated primitives: equal Morton codes -> the index tie-break of ConstructHLBVH.comp:64-67; build only
    "dups": dict(scene="dups", W=8, H=8, spp=1, depth=2, random_state=9, programs=("bvh",)),
    # a deeper tree (440 primitives): stack pushes / pops, long leaf sequences
    "mesh": dict(scene=dict(seed=7, n_tris=400, n_spheres=40), W=32, H=32, spp=1, depth=8, random_state=31337, programs=("bvh",)),
    # degenerate geometry: zero-area triangles (NaN normals -> NaN hit distances), a zero-radius sphere, coordinates that
    # overflow to inf in the cross products; axis-parallel rays through the odd-sized image centre
    "degenerate": dict(scene="degenerate", W=33, H=33, spp=2, depth=6, random_state=5, programs=("bvh", "linear")),
    # large zero-area triangles in front of the room: their boxes are hit, the normal is NaN, the NaN hit distance is ACCEPTED by
    # the shader's comparisons (raytraceBVH.comp:130,137) and poisons closestSoFar for the rest of that ray
    "slivers": dict(scene="slivers", W=32, H=32, spp=2, depth=6, random_state=6, programs=("bvh", "linear")),
    # smallest trees: N = 2 (one internal node) and N = 3
    "two": dict(scene=dict(seed=5, n_tris=1, n_spheres=1, room=False), W=8, H=8, spp=1, depth=2, random_state=1, programs=("bvh",)),
    "three": dict(scene=dict(seed=6, n_tris=2, n_spheres=1, room=False), W=8, H=8, spp=1, depth=2, random_state=2, programs=("bvh",)),
}
# edge cases (written as spirv_edge_*.npz): they pin the ORACLE on the corners of the parameter space
EDGE_CASES = {
    # N = 1: the root is the only node and a leaf (ConstructHLBVH.comp with zero internal nodes)
    "edge_one": dict(scene=dict(seed=31, n_tris=0, n_spheres=1, room=False), W=16, H=16, spp=2, depth=3, random_state=3, programs=("bvh", "linear")),
    # randomState + 1 wraps to 0: every pixel's base seed is 0 (random.glsl:10), only the alpha chain separates the samples
    "edge_seedwrap": dict(scene=dict(seed=32, n_tris=20, n_spheres=4), W=16, H=12, spp=3, depth=4, random_state=0xFFFFFFFF, programs=("bvh",)),
    # maxRayTraceDepth 0 (no ray at all: colour stays 0, the alpha chain still advances) and 1 (primary hit only)
    "edge_depth0": dict(scene=dict(seed=33, n_tris=20, n_spheres=4), W=12, H=12, spp=2, depth=0, random_state=5, programs=("bvh", "linear")),
    "edge_depth1": dict(scene=dict(seed=33, n_tris=20, n_spheres=4), W=12, H=12, spp=2, depth=1, random_state=5, programs=("bvh", "linear")),
    # wide field of view, non-square image whose sides are multiples of the 32x32 workgroup (one extra, empty workgroup row / column)
    "edge_fov90": dict(scene=dict(seed=34, n_tris=24, n_spheres=6), W=64, H=32, spp=1, depth=4, random_state=8, vfov=90.0, programs=("bvh",)),
    # six dispatches: a long alpha seed chain (raytraceBVH.comp:350,372)
    "edge_chain6": dict(scene=dict(seed=35, n_tris=16, n_spheres=3), W=10, H=10, spp=6, depth=3, random_state=13, programs=("bvh",)),
}
CASES_ALL = dict(CASES, **EDGE_CASES)
LOGISTIC = dict(points=2048, W=96, H=64, steps=3, seed=11)


def make_scene(spec):
    if spec == "dups":
        base = SU.random_scene(4, n_tris=12, n_spheres=4)
        sc = dict(base)
        sc["triangles"] = np.concatenate([base["triangles"], base["triangles"][6:], base["triangles"][6:12]])
        sc["spheres"] = np.concatenate([base["spheres"], base["spheres"], base["spheres"][:2]])
        return sc
    if spec == "degenerate":
        sc = SU.random_scene(8, n_tris=24, n_spheres=5)
        t = sc["triangles"]
        t["v1"][8] = t["v0"][8]                                   # zero-area: two equal vertices
        t["v1"][9] = t["v0"][9]; t["v2"][9] = t["v0"][9]          # a point
        t["v2"][10] = t["v0"][10] + 2.0 * (t["v1"][10] - t["v0"][10])   # collinear vertices
        t["v0"][11][:3] = (1e20, -1e20, 1e20)                     # products overflow to inf
        t["v0"][12][:3] = (0.0, 0.0, 0.0); t["v1"][12][:3] = (1e-30, 0.0, 0.0); t["v2"][12][:3] = (0.0, 1e-30, 0.0)   # denormal area
        sc["spheres"]["radius"][1] = 0.0
        sc["spheres"]["radius"][2] = -3.0
        return sc
    if spec == "slivers":
        sc = SU.random_scene(9, n_tris=16, n_spheres=3)
        t = sc["triangles"]
        for i, (a, b) in enumerate([((100, 100, 300), (400, 400, 320)), ((150, 400, 200), (420, 120, 260))]):
            k = 6 + i                                              # first random triangles (model index kept -> identity below)
            t["modelIndex"][k] = 0
            t["v0"][k][:3] = a; t["v1"][k][:3] = b
            t["v2"][k][:3] = 0.5 * (np.float32(a) + np.float32(b))    # collinear: cross(u, v) = 0 -> normalize -> NaN
        t["modelIndex"][8] = 0
        t["v0"][8][:3] = t["v1"][8][:3] = t["v2"][8][:3] = (275, 275, 100)   # a point
        return sc
    return SU.random_scene(**spec)


def _sincos(x):
    s, c = O.pin_sincos(x)
    return float(s), float(c)


def module(name):
    from oracle.spirv_interp import Module
    return Module(os.path.join(SPV_DIR, name + ".spv"), sincos=_sincos)


def run_build(sc, ubo):
    """S1 as recorded at RaytracerBVH.cpp:778-991: the six build dispatches, every intermediate buffer kept."""
    T, S = len(sc["triangles"]), len(sc["spheres"]); N = T + S
    ub = bytearray(ubo.tobytes())
    tris = bytearray(sc["triangles"].tobytes()) if T else bytearray(64)
    sphs = bytearray(sc["spheres"].tobytes()) if S else bytearray(32)
    models = bytearray(sc["models"].tobytes())
    scratch = bytearray(80)                                                         # 20 floats, RaytracerBVH.cpp:480-508
    out = {}
    module("ModelSpaceToWorldSpace.comp").dispatch((N // 32 + 1, 1, 1), {0: ub, 1: models, 2: tris, 3: sphs})            # :789
    out["tris_w"] = np.frombuffer(bytes(tris), O.TRIANGLE)[:T].copy(); out["sphs_w"] = np.frombuffer(bytes(sphs), O.SPHERE)[:S].copy()
    enc = bytearray(32)
    module("GetEnclosingAABB.comp").dispatch((1, 1, 1), {0: ub, 1: enc, 2: tris, 3: sphs, 4: scratch}, lockstep=True)     # :842
    out["enclosing"] = np.frombuffer(bytes(enc), O.ENCLOSING).copy()
    m1, m2 = bytearray(12 * N), bytearray(12 * N)
    module("GenerateMortonCodesOfPrimitives.comp").dispatch((N // 32 + 1, 1, 1), {0: ub, 1: enc, 2: tris, 3: sphs, 4: m1, 5: scratch})   # :879
    out["morton_unsorted"] = np.frombuffer(bytes(m1), O.MORTON).copy()
    module("RadixSortSimple.comp").dispatch((1, 1, 1), {0: ub, 1: m1, 2: m2}, lockstep=True)                               # :916
    out["morton"] = np.frombuffer(bytes(m1), O.MORTON).copy()
    nodes, cinfo = bytearray(40 * (2 * N - 1)), bytearray(8 * (2 * N - 1))
    module("ConstructHLBVH.comp").dispatch((N // 256 + 1, 1, 1), {0: ub, 1: tris, 2: sphs, 3: m1, 4: nodes, 5: cinfo})      # :954
    out["nodes_unfitted"] = np.frombuffer(bytes(nodes), O.NODE).copy(); out["cinfo_unfitted"] = np.frombuffer(bytes(cinfo), O.CINFO).copy()
    module("ConstructAABBsOfInternalNodes.comp").dispatch((N // 32 + 1, 1, 1), {0: ub, 1: nodes, 2: cinfo})                 # :991
    out["nodes"] = np.frombuffer(bytes(nodes), O.NODE).copy(); out["cinfo"] = np.frombuffer(bytes(cinfo), O.CINFO).copy()
    return out


def run_trace(shader, ubo, W, H, spp, tris_w, sphs_w, mats, nodes):
    """S2 as recorded at RaytracerBVH.cpp:998-1050 (Raytracer.cpp:485-509 for the non-BVH program): the image is cleared to
    (0, 0, 0, 1) (:772) and the shader is dispatched raysPerPixel times over (W/32+1, H/32+1) groups of 32x32."""
    from oracle.spirv_interp import Image
    img = np.zeros((H, W, 4), np.float32); img[..., 3] = 1.0
    im = Image(img.tolist())
    res = {0: bytearray(ubo.tobytes()), 1: im, 2: bytearray(tris_w.tobytes()) or bytearray(64), 3: bytearray(sphs_w.tobytes()) or bytearray(32),
           4: bytearray(mats.tobytes())}
    if nodes is not None:
        res[5] = bytearray(nodes.tobytes()); res[6] = bytearray(80)
    else:
        res[5] = bytearray(80)
    m = module(shader)
    per_dispatch = []
    for _ in range(spp):
        # threads outside the image return before touching memory (raytraceBVH.comp:346-347): skipped, not interpreted
        m.dispatch((W // 32 + 1, H // 32 + 1, 1), res, only=lambda g: g[0] < W and g[1] < H)
        per_dispatch.append(np.array(im.px, np.float32))
    # the reference binary's own work counts (OpFunctionCall executions): rays = hitBVH / sceneHit calls, node visits = AABBhitCheck
    # calls, triangle / sphere tests, material reads = emitted() calls -- the quantities bench.py's roofline bytes are built from
    counts = np.array([m.call_count("sceneHit"), m.call_count("AABBhitCheck"), m.call_count("triangleHit"), m.call_count("sphereHit"),
                       m.call_count("emitted")], np.uint64)
    return np.stack(per_dispatch), m.n_executed, counts


def run_resolve(image, rays_per_pixel):
    """SingleTriangleFullScreen.frag per pixel (uv = pixel centre, nearest texel), then the fixed-function UNORM8 conversion."""
    from oracle.spirv_interp import Image, f32
    H, W, _ = image.shape
    m = module("SingleTriangleFullScreen.frag")
    im = Image(image.tolist())
    ubo = bytearray(np.array([rays_per_pixel, 0, 0, 0], np.uint32).tobytes())
    out = np.zeros((H, W, 4), np.float32)
    for y in range(H):
        for x in range(W):
            uv = [f32((x + 0.5) / W), f32((y + 0.5) / H)]
            out[y, x] = m.run_stage({0: uv}, {0: ubo, 1: im})[0]
    return out


def run_logistic(cfg):
    """logistic.comp, LogisticMap.cpp:384: `steps` dispatches of points/1024 groups; the rgba8 image is never cleared in between."""
    from oracle.spirv_interp import Image
    rng = np.random.default_rng(cfg["seed"])
    n, W, H = cfg["points"], cfg["W"], cfg["H"]
    pts = np.zeros((n, 2), np.float32)
    pts[:, 0] = rng.uniform(0.0, 1.0, n); pts[:, 1] = rng.uniform(2.5, 4.0, n)
    pts[0] = (0.5, 4.0); pts[1] = (0.0, 3.0); pts[2] = (1.0, 3.9)              # edge values: x' = 1 -> y = 0; x' = 0 -> y = H (outside)
    color = np.array([0.25, 0.5, 1.0, 1.0], np.float32)
    ubo = bytearray(np.array([*color, 0.0, W, H, 0.0], np.float32).tobytes())
    buf = bytearray(pts.tobytes())
    im = Image(np.zeros((H, W, 4), np.float32).tolist())
    m = module("logistic.comp")
    states = []
    for _ in range(cfg["steps"]):
        m.dispatch((n // 1024, 1, 1), {0: ubo, 1: buf, 2: im})
        states.append(np.frombuffer(bytes(buf), np.float32).reshape(n, 2).copy())
    return dict(points0=pts, color=color, points=np.stack(states), plotted=(np.array(im.px, np.float32).sum(-1) > 0))


def generate_pixels(spec, W, H, n_pixels, spp, random_state, center_box=None, seed=2024, fixed=()):
    """A host-built scene (raytracergpu_mastersproject_b200/host/Scenes.cpp) through the reference binaries: S1 on the whole scene,
    S2 (raytraceBVH.comp.spv) on a seeded sample of pixels of the full-size image -- every pixel is independent."""
    from oracle.spirv_interp import Image
    from raytracergpu_mastersproject_b200 import scenes
    sc = scenes.load_scene(spec)
    ubo = SU.make_ubo(sc, max_depth=sc["max_depth"], random_state=random_state, vfov=sc["vfov"])
    t0 = time.time()
    b = run_build(sc, ubo)
    rng = np.random.default_rng(seed)
    px = set(fixed)
    while len(px) < n_pixels:
        if center_box is not None and rng.random() < 0.7:
            x0, y0, x1, y1 = center_box
            px.add((int(rng.integers(x0, x1)), int(rng.integers(y0, y1))))
        else:
            px.add((int(rng.integers(0, W)), int(rng.integers(0, H))))
    px = sorted(px)

    class SparseRow(dict):          # the image as {y: {x: texel}}: only the sampled pixels are ever touched
        def __missing__(self, k):
            v = [0.0, 0.0, 0.0, 1.0]
            self[k] = v
            return v

    class SparseImage(Image):
        def __init__(self, w, h):
            self.w, self.h = w, h

            class Rows(dict):
                def __missing__(self, k):
                    r = SparseRow()
                    self[k] = r
                    return r
            self.px = Rows()
    im = SparseImage(W, H)
    scratch = bytearray(80)
    res = {0: bytearray(ubo.tobytes()), 1: im, 2: bytearray(b["tris_w"].tobytes()), 3: bytearray(b["sphs_w"].tobytes()),
           4: bytearray(sc["materials"].tobytes()), 5: bytearray(b["nodes"].tobytes()), 6: scratch}
    m = module("raytraceBVH.comp")
    want = set(px)
    vals = []
    for _ in range(spp):
        m.dispatch((W // 32 + 1, H // 32 + 1, 1), res, only=lambda g: (g[0], g[1]) in want)
        vals.append(np.array([im.px[y][x] for (x, y) in px], np.float32))
    print(f"{spec}: {len(sc['triangles'])} triangles, {len(sc['spheres'])} spheres, {len(px)} pixels of {W}x{H} x{spp}: {m.n_executed} SPIR-V "
          f"instructions traced, {time.time() - t0:.1f} s")
    out = dict(spec=np.array(spec), models=sc["models"], triangles=sc["triangles"], spheres=sc["spheres"], materials=sc["materials"], ubo=ubo,
               W=np.int32(W), H=np.int32(H), spp=np.int32(spp), pixels=np.array(px, np.int32), values=np.stack(vals),
               scratch=np.frombuffer(bytes(scratch), np.float32).copy())
    out.update(b)
    return out


def generate_c1(n_pixels=1500, spp=2, random_state=12345):
    """BASELINE config C1 on its real scene: the host's complexScene (Scenes.cpp:337-398 with the stand-in models) at the
    reference's 800x800 window, S1 by the reference binaries on all 1 000+ primitives, S2 by raytraceBVH.comp.spv on a seeded
    sample of pixels (every pixel is independent) -- including the shader's debug pixel (175, 650), whose scratch[] writes
    (raytraceBVH.comp:40-45,256-263,288-313) are kept as well."""
    from oracle.spirv_interp import Image
    from raytracergpu_mastersproject_b200 import scenes
    sc = scenes.load_scene("complexScene")
    W = H = 800
    ubo = SU.make_ubo(sc, max_depth=sc["max_depth"], random_state=random_state, vfov=sc["vfov"])
    t0 = time.time()
    b = run_build(sc, ubo)
    rng = np.random.default_rng(2024)
    px = {(175, 650), (0, 0), (799, 799), (400, 400)}
    while len(px) < n_pixels:
        px.add((int(rng.integers(0, W)), int(rng.integers(0, H))))
    px = sorted(px)
    img = np.zeros((H, W, 4), np.float32); img[..., 3] = 1.0
    im = Image(img.tolist())
    scratch = bytearray(80)
    res = {0: bytearray(ubo.tobytes()), 1: im, 2: bytearray(b["tris_w"].tobytes()), 3: bytearray(b["sphs_w"].tobytes()),
           4: bytearray(sc["materials"].tobytes()), 5: bytearray(b["nodes"].tobytes()), 6: scratch}
    m = module("raytraceBVH.comp")
    want = set(px)
    vals = []
    for _ in range(spp):
        m.dispatch((W // 32 + 1, H // 32 + 1, 1), res, only=lambda g: (g[0], g[1]) in want)
        vals.append(np.array([im.px[y][x] for (x, y) in px], np.float32))
    print(f"c1pixels: {len(sc['triangles'])} triangles, {len(sc['spheres'])} spheres, {len(px)} pixels x{spp}: {m.n_executed} SPIR-V "
          f"instructions traced, {time.time() - t0:.1f} s")
    out = dict(models=sc["models"], triangles=sc["triangles"], spheres=sc["spheres"], materials=sc["materials"], ubo=ubo,
               W=np.int32(W), H=np.int32(H), spp=np.int32(spp), pixels=np.array(px, np.int32), values=np.stack(vals),
               scratch=np.frombuffer(bytes(scratch), np.float32).copy())
    out.update(b)
    return out


def generate(name):
    c = CASES_ALL[name]
    sc = make_scene(c["scene"])
    ubo = SU.make_ubo(sc, max_depth=c["depth"], random_state=c["random_state"], vfov=c.get("vfov", 40.0))
    t0 = time.time()
    out = dict(models=sc["models"], triangles=sc["triangles"], spheres=sc["spheres"], materials=sc["materials"], ubo=ubo,
               W=np.int32(c["W"]), H=np.int32(c["H"]), spp=np.int32(c["spp"]))
    b = run_build(sc, ubo)
    out.update(b)
    n_ins = 0
    if "bvh" in c["programs"]:
        out["images_bvh"], k, out["counts_bvh"] = run_trace("raytraceBVH.comp", ubo, c["W"], c["H"], c["spp"], b["tris_w"], b["sphs_w"], sc["materials"], b["nodes"])
        n_ins += k
    if "linear" in c["programs"]:
        out["images_linear"], k, out["counts_linear"] = run_trace("raytrace.comp", ubo, c["W"], c["H"], c["spp"], b["tris_w"], b["sphs_w"], sc["materials"], None)
        n_ins += k
    if "resolve" in c["programs"]:
        out["resolved"] = run_resolve(out["images_bvh"][-1], c["spp"])
    print(f"{name}: {len(sc['triangles'])} triangles, {len(sc['spheres'])} spheres, {c['W']}x{c['H']} x{c['spp']}: "
          f"{n_ins} SPIR-V instructions traced, {time.time() - t0:.1f} s")
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or (list(CASES_ALL) + ["logistic", "c1pixels", "c2pixels"])
    for name in names:
        if name == "c1pixels":
            np.savez_compressed(os.path.join(HERE, "spirv_c1pixels.npz"), **generate_c1())
        elif name == "c2pixels":
            # BASELINE config C2's generator at a size the interpreter can hold (meshRoom:70 = 9 806 primitives instead of 871 206): the
            # displaced sphere in the simpleScene room at C2's 1920x1080, depth 8 -- incl. rays tangent to the tessellated surface
            import hashlib
            out = generate_pixels("meshRoom:70:5", 1920, 1080, 700, 2, 1, center_box=(560, 140, 1360, 940))
            big = ("models", "triangles", "spheres", "materials", "tris_w", "sphs_w", "morton_unsorted", "morton", "nodes_unfitted",
                   "cinfo_unfitted", "nodes", "cinfo")       # kept as SHA-256 digests: the scene is rebuilt by the host at test time
            slim = {k: v for k, v in out.items() if k not in big}
            slim["digest_names"] = np.array(big)
            slim["digests"] = np.array([hashlib.sha256(np.ascontiguousarray(out[k]).tobytes()).hexdigest() for k in big])
            np.savez_compressed(os.path.join(HERE, "spirv_c2pixels.npz"), **slim)
        elif name == "logistic":
            np.savez_compressed(os.path.join(HERE, "spirv_logistic.npz"), **run_logistic(LOGISTIC))
            print("logistic done")
        else:
            np.savez_compressed(os.path.join(HERE, f"spirv_{name}.npz"), **generate(name))
