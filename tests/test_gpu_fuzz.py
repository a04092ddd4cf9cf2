"""GPU: randomized hardening of the nearest-first, t-culled walk (trace_wave.cu wave_step_u / trace_tail_kernel).

The walk drops record entries whose slab interval lies beyond the closest hit so far; that is sound only because the record boxes
were grown by the hit-point slack eta (bvh_build.cu eta_leaf_kernel), whose constants were derived by hand.  Two checks over
240 seeded adversarial scenes (tests/scene_util.py fuzz_scene: |coordinate| up to 1e6, slivers down to 1e-6 rad, nested /
coincident / overlapping spheres, duplicated triangles, depth 16, cameras outside and inside the scene box):
  1. the frame of the culled walk equals the frame of the exact-record walk in the reference's order, bit for bit;
  2. every primary hit the exact walk accepts lies within the eta the library computed for that primitive of the primitive's
     reference leaf box (the property the culling argument rests on), checked on the host from the exported eta values.
"""
import ctypes as C

import numpy as np
import pytest

import scene_util as SU
from oracle import oracle as O

pytestmark = pytest.mark.gpu
W, H, SPP = 40, 28, 3


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _primary_dirs(ubo, W, H):
    """primary rays as make_camera / primary_direction evaluate them (binary32, the shader's operation order)"""
    f = np.float32
    cam = ubo["camPos"][0, :3].astype(f); look = ubo["camLookAt"][0, :3].astype(f); up = ubo["camUpDir"][0, :3].astype(f)
    nrm = lambda v: (v / np.sqrt(f(f(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]), dtype=f)).astype(f)  # noqa: E731
    cross = lambda a, b: np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], f)  # noqa: E731
    theta = f(ubo["verticalFOV"][0]) * f(0.017453292519943295)
    h = f(np.tan(theta / f(2), dtype=f))
    vh = f(f(2) * h * f(10)); vw = f(vh * f(f(W) / f(H)))
    w_ = nrm(cam - look); u_ = nrm(cross(up, w_)); v_ = cross(w_, u_)
    vu = (vw * u_).astype(f); vv = (vh * (-v_)).astype(f)
    du = (vu / f(W)).astype(f); dv = (vv / f(H)).astype(f)
    ul = (((cam - f(10) * w_) - vu / f(2)) - vv / f(2)).astype(f)
    p00 = (ul + f(0.5) * (du + dv)).astype(f)
    xs = np.arange(W, dtype=f)[None, :, None]; ys = np.arange(H, dtype=f)[:, None, None]
    ps = ((p00 + xs * du) + ys * dv).astype(f) - cam
    n1 = (ps / np.sqrt(((ps[..., 0] * ps[..., 0] + ps[..., 1] * ps[..., 1]) + ps[..., 2] * ps[..., 2]).astype(f))[..., None]).astype(f)
    return cam, n1


@pytest.mark.parametrize("block", range(12))
def test_culled_walk_equals_exact_walk_and_hits_stay_within_eta(device, block):
    from raytracergpu_mastersproject_b200 import Buffer, Raytracer, capi
    L = capi.lib()
    checked_hits = culled_scenes = 0
    for seed in range(block * 20, block * 20 + 20):
        sc, cams = SU.fuzz_scene(seed)
        T, S = len(sc["triangles"]), len(sc["spheres"]); N = T + S
        for ci, (cam, look) in enumerate(cams):
            ubo = SU.ubo_with_camera(sc, cam, look, max_depth=16, random_state=seed + 1)
            rt = Raytracer(device, W, H)
            rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
            rt.build_bvh(ubo)
            hp = Buffer(device, 4, W * H); ht = Buffer(device, 4, W * H)
            rt.clear_image(); rt.raytrace(ubo, SPP, flags=capi.TRACE_EXACT_NODES | capi.TRACE_NO_PRIMARY_SHARING, hit_prim=hp, hit_t=ht); device.wait_idle()
            exact = rt.read_image(); prim = hp.read(np.uint32).reshape(H, W); t = ht.read(np.float32).reshape(H, W)
            for fl in (capi.TRACE_WIDE_NODES, capi.TRACE_WIDE_NODES | capi.TRACE_NO_PRIMARY_SHARING):
                rt.clear_image(); rt.raytrace(ubo, SPP, flags=fl); device.wait_idle()
                got = rt.read_image()
                bad = int((_bits(got) != _bits(exact)).any(axis=-1).sum())
                assert bad == 0, f"seed {seed} camera {ci} flags {fl}: {bad} pixels differ from the exact-record walk"
            # (2) accepted primary hits within eta of their leaf box
            eta = np.zeros(N, np.float32); region = np.zeros(6, np.float32)
            capi.check(L.rtb_export_hit_slack(device.handle, eta.ctypes.data_as(C.c_void_p), N, region.ctypes.data_as(C.c_void_p)))
            nodes = rt.nodes.read(O.NODE, 2 * N - 1)
            o, dirs = _primary_dirs(ubo, W, H)
            inside = bool(np.all(o >= region[:3]) and np.all(o <= region[3:]))
            hit = prim != 0xFFFFFFFF
            culled_scenes += int(np.isfinite(eta).all())
            hit &= np.isfinite(eta)[np.minimum(prim, N - 1)]
            if not inside or not hit.any():
                continue
            g = prim[hit].astype(np.int64)
            P = o[None, :].astype(np.float64) + t[hit][:, None].astype(np.float64) * dirs[hit].astype(np.float64)
            box = nodes["aabb"][N - 1 + g].astype(np.float64)                       # minX maxX minY maxY minZ maxZ
            lo, hi = box[:, 0::2], box[:, 1::2]
            dist = np.max(np.maximum(np.maximum(lo - P, P - hi), 0.0), axis=1)
            worst = dist - eta[g].astype(np.float64)
            assert np.all(worst <= 0), (f"seed {seed} camera {ci}: an accepted hit lies {dist[np.argmax(worst)]:.3e} outside its leaf box, "
                                        f"eta = {eta[g][np.argmax(worst)]:.3e}")
            checked_hits += int(hit.sum())
    assert checked_hits > 1000 and culled_scenes >= 20, (checked_hits, culled_scenes)    # most scenes really run the t-culled walk


@pytest.mark.parametrize("n_big", [0, 1, 3, 4, 7, 45, 46, 120])
def test_hoisted_big_primitives(device, n_big):
    """Scenes of >= 512 primitives take the fused build, whose derived records hoist the BIG leaves (boxes above 1/256 of the scene's
    surface area) out of the hierarchy into a chain of records in front of the root (bvh_build.cu pack_wide_kernel): none, fewer than one
    record's worth, exactly full records, the maximum (45), one too many (nothing is hoisted) and far too many; big triangles AND big
    spheres, overlapping each other and the small geometry, duplicated (equal-t ties between a hoisted and an in-tree primitive).  The
    production frame must equal the exact-record walk's bit for bit, shared and unshared primaries, and the reference-order walk over
    the un-hoisted records must too."""
    from raytracergpu_mastersproject_b200 import Raytracer, capi
    rng = np.random.default_rng(77 + n_big)
    nt = 9000
    T = np.zeros(nt + n_big + (2 if n_big else 0), O.TRIANGLE)
    c = rng.uniform(100, 450, (nt, 3)); sz = rng.uniform(1, 6, (nt, 1))
    T["v0"][:nt, :3] = c + rng.normal(size=(nt, 3)) * sz; T["v1"][:nt, :3] = c + rng.normal(size=(nt, 3)) * sz; T["v2"][:nt, :3] = c + rng.normal(size=(nt, 3)) * sz
    T["materialIndex"][:nt] = rng.integers(1, 4, nt)
    for k in range(n_big):                                   # big triangles: walls, floors, slanted sheets through the cloud
        a = rng.uniform(0, 550, 3); u = rng.normal(size=3); v = rng.normal(size=3)
        u *= rng.uniform(300, 700) / np.linalg.norm(u); v *= rng.uniform(300, 700) / np.linalg.norm(v)
        T["v0"][nt + k, :3] = a; T["v1"][nt + k, :3] = a + u; T["v2"][nt + k, :3] = a + v
        T["materialIndex"][nt + k] = 0 if k == 0 else int(rng.integers(1, 4))
    if n_big:                                                # an exact duplicate of a small triangle made big's neighbour, and of a big one
        T[nt + n_big] = T[nt]; T[nt + n_big + 1] = T[5]
    ns = 40 + (3 if n_big else 0)
    S = np.zeros(ns, O.SPHERE)
    S["center"][:, :3] = rng.uniform(100, 450, (ns, 3)); S["radius"] = rng.uniform(2, 12, ns); S["materialIndex"] = rng.integers(1, 4, ns)
    if n_big:
        S["radius"][-3:] = (160.0, 220.0, 90.0)              # big spheres: hoisted too
    M = np.zeros(1, O.MODEL); M["m"][0] = np.eye(4, dtype=np.float32).reshape(16)
    MT = np.zeros(4, O.MATERIAL)
    for i, (a, t) in enumerate([((12, 12, 12), 0), ((0.7, 0.7, 0.7), 1), ((0.6, 0.3, 0.2), 1), ((0.2, 0.5, 0.7), 1)]):
        MT["albedo"][i, :3] = a; MT["materialType"][i] = t
    sc = dict(models=M, triangles=T, spheres=S, materials=MT)
    W, H, spp = 64, 40, 3
    for cam, look in [((275.0, 275.0, -800.0), (275.0, 275.0, 0.0)), ((300.0, 260.0, 240.0), (100.0, 300.0, 500.0))]:
        ubo = SU.ubo_with_camera(sc, cam, look, max_depth=8, random_state=5)
        rt = Raytracer(device, W, H)
        rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
        rt.build_bvh(ubo)
        rt.clear_image(); rt.raytrace(ubo, spp, flags=capi.TRACE_EXACT_NODES | capi.TRACE_NO_PRIMARY_SHARING); device.wait_idle()
        exact = rt.read_image()
        for fl in (0, capi.TRACE_NO_PRIMARY_SHARING, capi.TRACE_REFERENCE_ORDER, capi.TRACE_CULLED):
            rt.clear_image(); rt.raytrace(ubo, spp, flags=fl); device.wait_idle()
            bad = int((_bits(rt.read_image()) != _bits(exact)).any(axis=-1).sum())
            assert bad == 0, f"{n_big} big primitives, camera {cam}, flags {fl}: {bad} pixels differ from the exact-record walk"


@pytest.mark.parametrize("seed", range(10))
def test_traversal_hierarchy_fuzz(device, seed):
    """Scenes of >= 512 primitives are walked over the library's own traversal hierarchy (traversal_tree.cu: adaptive-axis codes relative to
    the true bounds, big leaves in front of the root).  Adversarial inputs for THAT build: clouds far from the origin (|coordinate| up to 1e6:
    the reference's coord / span codes saturate there, the hierarchy's must not care), flat and needle-shaped extents (all code bits on one or
    two axes), thousands of coincident centres (equal codes: index tie-break), every primitive big (nothing hoisted), slivers, nested spheres.
    The frame must equal the exact-record walk's bit for bit, shared and unshared primaries and in the reference's order."""
    from raytracergpu_mastersproject_b200 import Raytracer, capi
    rng = np.random.default_rng(4000 + seed)
    off_mag = [0.0, 1.0e3, 1.0e5, 1.0e6, 0.0][seed % 5]
    shape = [(1, 1, 1), (1, 0.02, 1), (1, 1e-4, 1e-4), (1, 1, 0.0), (0.3, 1, 0.05)][(seed // 2) % 5]     # extent of the cloud per axis
    ext = max(400.0, off_mag * 0.01)
    offset = rng.normal(size=3); offset = offset / np.linalg.norm(offset) * off_mag
    nt, ns = 8300 + int(rng.integers(0, 600)), int(rng.integers(0, 300))
    T = np.zeros(nt + 2, O.TRIANGLE); S = np.zeros(ns, O.SPHERE)
    c = offset + rng.uniform(-0.5, 0.5, (nt, 3)) * ext * np.array(shape)
    if seed % 3 == 1:
        c[: nt // 3] = c[0]                                                  # thousands of coincident centres
    size = ext * 10.0 ** rng.uniform(-3.2, -1.2 if seed != 7 else 0.2, (nt, 1))      # seed 7: most triangles are BIG (nothing can be hoisted)
    v0 = c + rng.normal(size=(nt, 3)) * size; v1 = c + rng.normal(size=(nt, 3)) * size; v2 = c + rng.normal(size=(nt, 3)) * size
    sl = rng.random(nt) < 0.1
    v2 = np.where(sl[:, None], v0 + (v1 - v0) * rng.uniform(0.2, 0.8, (nt, 1)) + rng.normal(size=(nt, 3)) * size * 10.0 ** rng.uniform(-2.7, -1, (nt, 1)), v2)
    T["v0"][:nt, :3] = v0; T["v1"][:nt, :3] = v1; T["v2"][:nt, :3] = v2
    T["materialIndex"][:nt] = rng.integers(1, 4, nt)
    lo, hi = offset - 0.6 * ext, offset + 0.6 * ext
    T["v0"][nt, :3] = (lo[0], hi[1], lo[2]); T["v1"][nt, :3] = (hi[0], hi[1], lo[2]); T["v2"][nt, :3] = (hi[0], hi[1], hi[2])
    T["v0"][nt + 1, :3] = (lo[0], hi[1], lo[2]); T["v1"][nt + 1, :3] = (hi[0], hi[1], hi[2]); T["v2"][nt + 1, :3] = (lo[0], hi[1], hi[2])
    if ns:
        S["center"][:, :3] = offset + rng.uniform(-0.45, 0.45, (ns, 3)) * ext * np.array(shape)
        S["radius"] = ext * 10.0 ** rng.uniform(-2.5, -1.0, ns); S["materialIndex"] = rng.integers(1, 4, ns)
        S["center"][ns // 2:, :3] = S["center"][: ns - ns // 2, :3]           # nested / concentric
    M = np.zeros(1, O.MODEL); M["m"][0] = np.eye(4, dtype=np.float32).reshape(16)
    MT = np.zeros(4, O.MATERIAL)
    for i, (a, t) in enumerate([((12, 12, 12), 0), ((0.7, 0.7, 0.7), 1), ((0.6, 0.3, 0.2), 1), ((0.2, 0.5, 0.7), 2 if seed % 2 else 1)]):
        MT["albedo"][i, :3] = a; MT["materialType"][i] = t
    sc = dict(models=M, triangles=T, spheres=S, materials=MT)
    d = rng.normal(size=3); d /= np.linalg.norm(d)
    W, H, spp = 56, 40, 3
    for cam, look in [(offset + d * ext * 1.5, offset), (offset + rng.uniform(-0.1, 0.1, 3) * ext * np.array(shape), offset + d * ext)]:
        ubo = SU.ubo_with_camera(sc, cam, look, max_depth=8, random_state=seed + 11)
        rt = Raytracer(device, W, H)
        rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
        rt.build_bvh(ubo)
        rt.clear_image(); rt.raytrace(ubo, spp, flags=capi.TRACE_EXACT_NODES | capi.TRACE_NO_PRIMARY_SHARING); device.wait_idle()
        exact = rt.read_image()
        for fl in (0, capi.TRACE_NO_PRIMARY_SHARING, capi.TRACE_REFERENCE_ORDER):
            rt.clear_image(); rt.raytrace(ubo, spp, flags=fl); device.wait_idle()
            bad = int((_bits(rt.read_image()) != _bits(exact)).any(axis=-1).sum())
            assert bad == 0, f"seed {seed}, camera {cam}, flags {fl}: {bad} pixels differ from the exact-record walk"
