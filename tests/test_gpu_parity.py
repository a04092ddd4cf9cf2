"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bar: BIT-EXACT for everything on this path -- node arrays, Morton records, enclosing box, hit primitive ids,
RNG states, work counters AND the fp32 radiance image (the kernels keep the shader's operation order and are
built with --fmad=false).  The only tolerance is in the sample-range sharding test, where the fp32 summation
order differs by construction (tolerance stated there).
"""
import ctypes as C

import numpy as np
import pytest

import scene_util as SU
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _rt(device, W, H):
    from raytracergpu_mastersproject_b200 import Raytracer
    return Raytracer(device, W, H)


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _assert_nodes_equal(got, ref):
    assert got.tobytes() == ref.tobytes(), _first_diff(got, ref)


def _first_diff(got, ref):
    g = np.frombuffer(got.tobytes(), np.uint32); r = np.frombuffer(ref.tobytes(), np.uint32)
    idx = np.nonzero(g != r)[0]
    return f"{len(idx)} differing words, first at word {idx[0] if len(idx) else -1}"


SCENES = [
    dict(seed=1, n_tris=200, n_spheres=20),
    dict(seed=2, n_tris=1500, n_spheres=100, sort_morton=True),
    dict(seed=3, n_tris=0, n_spheres=50, room=False),      # spheres only
    dict(seed=4, n_tris=64, n_spheres=0, room=False),      # triangles only
    dict(seed=5, n_tris=1, n_spheres=0, room=False),       # N == 1: the root is a leaf
    dict(seed=6, n_tris=1, n_spheres=1, room=False),       # N == 2
    dict(seed=7, n_tris=5000, n_spheres=300, sort_morton=True),
]


@pytest.mark.parametrize("cfg", SCENES, ids=lambda c: f"seed{c['seed']}")
def test_s1_stage_by_stage(device, cfg):
    """K1..K6 one dispatch at a time, each compared with the oracle's restatement of the same shader."""
    from raytracergpu_mastersproject_b200 import Buffer, capi
    L = capi.lib()
    sc = SU.random_scene(**cfg)
    ubo = SU.make_ubo(sc)
    T, S = len(sc["triangles"]), len(sc["spheres"]); N = T + S
    h = device.handle
    up = ubo.ctypes.data_as(C.c_void_p)
    bm = Buffer(device, 64, len(sc["models"])); bm.write(sc["models"])
    bt = Buffer(device, 64, max(T, 1)); bt.write(sc["triangles"])
    bs = Buffer(device, 32, max(S, 1)); bs.write(sc["spheres"])
    # K1
    capi.check(L.rtb_model_to_world(h, up, bm._p, bt._p, bs._p))
    tw, sw = O.model_to_world(sc["models"], sc["triangles"], sc["spheres"])
    assert bt.read(O.TRIANGLE, T).tobytes() == tw.tobytes()
    assert bs.read(O.SPHERE, S).tobytes() == sw.tobytes()
    # K2
    be = Buffer(device, 32, 1)
    capi.check(L.rtb_enclosing_aabb(h, up, be._p, bt._p, bs._p, 0))
    enc = O.enclosing_aabb(tw, sw)
    assert be.read(O.ENCLOSING, 1).tobytes() == enc.tobytes()
    # K3
    m1 = Buffer(device, 12, N); m2 = Buffer(device, 12, N)
    capi.check(L.rtb_morton_codes(h, up, be._p, bt._p, bs._p, m1._p))
    mc = O.morton_codes(tw, sw, enc)
    assert m1.read(O.MORTON, N).tobytes() == mc.tobytes()
    # K4
    capi.check(L.rtb_sort_morton(h, up, m1._p, m2._p))
    ms = O.radix_sort(mc)
    assert m1.read(O.MORTON, N).tobytes() == ms.tobytes()
    # K5
    bn = Buffer(device, 40, 2 * N - 1); bc = Buffer(device, 8, 2 * N - 1)
    bn.zero(); bc.zero()
    capi.check(L.rtb_build_hlbvh(h, up, bt._p, bs._p, m1._p, bn._p, bc._p))
    nodes0, cinfo0 = O.construct_hlbvh(tw, sw, ms)
    _assert_nodes_equal(bn.read(O.NODE, 2 * N - 1), nodes0)
    assert bc.read(O.CINFO, 2 * N - 1).tobytes() == cinfo0.tobytes()
    # K6
    capi.check(L.rtb_refit_aabbs(h, up, bn._p, bc._p))
    nodes1, cinfo1 = O.refit_aabbs(nodes0, cinfo0, N)
    _assert_nodes_equal(bn.read(O.NODE, 2 * N - 1), nodes1)
    assert bc.read(O.CINFO, 2 * N - 1).tobytes() == cinfo1.tobytes()
    device.wait_idle()


@pytest.mark.parametrize("cfg", SCENES, ids=lambda c: f"seed{c['seed']}")
@pytest.mark.parametrize("spp", [1, 5])
@pytest.mark.parametrize("kernel", ["wave", "wave-exact-nodes", "wave-wide-nodes", "wave-wide-reference-order", "simple"])
def test_frame_bit_exact(device, cfg, spp, kernel):
    """Whole frame (S1 fused + S2): node array, hit ids, RNG states, counters and the fp32 image, all bit-exact.
    Both trace kernels: the production warp-coherent one and the straightforward one kept for A/B measurements."""
    from raytracergpu_mastersproject_b200 import Buffer, capi
    kflag = {"simple": capi.TRACE_SIMPLE_KERNEL, "stream": capi.TRACE_STREAM_KERNEL, "wave-compressed-nodes": capi.TRACE_COMPRESSED_NODES, "wave-wide-nodes": capi.TRACE_WIDE_NODES,
             "wave-exact-nodes": capi.TRACE_EXACT_NODES,
             "wave-wide-reference-order": capi.TRACE_WIDE_NODES | capi.TRACE_REFERENCE_ORDER}.get(kernel, 0)
    W, H = 96, 72
    sc = SU.random_scene(**cfg)
    ubo = SU.make_ubo(sc, max_depth=8, random_state=12345 + cfg["seed"])
    T, S = len(sc["triangles"]), len(sc["spheres"]); N = T + S
    ref = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    rr = O.raytrace(ubo, W, H, ref["tris"], ref["sphs"], sc["materials"], ref["nodes"], spp)

    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    _assert_nodes_equal(rt.nodes.read(O.NODE, 2 * N - 1), ref["nodes"])
    assert rt.morton1.read(O.MORTON, N).tobytes() == ref["morton"].tobytes()
    assert rt.enclosing.read(O.ENCLOSING, 1).tobytes() == ref["enclosing"].tobytes()
    assert rt.triangles.read(O.TRIANGLE, T).tobytes() == ref["tris"].tobytes()
    hp = Buffer(device, 4, W * H); ht = Buffer(device, 4, W * H); rg = Buffer(device, 4, W * H)
    rt.clear_image()
    rt.counters.zero()
    rt.raytrace(ubo, spp, flags=capi.TRACE_COUNT | kflag, hit_prim=hp, hit_t=ht, rng_out=rg)
    device.wait_idle()
    img = rt.read_image()
    assert np.array_equal(hp.read(np.uint32).reshape(H, W), rr["hit_prim"]), "primary hit primitive ids differ"
    assert np.array_equal(_bits(ht.read(np.float32)).reshape(H, W), _bits(rr["hit_t"])), "primary hit t differs"
    assert np.array_equal(rg.read(np.uint32).reshape(H, W), rr["rng"]), "RNG state after the last sample differs"
    assert np.array_equal(_bits(img[..., 3]), _bits(rr["image"][..., 3])), "alpha seed chain differs"
    nbad = int((_bits(img) != _bits(rr["image"])).any(axis=-1).sum())
    assert nbad == 0, f"{nbad} pixels differ bitwise; max abs diff {np.nanmax(np.abs(img - rr['image']))}"
    assert rt.read_counters() == rr["counters"]
    # the non-instrumented kernel variant must give the same image
    # (for spp > 1 the wave kernels share each pixel's primary hit between its samples; once more with that switched off)
    for extra in ([0, capi.TRACE_NO_PRIMARY_SHARING] if kernel.startswith("wave") else [0]):
        rt.clear_image(); hp.zero(); ht.zero(); rg.zero()
        rt.raytrace(ubo, spp, flags=kflag | extra, hit_prim=hp, hit_t=ht, rng_out=rg)
        device.wait_idle()
        assert np.array_equal(_bits(rt.read_image()), _bits(rr["image"]))
        assert np.array_equal(hp.read(np.uint32).reshape(H, W), rr["hit_prim"])
        assert np.array_equal(_bits(ht.read(np.float32)).reshape(H, W), _bits(rr["hit_t"]))
        assert np.array_equal(rg.read(np.uint32).reshape(H, W), rr["rng"])


def test_multi_pass_bit_identical(device, make_device):
    """A submission whose per-sample slots exceed the scratch budget runs in several passes (C4 / C5 do at full size):
    same image, hit ids and RNG states as one pass; the shared primary hits of pass 1 serve the later passes."""
    from raytracergpu_mastersproject_b200 import Buffer, capi
    W, H, spp = 128, 96, 20
    sc = SU.random_scene(21, n_tris=9000, n_spheres=40)          # >= 512 primitives: 4-ary records
    ubo = SU.make_ubo(sc, random_state=77)
    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    hp = Buffer(device, 4, W * H); rg = Buffer(device, 4, W * H)
    rt.clear_image(); rt.raytrace(ubo, spp, hit_prim=hp, rng_out=rg); device.wait_idle()
    one = rt.read_image(); hp1 = hp.read(np.uint32); rg1 = rg.read(np.uint32)
    device = make_device(RTB_WAVE_SAMPLE_BUF_MB=1)               # 128 * 96 * 16 B = 192 KiB per sample -> passes of 5 spp
    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    hp = Buffer(device, 4, W * H); rg = Buffer(device, 4, W * H)
    for flags in (0, capi.TRACE_NO_PRIMARY_SHARING, capi.TRACE_EXACT_NODES, capi.TRACE_REFERENCE_ORDER):
        hp.zero(); rg.zero()
        rt.clear_image(); rt.raytrace(ubo, spp, flags=flags, hit_prim=hp, rng_out=rg); device.wait_idle()
        assert np.array_equal(_bits(rt.read_image()), _bits(one))
        assert np.array_equal(hp.read(np.uint32), hp1) and np.array_equal(rg.read(np.uint32), rg1)
    ref = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    rr = O.raytrace(ubo, W, H, ref["tris"], ref["sphs"], sc["materials"], ref["nodes"], spp)
    assert np.array_equal(_bits(one), _bits(rr["image"]))


@pytest.mark.parametrize("n_prims", [511, 512, 513, 640])
def test_wide_threshold_boundary(device, n_prims):
    """Scenes around RTB_WIDE_MIN (512 primitives): just below, the exact child pairs in the reference's order; from 512 on, the fused
    build and the walk's own hierarchy.  The frame, primary hit ids and RNG states are the oracle's on both sides of the switch."""
    sc = SU.random_scene(300 + n_prims, n_tris=n_prims - 46, n_spheres=40, sort_morton=True)
    N = len(sc["triangles"]) + len(sc["spheres"])
    assert N == n_prims, N
    ubo = SU.make_ubo(sc, random_state=5)
    _check_against_oracle(device, sc, ubo, 72, 48, 3, [0])


def test_progressive_equals_fused(device):
    """spp dispatches one at a time (the reference's loop) == one fused launch == skip+count split."""
    W, H = 64, 48
    sc = SU.random_scene(11, n_tris=300, n_spheres=30)
    ubo = SU.make_ubo(sc, random_state=99)
    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    rt.clear_image(); rt.raytrace(ubo, 6); device.wait_idle()
    fused = rt.read_image()
    rt.clear_image()
    for _ in range(6):
        rt.raytrace(ubo, 1)
    device.wait_idle()
    assert np.array_equal(_bits(rt.read_image()), _bits(fused))


def test_tile_sharding_bit_identical(device):
    """Tile mode: every rank renders its interleaved row bands; the re-assembled image is bit-identical."""
    W, H = 80, 50
    sc = SU.random_scene(12, n_tris=400, n_spheres=40)
    ubo = SU.make_ubo(sc, random_state=7)
    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    rt.clear_image(); rt.raytrace(ubo, 3); device.wait_idle()
    full = rt.read_image()
    world, band = 4, 4
    nb = (H + band - 1) // band
    per = (nb + world - 1) // world
    rows = per * band
    out = np.zeros_like(full)
    for r in range(world):
        rt.clear_image(rows)
        rt.raytrace(ubo, 3, rows=rows, band_rows=band, band_first=r, band_step=world)
        device.wait_idle()
        loc = rt.read_image(rows)
        for j in range(rows):
            y = ((j // band) * world + r) * band + j % band
            if y < H:
                out[y] = loc[j]
    assert np.array_equal(_bits(out), _bits(full))


def test_sample_range_sharding(device):
    """Sample-range mode: ranks render disjoint sample ranges with the seed chain fast-forwarded.
    The alpha chain / RNG states are bit-exact; the radiance sum differs only by fp32 summation order:
    tolerance |sum_ranges - sequential| <= 1e-5 * (1 + sequential) per channel."""
    from raytracergpu_mastersproject_b200 import Buffer
    W, H, spp, world = 64, 40, 16, 4
    sc = SU.random_scene(13, n_tris=300, n_spheres=30)
    ubo = SU.make_ubo(sc, random_state=4242)
    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    rg = Buffer(device, 4, W * H)
    rt.clear_image(); rt.raytrace(ubo, spp, rng_out=rg); device.wait_idle()
    seq = rt.read_image(); seq_rng = rg.read(np.uint32)
    acc = np.zeros((H, W, 3), np.float32)
    last = None
    for r in range(world):
        rt.clear_image()
        rt.raytrace(ubo, spp // world, sample_skip=r * (spp // world), rng_out=rg)
        device.wait_idle()
        part = rt.read_image()
        acc += part[..., :3]
        last = part
    assert np.array_equal(_bits(last[..., 3]), _bits(seq[..., 3])), "alpha chain after fast-forward differs"
    assert np.array_equal(rg.read(np.uint32), seq_rng)
    assert np.all(np.abs(acc - seq[..., :3]) <= 1e-5 * (1 + np.abs(seq[..., :3])))


def test_resolve_matches_oracle(device):
    W, H = 64, 48
    sc = SU.random_scene(14)
    ubo = SU.make_ubo(sc)
    rt = _rt(device, W, H)
    img = rt.do_iteration(sc, ubo, 4)
    got = rt.resolve_rgba8(4)
    assert np.array_equal(got, O.resolve_rgba8(img, 4))


def test_extension_materials_bit_exact(device):
    """Extension N1 (metal / dielectric scatter; NOT reference behaviour): CUDA == oracle, bit for bit."""
    from raytracergpu_mastersproject_b200 import capi
    W, H = 96, 72
    sc = SU.random_scene(15, n_tris=300, n_spheres=60)
    ubo = SU.make_ubo(sc, random_state=5)
    ref = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    rr = O.raytrace(ubo, W, H, ref["tris"], ref["sphs"], sc["materials"], ref["nodes"], 4, opt=O.make_options(ext_materials=True))
    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    rt.clear_image(); rt.raytrace(ubo, 4, flags=capi.TRACE_EXT_MATERIALS); device.wait_idle()
    img = rt.read_image()
    nbad = int((_bits(img) != _bits(rr["image"])).any(axis=-1).sum())
    assert nbad == 0, f"{nbad} pixels differ"


@pytest.mark.parametrize("cfg", [SCENES[0], SCENES[2], SCENES[4]], ids=lambda c: f"seed{c['seed']}")
def test_non_bvh_program_bit_exact(device, cfg):
    """N2: Config::Programs::Raytracer = K1 + raytrace.comp (linear scan over all primitives, background (0.1,0.1,0.3))."""
    from raytracergpu_mastersproject_b200 import Buffer, capi
    W, H, spp = 64, 48, 3
    sc = SU.random_scene(**cfg)
    ubo = SU.make_ubo(sc, random_state=31 + cfg["seed"])
    tw, sw = O.model_to_world(sc["models"], sc["triangles"], sc["spheres"])
    rr = O.raytrace(ubo, W, H, tw, sw, sc["materials"], None, spp, opt=O.make_options(linear_scan=True))
    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.prepare_linear(ubo)
    hp = Buffer(device, 4, W * H); rg = Buffer(device, 4, W * H)
    rt.clear_image(); rt.counters.zero()
    rt.raytrace(ubo, spp, flags=capi.TRACE_LINEAR_SCAN | capi.TRACE_COUNT, hit_prim=hp, rng_out=rg)
    device.wait_idle()
    img = rt.read_image()
    assert np.array_equal(_bits(img), _bits(rr["image"]))
    assert np.array_equal(hp.read(np.uint32).reshape(H, W), rr["hit_prim"])
    assert np.array_equal(rg.read(np.uint32).reshape(H, W), rr["rng"])
    assert rt.read_counters() == rr["counters"]
    # without BVH nodes bound only the linear scan may run
    with pytest.raises(capi.RtbError):
        rt.raytrace(ubo, 1)


@pytest.mark.parametrize("cfg", [SCENES[0], SCENES[1], SCENES[2], SCENES[6]], ids=lambda c: f"seed{c['seed']}")
def test_culled_extension_matches_exact(device, cfg):
    """RTB_TRACE_CULLED (extension, not the reference's traversal): same image / hit ids / RNG states as the exact mode on
    the test scenes, with fewer node visits."""
    from raytracergpu_mastersproject_b200 import Buffer, capi
    W, H, spp = 96, 72, 4
    sc = SU.random_scene(**cfg)
    ubo = SU.make_ubo(sc, random_state=55 + cfg["seed"])
    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    out = {}
    for name, fl in (("exact", 0), ("culled", capi.TRACE_CULLED)):
        hp = Buffer(device, 4, W * H); rg = Buffer(device, 4, W * H)
        rt.clear_image(); rt.counters.zero()
        rt.raytrace(ubo, spp, flags=capi.TRACE_COUNT | fl, hit_prim=hp, rng_out=rg)
        device.wait_idle()
        out[name] = (rt.read_image(), hp.read(np.uint32), rg.read(np.uint32), rt.read_counters())
    assert np.array_equal(_bits(out["exact"][0]), _bits(out["culled"][0]))
    assert np.array_equal(out["exact"][1], out["culled"][1]) and np.array_equal(out["exact"][2], out["culled"][2])
    assert out["culled"][3]["rays"] == out["exact"][3]["rays"] and out["culled"][3]["matReads"] == out["exact"][3]["matReads"]
    assert out["culled"][3]["nodeVisits"] <= out["exact"][3]["nodeVisits"]


def test_error_behaviour(device):
    """The reference throws std::runtime_error on failed submissions; the C-ABI returns non-zero + message."""
    from raytracergpu_mastersproject_b200 import RtbError, capi
    sc = SU.random_scene(16, n_tris=10, n_spheres=2)
    ubo = SU.make_ubo(sc)
    rt = _rt(device, 32, 32)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    bad = ubo.copy(); bad["numTriangles"] += 1
    rt.clear_image()
    with pytest.raises(RtbError):
        rt.raytrace(bad, 1)
    with pytest.raises(RtbError):
        capi.check(capi.lib().rtb_raytrace(device.handle, None, None, None))


# ------------------------------------------------------------------------------------------------------------------
# edge cases
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("depth", [0, 1, 2])
@pytest.mark.parametrize("kernel", ["wave", "simple"])
def test_shallow_depths(device, depth, kernel):
    """maxRayTraceDepth 0 (the bounce loop never runs: no rays, colour 0, the alpha chain still advances), 1 and 2."""
    from raytracergpu_mastersproject_b200 import capi
    W, H, spp = 40, 30, 3
    sc = SU.random_scene(41, n_tris=120, n_spheres=10)
    ubo = SU.make_ubo(sc, max_depth=depth, random_state=3)
    ref = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    rr = O.raytrace(ubo, W, H, ref["tris"], ref["sphs"], sc["materials"], ref["nodes"], spp)
    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    rt.clear_image(); rt.counters.zero()
    rt.raytrace(ubo, spp, flags=capi.TRACE_COUNT | {"simple": capi.TRACE_SIMPLE_KERNEL, "stream": capi.TRACE_STREAM_KERNEL}.get(kernel, 0))
    device.wait_idle()
    assert np.array_equal(_bits(rt.read_image()), _bits(rr["image"]))
    assert rt.read_counters() == rr["counters"]
    if depth == 0:
        assert rr["counters"]["rays"] == 0 and np.all(rr["image"][..., :3] == 0)


@pytest.mark.parametrize("size", [(1, 1), (7, 5), (33, 3), (130, 67)])
def test_ragged_image_sizes(device, size):
    """image sizes that are not multiples of the 8x4 work tiles, down to a single pixel"""
    W, H = size
    sc = SU.random_scene(42, n_tris=150, n_spheres=12)
    ubo = SU.make_ubo(sc, random_state=8)
    ref = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    rr = O.raytrace(ubo, W, H, ref["tris"], ref["sphs"], sc["materials"], ref["nodes"], 2)
    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    rt.clear_image(); rt.raytrace(ubo, 2); device.wait_idle()
    assert np.array_equal(_bits(rt.read_image()), _bits(rr["image"]))


def test_degenerate_and_extreme_geometry_nan_parity(device):
    """A zero-area triangle makes the shader's normalize() divide 0 by 0: NaN hits that poison closest-so-far.  Whatever the
    reference arithmetic does with it (as pinned), the kernels must do the same, bit for bit -- NaN payloads included.
    Also: a far-away huge triangle (large magnitudes) and a needle triangle."""
    from raytracergpu_mastersproject_b200 import capi
    W, H, spp = 48, 36, 3
    sc = SU.random_scene(43, n_tris=60, n_spheres=6)
    t = sc["triangles"]
    t["v1"][10] = t["v0"][10]; t["v2"][10] = t["v0"][10]                        # a point
    t["v2"][11] = t["v1"][11]                                                     # a segment
    t["v0"][12, :3] = (-1e5, -1e5, 3e5); t["v1"][12, :3] = (1e5, -1e5, 3e5); t["v2"][12, :3] = (0, 2e5, 3e5)
    t["v0"][13, :3] = (0.1, 0.2, 0.1); t["v1"][13, :3] = (0.1000001, 0.2, 0.1); t["v2"][13, :3] = (0.1, 0.9, 0.4)
    ubo = SU.make_ubo(sc, random_state=17)
    ref = O.build_bvh(sc["models"], t, sc["spheres"])
    rr = O.raytrace(ubo, W, H, ref["tris"], ref["sphs"], sc["materials"], ref["nodes"], spp)
    for fl in (0, capi.TRACE_SIMPLE_KERNEL):
        rt = _rt(device, W, H)
        rt.update_scene(sc["models"], t, sc["spheres"], sc["materials"])
        rt.build_bvh(ubo)
        n = len(t) + len(sc["spheres"])
        assert rt.nodes.read(O.NODE, 2 * n - 1).tobytes() == ref["nodes"].tobytes()
        rt.clear_image(); rt.counters.zero()
        rt.raytrace(ubo, spp, flags=capi.TRACE_COUNT | fl); device.wait_idle()
        assert np.array_equal(_bits(rt.read_image()), _bits(rr["image"]))
        assert rt.read_counters() == rr["counters"]
        if fl == 0:   # 4-ary records: the zero-area triangles make the hit-point slack infinite -> reference-order walk
            for fl2 in (capi.TRACE_WIDE_NODES, capi.TRACE_WIDE_NODES | capi.TRACE_REFERENCE_ORDER):
                rt.clear_image(); rt.raytrace(ubo, spp, flags=fl2); device.wait_idle()
                assert np.array_equal(_bits(rt.read_image()), _bits(rr["image"]))


def _check_against_oracle(device, sc, ubo, W, H, spp, flag_sets):
    from raytracergpu_mastersproject_b200 import Buffer
    ref = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    rr = O.raytrace(ubo, W, H, ref["tris"], ref["sphs"], sc["materials"], ref["nodes"], spp)
    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    hp = Buffer(device, 4, W * H); ht = Buffer(device, 4, W * H); rg = Buffer(device, 4, W * H)
    for fl in flag_sets:
        hp.zero(); ht.zero(); rg.zero()
        rt.clear_image(); rt.raytrace(ubo, spp, flags=fl, hit_prim=hp, hit_t=ht, rng_out=rg); device.wait_idle()
        assert np.array_equal(hp.read(np.uint32).reshape(H, W), rr["hit_prim"]), f"flags {fl}: primary hit ids differ"
        assert np.array_equal(_bits(ht.read(np.float32)).reshape(H, W), _bits(rr["hit_t"])), f"flags {fl}: primary hit t differs"
        assert np.array_equal(rg.read(np.uint32).reshape(H, W), rr["rng"]), f"flags {fl}: RNG states differ"
        assert np.array_equal(_bits(rt.read_image()), _bits(rr["image"])), f"flags {fl}: image differs"
    return rr


def test_nearest_first_equal_t_ties(device):
    """Nearest-first traversal tests leaves in a different order than the reference; equal-t hits must still resolve to the
    primitive the reference's order ends with (the smallest id).  Every triangle of the scene exists three times (same
    vertices, different ids and materials), spheres twice, and the room's quads are split along a diagonal that pixels hit."""
    from raytracergpu_mastersproject_b200 import capi
    W, H, spp = 96, 72, 4
    sc = SU.random_scene(51, n_tris=250, n_spheres=12)
    t = sc["triangles"]; s = sc["spheres"]
    t2 = t.copy(); t2["materialIndex"] = (t2["materialIndex"] + 1) % len(sc["materials"])
    t3 = t.copy(); t3["materialIndex"] = (t3["materialIndex"] + 3) % len(sc["materials"])
    s2 = s.copy(); s2["materialIndex"] = (s2["materialIndex"] + 2) % len(sc["materials"])
    sc = dict(sc, triangles=np.concatenate([t, t2, t3]), spheres=np.concatenate([s, s2]))
    ubo = SU.make_ubo(sc, random_state=5)
    rr = _check_against_oracle(device, sc, ubo, W, H, spp,
                               [capi.TRACE_WIDE_NODES, capi.TRACE_WIDE_NODES | capi.TRACE_REFERENCE_ORDER, capi.TRACE_EXACT_NODES])
    hp = rr["hit_prim"]; T0, S0 = len(t), len(s)
    assert (hp != 0xFFFFFFFF).mean() > 0.2
    tri = hp[hp < 3 * T0]; sph = hp[(hp >= 3 * T0) & (hp != 0xFFFFFFFF)]
    assert len(tri) and (tri < T0).all(), "the reference's order ends with the first copy of a triplicated triangle"
    assert (sph < 3 * T0 + S0).all()


@pytest.mark.parametrize("angle", [3e-2, 1e-3, 1e-5, 1e-7])
def test_nearest_first_sliver_triangles(device, angle):
    """Thin triangles stretch the slack between an accepted hit point and the triangle (the barycentric tests are evaluated with
    rounding): moderate slivers widen the culling boxes, extreme ones switch the scene to the reference-order walk.  Either
    way the frame is the oracle's, bit for bit."""
    from raytracergpu_mastersproject_b200 import capi
    W, H, spp = 80, 60, 3
    sc = SU.random_scene(61, n_tris=400, n_spheres=0)
    t = sc["triangles"]
    rng = np.random.default_rng(7)
    for i in range(20, len(t), 3):                       # every third random triangle becomes a sliver of the given apex angle
        v0 = t["v0"][i, :3]; e = t["v1"][i, :3] - v0
        perp = np.cross(e, rng.uniform(-1, 1, 3)); perp /= np.linalg.norm(perp) + 1e-30
        t["v2"][i, :3] = v0 + e * np.float32(0.9) + perp * np.float32(np.linalg.norm(e) * angle)
    ubo = SU.make_ubo(sc, random_state=23)
    _check_against_oracle(device, sc, ubo, W, H, spp, [capi.TRACE_WIDE_NODES, capi.TRACE_WIDE_NODES | capi.TRACE_REFERENCE_ORDER])


@pytest.mark.parametrize("scale", [1e-3, 1.0, 1e4])
def test_nearest_first_scene_scales(device, scale):
    """The slack terms scale with the scene: the same scene at three magnitudes (model matrices and camera scaled)."""
    from raytracergpu_mastersproject_b200 import capi
    W, H, spp = 80, 60, 3
    sc = SU.random_scene(71, n_tris=500, n_spheres=25)
    m = sc["models"]["m"].reshape(-1, 4, 4).copy()
    m[:, :, :3] *= np.float32(scale)                     # m[col][row]: scales rotation/scale columns and the translation
    sc["models"]["m"] = m.reshape(-1, 16)
    sc["spheres"]["radius"] *= np.float32(scale)
    ubo = SU.make_ubo(sc, random_state=31)
    ubo["camPos"][0, :3] *= np.float32(scale); ubo["camLookAt"][0, :3] *= np.float32(scale)
    _check_against_oracle(device, sc, ubo, W, H, spp, [capi.TRACE_WIDE_NODES, capi.TRACE_WIDE_NODES | capi.TRACE_REFERENCE_ORDER])


@pytest.mark.parametrize("spec,rs", [("meshRoom:70:5", 3), ("meshRoom:110:9", 4), ("heightField:90:70:6:40", 5),
                                     ("sphereField:12000:7", 6), ("simpleScene", 7), ("complexScene", 8)])
def test_nearest_first_equals_reference_order_mid_size(device, spec, rs):
    """The generators behind C2..C5 at sizes where a 320x240 / 8 spp / depth 8 frame takes milliseconds: the nearest-first
    t-culled walk (library default for these sizes, and forced on the two small reference scenes), the same records in the
    reference's order, and the exact records must agree on the whole frame, the primary hit ids and the RNG states; the
    exact-record frame is checked against the oracle on two rows."""
    from raytracergpu_mastersproject_b200 import Buffer, Raytracer, capi, make_ubo, scenes
    W, H, spp = 320, 240, 8
    sc = scenes.load_scene(spec)
    T, S = len(sc["triangles"]), len(sc["spheres"])
    ubo = make_ubo(T, S, len(sc["materials"]), 8, rs, sc["vfov"])
    rt = Raytracer(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    hp = Buffer(device, 4, W * H); rg = Buffer(device, 4, W * H)
    out = {}
    for name, fl in (("exact", capi.TRACE_EXACT_NODES), ("default", 0), ("nearest", capi.TRACE_WIDE_NODES),
                     ("ordered", capi.TRACE_WIDE_NODES | capi.TRACE_REFERENCE_ORDER),
                     ("nearest-unshared", capi.TRACE_WIDE_NODES | capi.TRACE_NO_PRIMARY_SHARING)):
        hp.zero(); rg.zero()
        rt.clear_image(); rt.raytrace(ubo, spp, flags=fl, hit_prim=hp, rng_out=rg); device.wait_idle()
        out[name] = (rt.read_image().copy(), hp.read(np.uint32), rg.read(np.uint32))
    for name in out:
        assert np.array_equal(_bits(out[name][0]), _bits(out["exact"][0])), f"{spec}: {name} image differs from the exact-record walk"
        assert np.array_equal(out[name][1], out["exact"][1]), f"{spec}: {name} primary hit ids differ"
        assert np.array_equal(out[name][2], out["exact"][2]), f"{spec}: {name} RNG states differ"
    ref = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    r = O.raytrace(ubo, W, H, ref["tris"], ref["sphs"], sc["materials"], ref["nodes"], spp, rows=(100, 102), want_hits=False, want_rng=False)
    assert np.array_equal(_bits(out["exact"][0][100:102]), _bits(r["image"][100:102]))


def test_empty_and_zero_sample_submissions(device):
    from raytracergpu_mastersproject_b200 import RtbError, capi
    sc = SU.random_scene(44, n_tris=20, n_spheres=2)
    ubo = SU.make_ubo(sc)
    rt = _rt(device, 16, 16)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    empty = ubo.copy(); empty["numTriangles"] = 0; empty["numSpheres"] = 0
    with pytest.raises(RtbError, match="empty scene"):
        rt.build_bvh(empty)
    rt.build_bvh(ubo)
    rt.clear_image(); rt.raytrace(ubo, 0); device.wait_idle()          # zero dispatches: the cleared image is untouched
    img = rt.read_image()
    assert np.all(img[..., :3] == 0) and np.all(img[..., 3] == 1)


@pytest.mark.parametrize("coop,turns", [(0, 1), (32, 0), (32, 1), (8, 3)])
def test_tail_handover_forced(make_device, coop, turns):
    """Long-ray / tail hand-over (trace_wave.cu: parked rays and paths finished by trace_tail_kernel, one ray per warp) with the
    thresholds forced so low that nearly every ray -- primary launch included -- takes that path: rays parked in flight after
    `turns` turns (with their stack and closest hit), paths parked at a ray boundary once the queue is drained.  Image, primary
    hit ids / t, RNG states must be the oracle's, with and without the material extension, shared and unshared primaries,
    1 and 6 samples (1 sample: the primary ray itself is walked by the main launch and may be parked at depth 0)."""
    from raytracergpu_mastersproject_b200 import capi
    device = make_device(RTB_WAVE_COOP=coop, RTB_WAVE_COOP_TURNS=turns)
    W, H = 80, 56
    sc = SU.random_scene(23, n_tris=900, n_spheres=120, sort_morton=True)
    ubo = SU.make_ubo(sc, random_state=77)
    for spp in (1, 6):
        _check_against_oracle(device, sc, ubo, W, H, spp, [capi.TRACE_WIDE_NODES, capi.TRACE_WIDE_NODES | capi.TRACE_NO_PRIMARY_SHARING])
    # extension materials through trace_tail_kernel<true>
    ref = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    rr = O.raytrace(ubo, W, H, ref["tris"], ref["sphs"], sc["materials"], ref["nodes"], 5, opt=O.make_options(ext_materials=True))
    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    rt.clear_image(); rt.raytrace(ubo, 5, flags=capi.TRACE_EXT_MATERIALS | capi.TRACE_WIDE_NODES); device.wait_idle()
    assert np.array_equal(_bits(rt.read_image()), _bits(rr["image"])), "extension materials through the tail kernel differ"


def test_tail_handover_off_equals_on(make_device):
    """the hand-over is a scheduling decision: switched off (RTB_WAVE_COOP=0 RTB_WAVE_COOP_TURNS=0) the frame is the same"""
    from raytracergpu_mastersproject_b200 import capi, make_ubo, scenes
    sc = scenes.load_scene("meshRoom:110:9")
    W, H, spp = 160, 90, 8
    ubo = make_ubo(len(sc["triangles"]), len(sc["spheres"]), len(sc["materials"]), sc["max_depth"], 11, sc["vfov"])
    frames = []
    for coop, turns, shape in ((0, 0, {}), (8, 32, {}), (32, 2, {}), (32, 2, dict(RTB_WAVE_MAIN_CTAS=6, RTB_WAVE_TAIL_THREADS=64))):
        device = make_device(RTB_WAVE_COOP=coop, RTB_WAVE_COOP_TURNS=turns, **shape)     # shape: launch-shape knobs (frames in flight)
        rt = _rt(device, W, H)
        rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
        rt.build_bvh(ubo)
        rt.clear_image(); rt.raytrace(ubo, spp); device.wait_idle()
        frames.append(rt.read_image())
        del rt
    assert all(np.array_equal(_bits(frames[0]), _bits(f)) for f in frames[1:])


@pytest.mark.parametrize("spec,spp,ext", [("meshRoom:70:5", 6, False), ("heightField:90:70:6:40", 4, True), ("sphereField:9000:7", 3, False)])
def test_walk_counters_of_the_production_kernels(device, spec, spp, ext):
    """RTB_TRACE_WALK_COUNT: the production walk (4-ary records, nearest-first, hand-over, shared primary hits) instrumented with
    what it fetches.  Same frame as the un-instrumented call; the counters obey the identities the walk implies and relate to the
    reference-equivalent counters (RTB_TRACE_COUNT) as primary-hit sharing says they must."""
    from raytracergpu_mastersproject_b200 import Buffer, capi, make_ubo, scenes
    sc = scenes.load_scene(spec)
    W, H = 160, 96
    ubo = make_ubo(len(sc["triangles"]), len(sc["spheres"]), len(sc["materials"]), sc["max_depth"], 5, sc["vfov"])
    rt = _rt(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    e = capi.TRACE_EXT_MATERIALS if ext else 0
    rt.clear_image(); rt.raytrace(ubo, spp, flags=e); device.wait_idle()
    plain = rt.read_image()
    wc = Buffer(device, 8, 16); wc.zero()
    rt.clear_image(); rt.raytrace(ubo, spp, flags=e | capi.TRACE_WALK_COUNT, walk_counters=wc); device.wait_idle()
    assert np.array_equal(_bits(rt.read_image()), _bits(plain)), "the instrumented production walk renders a different frame"
    w = dict(zip(capi.WALK_COUNTER_FIELDS, (int(x) for x in wc.read(np.uint64, 16))))
    rt.clear_image(); rt.counters.zero(); rt.raytrace(ubo, spp, flags=e | capi.TRACE_COUNT); device.wait_idle()
    assert np.array_equal(_bits(rt.read_image()), _bits(plain))
    ref = rt.read_counters()
    # active pixels = pixels whose primary ray enters the root box = items of the primary launch
    active = w["items"] // (spp + 1)
    assert w["items"] == active * (spp + 1), "items = one primary-hit item + spp sample items per active pixel"
    assert w["paths"] == active * spp, "every (pixel, sample) item ends in exactly one colour store"
    # rays walked = the reference's rays minus the (spp - 1) repeats of every active pixel's primary ray and minus the rays of
    # root-missing pixels (finished by the pre-pass without a walk)
    missed = W * H - active
    assert w["rays"] == ref["rays"] - active * (spp - 1) - missed * spp
    assert w["matReads"] == ref["matReads"], "every sample still shades its own primary hit"
    assert 0 < w["triTests"] + w["sphTests"] <= ref["triTests"] + ref["sphTests"], "t-culling can only drop primitive tests"
    assert w["leafBoxFetches"] >= w["triTests"] + w["sphTests"], "a primitive is tested only after its exact leaf box passed"
    assert w["recordFetches"] > 0 and w["laneSteps"] + 32 * w["tailTurns"] >= w["recordFetches"] and w["warpSteps"] * 32 >= w["laneSteps"]
    assert w["tailRays"] >= w["parked"] or w["parked"] == 0, "every parked path / ray is finished by the tail kernel"
    b = capi.walk_bytes(w, primary_sharing=True)
    assert b["load_bytes"] > b["record_bytes"] > 0


def test_ab_variants_are_not_in_the_product_build(device):
    """The streaming kernel and the 32-byte compressed records were measured and not adopted: the default build does not
    carry them and says so instead of silently running something else."""
    from raytracergpu_mastersproject_b200 import RtbError, capi
    sc = SU.random_scene(1, n_tris=50, n_spheres=5)
    ubo = SU.make_ubo(sc)
    rt = _rt(device, 16, 16)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(ubo)
    rt.clear_image()
    for fl in (capi.TRACE_STREAM_KERNEL, capi.TRACE_COMPRESSED_NODES):
        with pytest.raises(RtbError, match="A/B variants"):
            rt.raytrace(ubo, 1, flags=fl)
