"""CPU: the oracle against the known-answer vectors derived by hand from the shader sources
and against structural properties.  The reference ships no golden vectors of its own; the stronger pin -- outputs of the
reference's compiled shaders -- lives in tests/test_spirv_golden.py, these vectors complement it."""
import math
import struct

import numpy as np
import pytest

import scene_util as SU
from oracle import oracle as O


def f32bits(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


# ---------------------------------------------------------------- RNG (random.glsl:10-27)
def test_pcg_sequence_from_zero():
    s = 0; states = []; words = []; floats = []
    for _ in range(4):
        f, s = O.pcg_float(s)
        states.append(s); words.append(O.pcg_word(s)); floats.append(f)
    assert states == [0x00000001, 0x2C9277B6, 0x27003DAF, 0xF1A5F5BC]
    assert words == [0x108EF29B, 0x00033628, 0xDCCC2102, 0xD3BB3506]
    assert np.allclose(floats, [0.06468121, 4.900433e-05, 0.86248976, 0.8270753], rtol=1e-7, atol=0)


def test_seed_formula_and_known_pixel():
    s = O.seed_base(175, 650, 12345)
    assert s == 0x4DBEE054 == ((600 * 175 + 650) * (12345 + 1)) & 0xFFFFFFFF
    w = []
    for _ in range(4):
        _, s = O.pcg_float(s); w.append(O.pcg_word(s))
    assert w == [0x0DD0D599, 0xEE2725C8, 0xF2C5541F, 0x08A7B69F]
    _, s = O.pcg_float(0xFFFFFFFF)
    assert O.pcg_word(s) == 0x106EE0AB
    assert O.seed_base(0, 600, 7) == O.seed_base(1, 0, 7)          # U12: 600*x + y collides
    assert O.seed_base(0, 0, 99) == 0


def test_random_float_is_exact_scaling_and_saturates_to_one():
    # float(word)/4294967295.0f == float(word) * 2^-32 ; words >= 0xFFFFFF80 give exactly 1.0
    for st in (1, 12345, 0xDEADBEEF):
        f, s2 = O.pcg_float(st)
        assert f == float(np.float32(np.float32(O.pcg_word(s2)) * np.float32(2.0 ** -32)))
    assert float(np.float32(0xFFFFFF80)) * 2.0 ** -32 == 1.0


def test_alpha_reseed_pin_U1():
    assert O.alpha_to_u32(1.0) == 0xFFFFFFFF                       # cleared image: alpha 1.0 saturates
    assert O.alpha_to_u32(0.5) == 0x80000000
    assert O.alpha_to_u32(0.0) == 0
    a = float(np.float32(0x12345600) * np.float32(2.0 ** -32))
    assert O.alpha_to_u32(a) == 0x12345600                         # exact round trip of a stored nextRandom
    assert O.float_to_u32_sat(-1.5) == 0 and O.float_to_u32_sat(float("nan")) == 0
    assert O.float_to_u32_sat(4294967296.0) == 0xFFFFFFFF and O.float_to_u32_sat(1023.99) == 1023


def test_fp32_constants():
    assert f32bits(3.1415926535897932385) == 0x40490FDB
    assert f32bits(0.001) == 0x3A83126F and f32bits(10000000.0) == 0x4B189680


def test_pinned_sincos_accuracy_and_quadrants():
    xs = np.linspace(0, 2 * math.pi, 4001).astype(np.float32)
    err = 0.0
    for x in xs:
        s, c = O.pin_sincos(x)
        err = max(err, abs(float(s) - math.sin(float(x))), abs(float(c) - math.cos(float(x))))
    assert err < 2.5e-7
    s, c = O.pin_sincos(0.0)
    assert float(s) == 0.0 and float(c) == 1.0


# ---------------------------------------------------------------- Morton (GenerateMortonCodesOfPrimitives.comp:41-58)
def test_morton_vectors():
    assert O.separate_bits(1023) == O.separate_bits(1024) == 0x09249249
    assert (O.morton3(1, 0, 0), O.morton3(0, 1, 0), O.morton3(0, 0, 1)) == (1, 2, 4)
    assert O.morton3(1023, 1023, 1023) == 0x3FFFFFFF
    assert O.morton3(512, 256, 128) == 0x0A800000
    assert O.morton3(5, 9, 1) == 0x447


def test_morton_uses_coord_over_span_pin_U5():
    sc = SU.random_scene(3, n_tris=0, n_spheres=6, room=False)
    tw, sw = O.model_to_world(sc["models"], sc["triangles"], sc["spheres"])
    enc = O.enclosing_aabb(tw, sw)
    codes = O.morton_codes(tw, sw, enc)
    span = enc["eMax"][0, :3] - enc["eMin"][0, :3]
    for i, s in enumerate(sw):
        q = np.float32(s["center"][:3] / span) * np.float32(1024.0)
        q = [O.float_to_u32_sat(v) for v in q]
        assert codes["code"][i] == O.morton3(*q)
        assert codes["primitiveIndex"][i] == i and codes["primitiveType"][i] == 0


def test_enclosing_box_pin_U4_includes_origin_and_pads():
    sc = SU.random_scene(4, n_tris=10, n_spheres=0, room=False)
    tw, sw = O.model_to_world(sc["models"], sc["triangles"], sc["spheres"])
    enc = O.enclosing_aabb(tw, sw)
    cen = (tw["v0"][:, :3] + tw["v1"][:, :3] + tw["v2"][:, :3]) / np.float32(3)
    assert np.all(enc["eMin"][0, :3] <= np.minimum(0, cen.min(0)) + 1e-3)
    assert np.all(enc["eMax"][0, :3] >= np.maximum(0, cen.max(0)) - 1e-3)
    assert enc["eMin"][0, 3] == 0 and enc["eMax"][0, 3] == 0
    inf = O.enclosing_aabb(tw, sw, O.make_options(enclosing_init_inf=True))
    assert np.allclose(inf["eMin"][0, :3], cen.min(0)) and np.allclose(inf["eMax"][0, :3], cen.max(0))


# ---------------------------------------------------------------- sort + LBVH (RadixSortSimple / ConstructHLBVH / ConstructAABBs)
def test_radix_sort_is_stable_ascending():
    rng = np.random.default_rng(0)
    m = np.zeros(5000, O.MORTON)
    m["code"] = rng.integers(0, 50, 5000)            # heavy duplicates
    m["primitiveIndex"] = np.arange(5000)
    s = O.radix_sort(m)
    order = np.argsort(m["code"], kind="stable")
    assert np.array_equal(s["code"], m["code"][order]) and np.array_equal(s["primitiveIndex"], order)


def _check_tree(nodes, n):
    if n == 1:
        assert nodes["leftIndex"][0] == 0 and nodes["rightIndex"][0] == 0
        return 0
    seen = np.zeros(2 * n - 1, np.int32)
    depth_max = 0
    stack = [(0, 0)]
    while stack:
        i, d = stack.pop()
        seen[i] += 1
        depth_max = max(depth_max, d)
        if i < n - 1:
            l, r = int(nodes["leftIndex"][i]), int(nodes["rightIndex"][i])
            assert l != 0 and r != 0                                   # 0 is the "invalid index" sentinel and the root
            a, bl, br = nodes["aabb"][i], nodes["aabb"][l], nodes["aabb"][r]
            for k in (0, 2, 4):                                        # min lanes
                assert a[k] == min(bl[k], br[k]) and a[k + 1] == max(bl[k + 1], br[k + 1])   # refit = exact union
            stack += [(l, d + 1), (r, d + 1)]
        else:
            assert nodes["leftIndex"][i] == 0 and nodes["rightIndex"][i] == 0
    assert np.all(seen == 1), "every node must be reached exactly once from the root"
    return depth_max


@pytest.mark.parametrize("n", [1, 2, 3, 7, 1000])
def test_lbvh_topology_valid(n):
    sc = SU.random_scene(n, n_tris=n, n_spheres=0, room=False)
    b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    d = _check_tree(b["nodes"], n)
    assert d <= 62
    if n == 2:
        assert (b["nodes"]["leftIndex"][0], b["nodes"]["rightIndex"][0]) == (1, 2)
    # leaves are in ORIGINAL primitive order (D8)
    assert np.array_equal(b["nodes"]["primitiveIndex"][n - 1:], np.arange(n))
    assert np.all(b["nodes"]["primitiveType"][n - 1:] == 1)


def test_lbvh_with_heavy_duplicate_codes():
    sc = SU.random_scene(9, n_tris=300, n_spheres=0, room=False)
    t = sc["triangles"]
    t["v0"][:] = t["v0"][0]; t["v1"][:] = t["v1"][0]; t["v2"][:] = t["v2"][0]; t["modelIndex"][:] = t["modelIndex"][0]
    b = O.build_bvh(sc["models"], t, sc["spheres"])
    assert len(set(b["morton"]["code"].tolist())) == 1
    _check_tree(b["nodes"], 300)                     # ties are broken by position (ConstructHLBVH.comp:64-67)


def test_delta_function():
    m = np.zeros(4, O.MORTON); m["code"] = [1, 1, 2, 0x40000000]
    assert O.delta(m, 0, -1) == -1 and O.delta(m, 0, 4) == -1
    assert O.delta(m, 0, 1) == 32 + 31 - 0          # equal codes: 32 + clz(i ^ j)
    assert O.delta(m, 1, 2) == 31 - 1               # codes 1 ^ 2 = 3 -> msb 1
    assert O.delta(m, 0, 3) == 31 - 30


# ---------------------------------------------------------------- trace
def test_bvh_hits_agree_with_bruteforce_up_to_exact_ties():
    """raytrace.comp's linear scan is an independent oracle for the closest hit: ids may differ only where t ties (U9)."""
    sc = SU.random_scene(21, n_tris=400, n_spheres=40)
    ubo = SU.make_ubo(sc)
    b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    r = O.raytrace(ubo, 120, 90, b["tris"], b["sphs"], sc["materials"], b["nodes"], 1)
    hp, ht = O.primary_hits_bruteforce(ubo, 120, 90, b["tris"], b["sphs"])
    assert np.array_equal(ht.view(np.uint32), r["hit_t"].view(np.uint32))
    differ = hp != r["hit_prim"]
    assert differ.mean() < 0.02


def test_non_bvh_program_agrees_with_bvh_program_on_primary_hits():
    """raytrace.comp (linear scan) and raytraceBVH.comp find the same closest primary hit t; the image differs only by the
    background colour (0.1,0.1,0.3) and by exact-t tie-breaks."""
    sc = SU.random_scene(24, n_tris=150, n_spheres=15)
    ubo = SU.make_ubo(sc, max_depth=1)
    b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    a = O.raytrace(ubo, 80, 60, b["tris"], b["sphs"], sc["materials"], b["nodes"], 1)
    l = O.raytrace(ubo, 80, 60, b["tris"], b["sphs"], sc["materials"], None, 1, opt=O.make_options(linear_scan=True))
    assert np.array_equal(a["hit_t"].view(np.uint32), l["hit_t"].view(np.uint32))
    miss = l["hit_prim"] == 0xFFFFFFFF
    assert np.all(l["image"][miss][:, :3] == np.float32([0.1, 0.1, 0.3])) and np.all(a["image"][miss][:, :3] == 0)
    assert l["counters"]["nodeVisits"] == 0 and l["counters"]["triTests"] == l["counters"]["rays"] * len(sc["triangles"])


def test_alpha_chain_and_accumulation_over_dispatches():
    sc = SU.random_scene(22, n_tris=100, n_spheres=10)
    ubo = SU.make_ubo(sc, random_state=777)
    b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    W, H = 40, 30
    r3 = O.raytrace(ubo, W, H, b["tris"], b["sphs"], sc["materials"], b["nodes"], 3)
    img = None
    for _ in range(3):                                # three separate dispatches over the same image
        rr = O.raytrace(ubo, W, H, b["tris"], b["sphs"], sc["materials"], b["nodes"], 1, image=img)
        img = rr["image"]
    assert np.array_equal(img.view(np.uint32), r3["image"].view(np.uint32))
    # the alpha written by a dispatch is random() of seed base + uint(previous alpha * 2^32)
    x, y = 7, 11
    a = 1.0
    for _ in range(3):
        f, _s = O.pcg_float((O.seed_base(x, y, 777) + O.alpha_to_u32(a)) & 0xFFFFFFFF)
        a = f
    assert np.float32(a) == r3["image"][y, x, 3]


def test_absorbing_materials_and_light_termination():
    """D4: METALLIC / DIELECTRIC absorb (radiance 0); LIGHT emits its albedo and terminates the path."""
    sc = SU.random_scene(23, n_tris=0, n_spheres=0)
    ubo = SU.make_ubo(sc)
    b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    base = O.raytrace(ubo, 64, 48, b["tris"], b["sphs"], sc["materials"], b["nodes"], 2)
    mats = sc["materials"].copy(); mats["materialType"][1:] = 2
    absorbed = O.raytrace(ubo, 64, 48, b["tris"], b["sphs"], mats, b["nodes"], 2)
    lit = base["hit_prim"] < 2                         # the two light triangles
    assert np.all(absorbed["image"][..., :3][~lit] == 0)
    assert np.all(absorbed["image"][..., :3][lit] == 30.0)    # 2 samples x albedo 15
    assert base["image"][..., :3].sum() > absorbed["image"][..., :3].sum()


def test_resolve_formula():
    img = np.zeros((1, 4, 4), np.float32)
    img[0, :, 0] = [0.0, 4.0, 1.0, 100.0]
    out = O.resolve_rgba8(img, 4)
    assert out[0, :, 0].tolist() == [0, 255, 128, 255]           # clamp(sqrt(x / 4), 0, 1) * 255 rounded
    assert np.all(out[..., 3] == 255)


def test_logistic_map_step():
    """logistic.comp:26-34: x' = x r (1 - x); plot at (int(r/4 W), int((1 - x') H)); out-of-image stores are discarded"""
    pts = np.float32([[0.5, 4.0], [0.5, 2.0], [0.25, 3.0], [0.0, 1.0]])
    img = np.zeros((10, 8, 4), np.uint8)
    O.logistic_step(pts, img)
    assert pts[:, 0].tolist() == [1.0, 0.5, 0.5625, 0.0]
    # (0.5, 4.0): x = int(1.0 * 8) = 8 -> outside; (0.5, 2.0): (4, 5); (0.25, 3.0): (6, int(0.4375 * 10) = 4); (0, 1): y = 10 -> outside
    lit = {(int(y), int(x)) for y, x in zip(*np.nonzero(img[..., 0]))}
    assert lit == {(5, 4), (4, 6)}
    assert img[5, 4].tolist() == [255, 255, 255, 255]
