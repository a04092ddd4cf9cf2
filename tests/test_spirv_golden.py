"""Parity against THE REFERENCE'S OWN BINARIES.

tests/golden/spirv_*.npz hold the outputs of the reference's compiled shaders (shaders/compiled/*.spv), executed in the build
container by oracle/spirv_interp.py (generator: tests/golden/make_spirv_golden.py; every buffer after every dispatch of S1, the
image after every dispatch of S2, for both ray tracing programs, plus the logistic-map program and the fragment resolve).
  * CPU tests: the C oracle must reproduce every one of those buffers bit for bit, stage by stage (each stage is fed the
    golden output of the previous one, so a mismatch names the shader it belongs to) -- this is what pins the oracle;
  * GPU tests: the CUDA path, through the C-ABI, must reproduce them as well.
Nothing here reads /root/reference: the fixtures are committed.
"""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
ALL_CASES = sorted(os.path.basename(p)[len("spirv_"):-len(".npz")] for p in glob.glob(os.path.join(HERE, "golden", "spirv_*.npz"))
                   if not p.endswith(("spirv_logistic.npz", "spirv_c1pixels.npz", "spirv_c2pixels.npz")))
# the edge_* fixtures cover corners of the parameter space (N = 1, depth 0 / 1, wrapped seed, ...)
CASES = [c for c in ALL_CASES if not c.startswith("edge_")]


def _load(name):
    return np.load(os.path.join(HERE, "golden", f"spirv_{name}.npz"))


def _same(a, b):
    return np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes()


def test_fixture_set_is_complete():
    assert CASES == ["degenerate", "dups", "mesh", "room", "slivers", "sorted", "spheres", "three", "two"]
    assert [c for c in ALL_CASES if c.startswith("edge_")] == ["edge_chain6", "edge_depth0", "edge_depth1", "edge_fov90", "edge_one", "edge_seedwrap"]
    assert os.path.exists(os.path.join(HERE, "golden", "spirv_c1pixels.npz")) and os.path.exists(os.path.join(HERE, "golden", "spirv_c2pixels.npz"))
    assert os.path.exists(os.path.join(HERE, "golden", "spirv_logistic.npz"))


@pytest.mark.parametrize("name", ALL_CASES)
def test_oracle_build_stages_match_reference_binaries(name):
    g = _load(name)
    T, S = len(g["triangles"]), len(g["spheres"]); N = T + S
    # K1 ModelSpaceToWorldSpace.comp.spv
    tw, sw = O.model_to_world(g["models"], g["triangles"], g["spheres"])
    assert _same(tw, g["tris_w"]) and _same(sw, g["sphs_w"]), "K1 model -> world"
    # K2 GetEnclosingAABB.comp.spv (pin U4: uninitialised locals read as zero -- the binary's OpVariable has no initialiser)
    assert _same(O.enclosing_aabb(g["tris_w"], g["sphs_w"]), g["enclosing"]), "K2 enclosing box"
    # K3 GenerateMortonCodesOfPrimitives.comp.spv
    assert _same(O.morton_codes(g["tris_w"], g["sphs_w"], g["enclosing"]), g["morton_unsorted"]), "K3 Morton codes"
    # K4 RadixSortSimple.comp.spv (whole 12-byte records: the sort is stable)
    assert _same(O.radix_sort(g["morton_unsorted"]), g["morton"]), "K4 radix sort"
    # K5 ConstructHLBVH.comp.spv
    nodes, cinfo = O.construct_hlbvh(g["tris_w"], g["sphs_w"], g["morton"])
    assert _same(nodes, g["nodes_unfitted"]) and _same(cinfo, g["cinfo_unfitted"]), "K5 topology / leaves / parents"
    # K6 ConstructAABBsOfInternalNodes.comp.spv
    nodes2, cinfo2 = O.refit_aabbs(g["nodes_unfitted"], g["cinfo_unfitted"], N)
    assert _same(nodes2, g["nodes"]) and _same(cinfo2, g["cinfo"]), "K6 refit"
    # and the fused entry point
    b = O.build_bvh(g["models"], g["triangles"], g["spheres"])
    assert _same(b["nodes"], g["nodes"]) and _same(b["morton"], g["morton"]) and _same(b["enclosing"], g["enclosing"])


@pytest.mark.parametrize("name", ALL_CASES)
def test_oracle_trace_matches_reference_binaries(name):
    g = _load(name)
    W, H, spp = int(g["W"]), int(g["H"]), int(g["spp"])
    image = None
    for s in range(spp):        # one dispatch of raytraceBVH.comp.spv at a time: the image (rgb sum + alpha seed) after each
        r = O.raytrace(g["ubo"], W, H, g["tris_w"], g["sphs_w"], g["materials"], g["nodes"], 1, image=image, want_hits=False, want_rng=False)
        image = r["image"]
        assert np.array_equal(image.view(np.uint32), g["images_bvh"][s].view(np.uint32)), f"raytraceBVH.comp dispatch {s}"
    r = O.raytrace(g["ubo"], W, H, g["tris_w"], g["sphs_w"], g["materials"], g["nodes"], spp, want_hits=False, want_rng=False)
    assert np.array_equal(r["image"].view(np.uint32), g["images_bvh"][-1].view(np.uint32))
    if "images_linear" in g.files:      # raytrace.comp.spv, the non-BVH program
        r = O.raytrace(g["ubo"], W, H, g["tris_w"], g["sphs_w"], g["materials"], None, spp, opt=O.make_options(linear_scan=True),
                       want_hits=False, want_rng=False)
        assert np.array_equal(r["image"].view(np.uint32), g["images_linear"][-1].view(np.uint32)), "raytrace.comp"
    if "resolved" in g.files:           # SingleTriangleFullScreen.frag.spv + UNORM8 conversion
        want = np.floor(np.clip(g["resolved"], 0.0, 1.0).astype(np.float32) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
        assert np.array_equal(O.resolve_rgba8(g["images_bvh"][-1], spp), want), "fragment resolve"


@pytest.mark.parametrize("name", ALL_CASES)
def test_oracle_work_counters_match_reference_binaries(name):
    """The quantities bench.py's roofline bytes are built from (rays, node visits V, triangle / sphere tests Tt / St,
    material reads H) are the reference binary's own: counted as executions of sceneHit / AABBhitCheck / triangleHit / sphereHit /
    emitted in the interpreted raytraceBVH.comp.spv (raytrace.comp.spv for the non-BVH program, which has no box tests)."""
    g = _load(name)
    W, H, spp = int(g["W"]), int(g["H"]), int(g["spp"])
    c = O.raytrace(g["ubo"], W, H, g["tris_w"], g["sphs_w"], g["materials"], g["nodes"], spp, want_hits=False, want_rng=False)["counters"]
    assert [c["rays"], c["nodeVisits"], c["triTests"], c["sphTests"], c["matReads"]] == g["counts_bvh"].tolist()
    assert c["samples"] == W * H * spp
    if "counts_linear" in g.files:
        c = O.raytrace(g["ubo"], W, H, g["tris_w"], g["sphs_w"], g["materials"], None, spp, opt=O.make_options(linear_scan=True),
                       want_hits=False, want_rng=False)["counters"]
        want = g["counts_linear"].tolist()
        assert [c["rays"], c["triTests"], c["sphTests"], c["matReads"]] == [want[0], want[2], want[3], want[4]] and want[1] == 0


def test_golden_traces_are_not_trivial():
    """the fixtures exercise what they claim: hits, misses, multi-bounce paths, both primitive types, every material type"""
    g = _load("room")
    img = g["images_bvh"][-1]
    lit = (img[..., :3].sum(-1) > 0).sum()
    assert lit > 50 and lit < img.shape[0] * img.shape[1]
    assert len(np.unique(g["images_bvh"][0][..., 3])) > 0.9 * img.shape[0] * img.shape[1]      # alpha = per-pixel seed chain
    types = set()
    for n in ("room", "sorted", "spheres"):
        types |= set(_load(n)["materials"]["materialType"].tolist())
    assert types == {0, 1, 2, 3}            # light, diffuse, and the two absorbing types (D4) all occur
    r = O.raytrace(g["ubo"], int(g["W"]), int(g["H"]), g["tris_w"], g["sphs_w"], g["materials"], g["nodes"], int(g["spp"]))
    c = r["counters"]
    assert c["rays"] > 1.3 * c["samples"] and c["triTests"] > 0 and c["sphTests"] > 0
    d = _load("dups")
    codes = d["morton"]["code"]
    assert (np.diff(codes.astype(np.int64)) == 0).sum() >= 10, "dups must contain equal Morton codes"


def test_oracle_c1_matches_reference_binaries():
    """BASELINE config C1 (the default complexScene, 800x800): S1 of the reference binaries on the whole scene, S2 on 1 500 sampled
    pixels (incl. the shader's debug pixel), two dispatches."""
    g = _load("c1pixels")
    W, H, spp = int(g["W"]), int(g["H"]), int(g["spp"])
    b = O.build_bvh(g["models"], g["triangles"], g["spheres"])
    for k in ("tris_w", "sphs_w", "enclosing", "morton", "nodes", "cinfo"):
        assert _same(b[{"tris_w": "tris", "sphs_w": "sphs"}.get(k, k)], g[k]), k
    xs, ys = g["pixels"][:, 0], g["pixels"][:, 1]
    image = None
    for s in range(spp):
        image = O.raytrace(g["ubo"], W, H, b["tris"], b["sphs"], g["materials"], b["nodes"], 1, image=image, want_hits=False, want_rng=False)["image"]
        assert np.array_equal(image[ys, xs].view(np.uint32), g["values"][s].view(np.uint32)), f"dispatch {s}"
    # the scene arrays in the fixture are what the C++ host builds today (Scenes.cpp complexScene + flatten)
    from raytracergpu_mastersproject_b200 import scenes
    sc = scenes.load_scene("complexScene")
    assert _same(sc["triangles"], g["triangles"]) and _same(sc["models"], g["models"]) and _same(sc["materials"], g["materials"])


def _c2_scene_and_digests():
    import hashlib

    from raytracergpu_mastersproject_b200 import scenes
    g = _load("c2pixels")
    sc = scenes.load_scene(str(g["spec"]))
    want = dict(zip(g["digest_names"].tolist(), g["digests"].tolist()))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    for k in ("models", "triangles", "spheres", "materials"):
        assert sha(sc[k]) == want[k], f"the host's {g['spec']} scene changed: {k}"
    return g, sc, want, sha


def test_oracle_c2_generator_matches_reference_binaries():
    """BASELINE config C2's scene generator (displaced sphere in the simpleScene room) at 9 806 primitives, C2's 1920x1080 and depth 8:
    every S1 buffer of the reference binaries (as SHA-256) and 700 sampled pixels of two raytraceBVH.comp dispatches."""
    g, sc, want, sha = _c2_scene_and_digests()
    W, H, spp = int(g["W"]), int(g["H"]), int(g["spp"])
    T, S = len(sc["triangles"]), len(sc["spheres"]); N = T + S
    tw, sw = O.model_to_world(sc["models"], sc["triangles"], sc["spheres"])
    assert sha(tw) == want["tris_w"] and sha(sw) == want["sphs_w"]
    enc = O.enclosing_aabb(tw, sw); assert _same(enc, g["enclosing"])
    mu = O.morton_codes(tw, sw, enc); assert sha(mu) == want["morton_unsorted"]
    ms = O.radix_sort(mu); assert sha(ms) == want["morton"]
    n0, c0 = O.construct_hlbvh(tw, sw, ms); assert sha(n0) == want["nodes_unfitted"] and sha(c0) == want["cinfo_unfitted"]
    n1, c1 = O.refit_aabbs(n0, c0, N); assert sha(n1) == want["nodes"] and sha(c1) == want["cinfo"]
    xs, ys = g["pixels"][:, 0], g["pixels"][:, 1]
    image = None
    for s in range(spp):
        image = O.raytrace(g["ubo"], W, H, tw, sw, sc["materials"], n1, 1, image=image, want_hits=False, want_rng=False)["image"]
        assert np.array_equal(image[ys, xs].view(np.uint32), g["values"][s].view(np.uint32)), f"dispatch {s}"
    assert (g["values"][-1][:, :3].sum(-1) > 0).sum() >= 5        # some sampled paths reach the light


def test_oracle_logistic_matches_reference_binary():
    g = np.load(os.path.join(HERE, "golden", "spirv_logistic.npz"))
    pts = g["points0"].copy()
    H, W = g["plotted"].shape
    img = np.zeros((H, W, 4), np.uint8)
    for s in range(len(g["points"])):
        O.logistic_step(pts, img, tuple(float(x) for x in g["color"]))
        assert _same(pts, g["points"][s]), f"logistic.comp dispatch {s}"
    assert np.array_equal(img.sum(-1) > 0, g["plotted"])
    want = np.floor(g["color"] * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    assert (img[g["plotted"]] == want).all()


# ------------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ALL_CASES)
def test_cuda_matches_reference_binaries(device, name):
    from raytracergpu_mastersproject_b200 import Raytracer, capi
    g = _load(name)
    W, H, spp = int(g["W"]), int(g["H"]), int(g["spp"])
    T, S = len(g["triangles"]), len(g["spheres"]); N = T + S
    rt = Raytracer(device, W, H)
    rt.update_scene(g["models"], g["triangles"], g["spheres"], g["materials"])
    rt.build_bvh(g["ubo"])
    device.wait_idle()
    assert T == 0 or _same(rt.triangles.read(O.TRIANGLE, T), g["tris_w"])
    assert S == 0 or _same(rt.spheres.read(O.SPHERE, S), g["sphs_w"])
    assert _same(rt.enclosing.read(O.ENCLOSING, 1), g["enclosing"])
    assert _same(rt.morton1.read(O.MORTON, N), g["morton"])
    assert _same(rt.nodes.read(O.NODE, 2 * N - 1), g["nodes"])
    # progressive: one sample per submission == one dispatch of raytraceBVH.comp each
    rt.clear_image()
    for s in range(spp):
        rt.raytrace(g["ubo"], 1); device.wait_idle()
        assert np.array_equal(rt.read_image().view(np.uint32), g["images_bvh"][s].view(np.uint32)), f"dispatch {s}"
    # fused submission, every traversal variant
    variants = [0, capi.TRACE_WIDE_NODES, capi.TRACE_WIDE_NODES | capi.TRACE_REFERENCE_ORDER, capi.TRACE_EXACT_NODES,
                capi.TRACE_SIMPLE_KERNEL, capi.TRACE_NO_PRIMARY_SHARING]
    for fl in variants:
        rt.clear_image(); rt.raytrace(g["ubo"], spp, flags=fl); device.wait_idle()
        assert np.array_equal(rt.read_image().view(np.uint32), g["images_bvh"][-1].view(np.uint32)), f"flags {fl}"
    # the instrumented launch counts the reference binary's work (roofline bytes)
    rt.clear_image(); rt.counters.zero()
    rt.raytrace(g["ubo"], spp, flags=capi.TRACE_COUNT); device.wait_idle()
    c = rt.read_counters()
    assert [c["rays"], c["nodeVisits"], c["triTests"], c["sphTests"], c["matReads"]] == g["counts_bvh"].tolist(), "work counters"
    assert np.array_equal(rt.read_image().view(np.uint32), g["images_bvh"][-1].view(np.uint32))
    if "resolved" in g.files:
        want = np.floor(np.clip(g["resolved"], 0.0, 1.0).astype(np.float32) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
        assert np.array_equal(rt.resolve_rgba8(spp), want)
    if "images_linear" in g.files:
        rt.update_scene(g["models"], g["triangles"], g["spheres"], g["materials"])
        rt.prepare_linear(g["ubo"])
        rt.clear_image(); rt.raytrace(g["ubo"], spp, flags=capi.TRACE_LINEAR_SCAN); device.wait_idle()
        assert np.array_equal(rt.read_image().view(np.uint32), g["images_linear"][-1].view(np.uint32)), "raytrace.comp"


@pytest.mark.gpu
def test_cuda_c1_matches_reference_binaries(device):
    from raytracergpu_mastersproject_b200 import Raytracer, capi
    g = _load("c1pixels")
    W, H, spp = int(g["W"]), int(g["H"]), int(g["spp"])
    N = len(g["triangles"]) + len(g["spheres"])
    rt = Raytracer(device, W, H)
    rt.update_scene(g["models"], g["triangles"], g["spheres"], g["materials"])
    rt.build_bvh(g["ubo"])
    device.wait_idle()
    assert _same(rt.nodes.read(O.NODE, 2 * N - 1), g["nodes"]) and _same(rt.morton1.read(O.MORTON, N), g["morton"])
    assert _same(rt.enclosing.read(O.ENCLOSING, 1), g["enclosing"]) and _same(rt.cinfo.read(O.CINFO, 2 * N - 1), g["cinfo"])
    xs, ys = g["pixels"][:, 0], g["pixels"][:, 1]
    rt.clear_image()
    for s in range(spp):
        rt.raytrace(g["ubo"], 1); device.wait_idle()
        assert np.array_equal(rt.read_image()[ys, xs].view(np.uint32), g["values"][s].view(np.uint32)), f"dispatch {s}"
    for fl in (0, capi.TRACE_WIDE_NODES, capi.TRACE_EXACT_NODES, capi.TRACE_SIMPLE_KERNEL):
        rt.clear_image(); rt.raytrace(g["ubo"], spp, flags=fl); device.wait_idle()
        assert np.array_equal(rt.read_image()[ys, xs].view(np.uint32), g["values"][-1].view(np.uint32)), f"flags {fl}"


@pytest.mark.gpu
def test_cuda_c2_generator_matches_reference_binaries(device):
    from raytracergpu_mastersproject_b200 import Raytracer, capi
    g, sc, want, sha = _c2_scene_and_digests()
    W, H, spp = int(g["W"]), int(g["H"]), int(g["spp"])
    N = len(sc["triangles"]) + len(sc["spheres"])
    rt = Raytracer(device, W, H)
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    rt.build_bvh(g["ubo"])
    device.wait_idle()
    assert sha(rt.nodes.read(O.NODE, 2 * N - 1)) == want["nodes"] and sha(rt.morton1.read(O.MORTON, N)) == want["morton"]
    assert sha(rt.cinfo.read(O.CINFO, 2 * N - 1)) == want["cinfo"] and _same(rt.enclosing.read(O.ENCLOSING, 1), g["enclosing"])
    xs, ys = g["pixels"][:, 0], g["pixels"][:, 1]
    rt.clear_image()
    for s in range(spp):
        rt.raytrace(g["ubo"], 1); device.wait_idle()
        assert np.array_equal(rt.read_image()[ys, xs].view(np.uint32), g["values"][s].view(np.uint32)), f"dispatch {s}"
    for fl in (0, capi.TRACE_REFERENCE_ORDER, capi.TRACE_EXACT_NODES, capi.TRACE_NO_PRIMARY_SHARING):   # 9 806 primitives: the 4-ary walk is the default
        rt.clear_image(); rt.raytrace(g["ubo"], spp, flags=fl); device.wait_idle()
        assert np.array_equal(rt.read_image()[ys, xs].view(np.uint32), g["values"][-1].view(np.uint32)), f"flags {fl}"


@pytest.mark.gpu
def test_cuda_logistic_matches_reference_binary(device):
    import ctypes as C

    from raytracergpu_mastersproject_b200 import Buffer, capi
    g = np.load(os.path.join(HERE, "golden", "spirv_logistic.npz"))
    H, W = g["plotted"].shape
    n = len(g["points0"])
    pts = Buffer(device, 8, n); pts.write(g["points0"])
    img = Buffer(device, 4, W * H); img.zero()
    col = np.ascontiguousarray(g["color"], np.float32)
    for s in range(len(g["points"])):
        capi.check(capi.lib().rtb_logistic_step(device.handle, pts._p, n, img._p, W, H, col.ctypes.data_as(C.c_void_p)))
        device.wait_idle()
        assert _same(pts.read(np.float32, 2 * n).reshape(n, 2), g["points"][s]), f"logistic.comp dispatch {s}"
    out = img.read(np.uint8, W * H * 4).reshape(H, W, 4)
    assert np.array_equal(out.sum(-1) > 0, g["plotted"])


# ------------------------------------------------------------------------------------------------ the generator itself
_SPV = "/root/reference/RaytracerGPU_MastersProject/shaders/compiled"


@pytest.mark.skipif(not os.path.isdir(_SPV), reason="the reference tree only exists in the build container")
def test_interpreter_regenerates_committed_fixtures():
    """Build container only: re-run the reference binaries through oracle/spirv_interp.py on the two smallest cases and on the
    logistic program and compare with the committed fixtures -- the fixtures are reproducible from the reference tree."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_spirv_golden as M
    for name in ("two", "three"):
        out = M.generate(name)
        g = _load(name)
        for k in ("tris_w", "sphs_w", "enclosing", "morton_unsorted", "morton", "nodes_unfitted", "cinfo_unfitted", "nodes", "cinfo"):
            assert _same(out[k], g[k]), (name, k)
        assert np.array_equal(out["images_bvh"].view(np.uint32), g["images_bvh"].view(np.uint32)), name
    lg = M.run_logistic(M.LOGISTIC)
    g = np.load(os.path.join(HERE, "golden", "spirv_logistic.npz"))
    assert _same(lg["points"], g["points"]) and np.array_equal(lg["plotted"], g["plotted"])


@pytest.mark.skipif(not os.path.isdir(_SPV), reason="the reference tree only exists in the build container")
def test_interpreter_reads_the_reference_layouts():
    """the record layouts the whole repository relies on, read from the Offset / ArrayStride decorations of the reference binary"""
    from oracle.spirv_interp import DEC_ARRAY_STRIDE, DEC_OFFSET, Module
    m = Module(os.path.join(_SPV, "raytraceBVH.comp.spv"))
    b = m.bindings()
    assert m.local_size == (32, 32, 1) and sorted(b) == [0, 1, 2, 3, 4, 5, 6]

    def element(binding):       # struct { T data[]; } -> (stride, struct type id of T)
        st = m.types[b[binding][1]]
        arr = st[1][0]
        return m.decor[arr][DEC_ARRAY_STRIDE][0], m.types[arr][1]

    def offsets(tid):
        return [m.mdecor[(tid, i)][DEC_OFFSET][0] for i in range(len(m.types[tid][1]))]

    stride, tri = element(2); assert stride == 64 and offsets(tri) == [0, 16, 32, 48, 52]        # Triangle
    stride, sph = element(3); assert stride == 32 and offsets(sph) == [0, 16, 20, 24]            # Sphere
    stride, mat = element(4); assert stride == 32 and offsets(mat) == [0, 16]                    # Material
    stride, node = element(5); assert stride == 40 and offsets(node) == [0, 24, 28, 32, 36]      # HLBVHNode
    assert offsets(m.types[node][1][0]) == [0, 4, 8, 12, 16, 20]                                  # AABB: minX maxX minY maxY minZ maxZ
    assert offsets(b[0][1]) == [0, 16, 32, 48, 52, 56, 60, 64, 68, 72]                            # ParameterUBO (80 bytes)
