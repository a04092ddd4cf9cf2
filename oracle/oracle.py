"""ctypes front-end of the CPU ORACLE (oracle/rt_oracle.c).

TEST INFRASTRUCTURE ONLY.  Pinned against the reference's own compiled shaders (tests/golden/spirv_*.npz, produced by
oracle/spirv_interp.py from shaders/compiled/*.spv; see rt_oracle.h for what that does and does not cover).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
Nothing under raytracergpu_mastersproject_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _cpu_tag() -> str:
    """The oracle is built -march=native: one build per distinct CPU feature set, so that a library built in
    the build container is never executed on a GPU box with another CPU (it is rebuilt there; gcc is in the image)."""
    import hashlib
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return hashlib.sha1(" ".join(sorted(line.split(":", 1)[1].split())).encode()).hexdigest()[:10]
    except OSError:
        pass
    return "generic"


_SO = os.path.join(_HERE, "_build", f"liboracle-{_cpu_tag()}.so")

# record layouts (shaders/include/definitions.glsl:6-77, VulkanWrapper/SceneTypes.hpp:32-123)
MODEL = np.dtype([("m", "<f4", (16,))])
TRIANGLE = np.dtype([("v0", "<f4", (4,)), ("v1", "<f4", (4,)), ("v2", "<f4", (4,)),
                     ("materialIndex", "<u4"), ("modelIndex", "<u4"), ("_pad", "<u4", (2,))])
SPHERE = np.dtype([("center", "<f4", (4,)), ("radius", "<f4"), ("materialIndex", "<u4"),
                   ("modelIndex", "<u4"), ("_pad", "<u4")])
MATERIAL = np.dtype([("albedo", "<f4", (4,)), ("materialType", "<u4"), ("_pad", "<u4", (3,))])
NODE = np.dtype([("aabb", "<f4", (6,)), ("leftIndex", "<u4"), ("rightIndex", "<u4"),
                 ("primitiveIndex", "<u4"), ("primitiveType", "<u4")])
MORTON = np.dtype([("code", "<u4"), ("primitiveIndex", "<u4"), ("primitiveType", "<u4")])
CINFO = np.dtype([("parent", "<u4"), ("visitationCount", "<i4")])
ENCLOSING = np.dtype([("eMin", "<f4", (4,)), ("eMax", "<f4", (4,))])
UBO = np.dtype([("camPos", "<f4", (4,)), ("camLookAt", "<f4", (4,)), ("camUpDir", "<f4", (4,)),
                ("verticalFOV", "<f4"), ("numTriangles", "<u4"), ("numSpheres", "<u4"),
                ("numMaterials", "<u4"), ("numLights", "<u4"), ("maxRayTraceDepth", "<u4"),
                ("randomState", "<u4"), ("_pad", "<u4")])
assert (MODEL.itemsize, TRIANGLE.itemsize, SPHERE.itemsize, MATERIAL.itemsize, NODE.itemsize,
        MORTON.itemsize, CINFO.itemsize, ENCLOSING.itemsize, UBO.itemsize) == (64, 64, 32, 32, 40, 12, 8, 32, 80)


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "nodeVisits", "triTests", "sphTests", "matReads", "samples")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Options(C.Structure):
    _fields_ = [("enclosingInitInf", C.c_int), ("extMaterials", C.c_int), ("threads", C.c_int), ("linearScan", C.c_int)]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("rt_oracle.c", "rt_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "OUT=" + os.path.relpath(_SO, _HERE)], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
        L.orc_pcg_step.restype = u32; L.orc_pcg_step.argtypes = [u32]
        L.orc_pcg_word.restype = u32; L.orc_pcg_word.argtypes = [u32]
        L.orc_pcg_float.restype = f32; L.orc_pcg_float.argtypes = [C.POINTER(u32)]
        L.orc_seed_base.restype = u32; L.orc_seed_base.argtypes = [u32, u32, u32]
        L.orc_alpha_to_u32.restype = u32; L.orc_alpha_to_u32.argtypes = [f32]
        L.orc_pin_sincos.restype = None; L.orc_pin_sincos.argtypes = [f32, C.POINTER(f32), C.POINTER(f32)]
        L.orc_morton3.restype = u32; L.orc_morton3.argtypes = [u32, u32, u32]
        L.orc_separate_bits.restype = u32; L.orc_separate_bits.argtypes = [u32]
        L.orc_float_to_u32_sat.restype = u32; L.orc_float_to_u32_sat.argtypes = [f32]
        L.orc_delta.restype = C.c_int; L.orc_delta.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.orc_model_to_world.restype = None; L.orc_model_to_world.argtypes = [vp, vp, u32, vp, u32]
        L.orc_enclosing_aabb.restype = None; L.orc_enclosing_aabb.argtypes = [vp, u32, vp, u32, vp, vp]
        L.orc_morton_codes.restype = None; L.orc_morton_codes.argtypes = [vp, u32, vp, u32, vp, vp]
        L.orc_radix_sort.restype = None; L.orc_radix_sort.argtypes = [vp, vp, u32]
        L.orc_construct_hlbvh.restype = None; L.orc_construct_hlbvh.argtypes = [vp, u32, vp, u32, vp, vp, vp]
        L.orc_refit_aabbs.restype = None; L.orc_refit_aabbs.argtypes = [vp, vp, u32]
        L.orc_build_bvh.restype = None; L.orc_build_bvh.argtypes = [vp, vp, u32, vp, u32, vp, vp, vp, vp, vp, vp]
        L.orc_clear_image.restype = None; L.orc_clear_image.argtypes = [vp, u32, u32]
        L.orc_raytrace.restype = C.c_int
        L.orc_raytrace.argtypes = [vp, vp, u32, u32, u32, u32, vp, vp, vp, vp, u32, vp, vp, vp, vp, vp]
        L.orc_primary_hits_bruteforce.restype = None
        L.orc_primary_hits_bruteforce.argtypes = [vp, u32, u32, vp, vp, vp, vp]
        L.orc_resolve_rgba8.restype = None; L.orc_resolve_rgba8.argtypes = [vp, u32, u32, u32, vp]
        L.orc_max_threads.restype = C.c_int; L.orc_max_threads.argtypes = []
        L.orc_logistic_step.restype = None; L.orc_logistic_step.argtypes = [vp, u32, vp, u32, u32, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    a = np.ascontiguousarray(a)
    assert a.dtype.itemsize == np.dtype(dt).itemsize, (a.dtype, dt)
    return a


# ---------------------------------------------------------------- scalar helpers (KAT surface)
def pcg_step(s): return int(lib().orc_pcg_step(s & 0xFFFFFFFF))
def pcg_word(s): return int(lib().orc_pcg_word(s & 0xFFFFFFFF))


def pcg_float(state):
    """returns (value, new_state)"""
    s = C.c_uint32(state & 0xFFFFFFFF)
    v = lib().orc_pcg_float(C.byref(s))
    return float(np.float32(v)), int(s.value)


def seed_base(x, y, rs): return int(lib().orc_seed_base(x, y, rs & 0xFFFFFFFF))
def alpha_to_u32(a): return int(lib().orc_alpha_to_u32(float(a)))
def morton3(x, y, z): return int(lib().orc_morton3(x, y, z))
def separate_bits(v): return int(lib().orc_separate_bits(v))
def float_to_u32_sat(f): return int(lib().orc_float_to_u32_sat(float(f)))


def pin_sincos(x):
    s, c = C.c_float(), C.c_float()
    lib().orc_pin_sincos(float(x), C.byref(s), C.byref(c))
    return np.float32(s.value), np.float32(c.value)


def delta(sorted_morton, i, j):
    m = _c(sorted_morton, MORTON)
    return int(lib().orc_delta(_p(m), len(m), i, j))


def max_threads(): return int(lib().orc_max_threads())


def host_threads() -> int:
    """All host cores this process may run on -- NOT omp_get_max_threads(), which launchers cap (torchrun exports OMP_NUM_THREADS=1).
    Pass it as make_options(threads=...) to time the oracle on the whole host."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------- S1
def make_options(enclosing_init_inf=False, ext_materials=False, threads=0, linear_scan=False):
    return Options(int(enclosing_init_inf), int(ext_materials), int(threads), int(linear_scan))


def model_to_world(models, tris, sphs):
    tris = _c(tris, TRIANGLE).copy(); sphs = _c(sphs, SPHERE).copy()
    lib().orc_model_to_world(_p(_c(models, MODEL)), _p(tris), len(tris), _p(sphs), len(sphs))
    return tris, sphs


def enclosing_aabb(tris_w, sphs_w, opt=None):
    out = np.zeros(1, ENCLOSING)
    opt = opt or make_options()
    lib().orc_enclosing_aabb(_p(_c(tris_w, TRIANGLE)), len(tris_w), _p(_c(sphs_w, SPHERE)), len(sphs_w),
                             C.addressof(opt), _p(out))
    return out


def morton_codes(tris_w, sphs_w, enc):
    out = np.zeros(len(tris_w) + len(sphs_w), MORTON)
    lib().orc_morton_codes(_p(_c(tris_w, TRIANGLE)), len(tris_w), _p(_c(sphs_w, SPHERE)), len(sphs_w),
                           _p(_c(enc, ENCLOSING)), _p(out))
    return out


def radix_sort(morton):
    m1 = _c(morton, MORTON).copy(); m2 = np.zeros_like(m1)
    lib().orc_radix_sort(_p(m1), _p(m2), len(m1))
    return m1


def construct_hlbvh(tris_w, sphs_w, sorted_morton):
    n = len(tris_w) + len(sphs_w)
    nodes = np.zeros(2 * n - 1, NODE); cinfo = np.zeros(2 * n - 1, CINFO)
    lib().orc_construct_hlbvh(_p(_c(tris_w, TRIANGLE)), len(tris_w), _p(_c(sphs_w, SPHERE)), len(sphs_w),
                              _p(_c(sorted_morton, MORTON)), _p(nodes), _p(cinfo))
    return nodes, cinfo


def refit_aabbs(nodes, cinfo, n):
    nodes = _c(nodes, NODE).copy(); cinfo = _c(cinfo, CINFO).copy()
    lib().orc_refit_aabbs(_p(nodes), _p(cinfo), n)
    return nodes, cinfo


def build_bvh(models, tris, sphs, opt=None):
    """Full S1.  Returns dict(tris, sphs (world space), enclosing, morton (sorted), nodes, cinfo)."""
    tris = _c(tris, TRIANGLE).copy(); sphs = _c(sphs, SPHERE).copy()
    n = len(tris) + len(sphs)
    enc = np.zeros(1, ENCLOSING); m1 = np.zeros(n, MORTON); m2 = np.zeros(n, MORTON)
    nodes = np.zeros(2 * n - 1, NODE); cinfo = np.zeros(2 * n - 1, CINFO)
    opt = opt or make_options()
    lib().orc_build_bvh(_p(_c(models, MODEL)), _p(tris), len(tris), _p(sphs), len(sphs), C.addressof(opt),
                        _p(enc), _p(m1), _p(m2), _p(nodes), _p(cinfo))
    return dict(tris=tris, sphs=sphs, enclosing=enc, morton=m1, nodes=nodes, cinfo=cinfo)


# ---------------------------------------------------------------- S2
def raytrace(ubo, W, H, tris_w, sphs_w, mats, nodes, spp, rows=None, image=None, opt=None,
             want_hits=True, want_rng=True):
    """`spp` dispatches of raytraceBVH.comp.  Returns dict(image[H,W,4], hit_prim, hit_t, rng, counters)."""
    if image is None:
        image = np.empty((H, W, 4), np.float32)
        lib().orc_clear_image(_p(image), W, H)
    y0, y1 = rows if rows is not None else (0, H)
    hit_prim = np.full((H, W), 0xFFFFFFFF, np.uint32) if want_hits else None
    hit_t = np.zeros((H, W), np.float32) if want_hits else None
    rng = np.zeros((H, W), np.uint32) if want_rng else None
    cnt = Counters()
    opt = opt or make_options()
    ubo = _c(ubo, UBO)
    rc = lib().orc_raytrace(_p(ubo), _p(image), W, H, y0, y1, _p(_c(tris_w, TRIANGLE)), _p(_c(sphs_w, SPHERE)),
                            _p(_c(mats, MATERIAL)), _p(_c(nodes, NODE)) if nodes is not None else None, spp, C.addressof(opt),
                            _p(hit_prim), _p(hit_t), _p(rng), C.addressof(cnt))
    if rc != 0:
        raise RuntimeError("oracle: traversal stack overflow (MAX_STACK_DEPTH 128)")
    return dict(image=image, hit_prim=hit_prim, hit_t=hit_t, rng=rng, counters=cnt.as_dict())


def primary_hits_bruteforce(ubo, W, H, tris_w, sphs_w):
    hit_prim = np.zeros((H, W), np.uint32); hit_t = np.zeros((H, W), np.float32)
    lib().orc_primary_hits_bruteforce(_p(_c(ubo, UBO)), W, H, _p(_c(tris_w, TRIANGLE)), _p(_c(sphs_w, SPHERE)),
                                      _p(hit_prim), _p(hit_t))
    return hit_prim, hit_t


def resolve_rgba8(image, rays_per_pixel):
    H, W, _ = image.shape
    out = np.zeros((H, W, 4), np.uint8)
    lib().orc_resolve_rgba8(_p(np.ascontiguousarray(image, np.float32)), W, H, rays_per_pixel, _p(out))
    return out


def logistic_step(points, image_rgba8, pixel_color=(1.0, 1.0, 1.0, 1.0)):
    """one dispatch of logistic.comp, in place: points float32 [n, 2] = (x, r); image uint8 [H, W, 4]"""
    H, W, _ = image_rgba8.shape
    col = np.asarray(pixel_color, np.float32)
    lib().orc_logistic_step(_p(points), len(points), _p(image_rgba8), W, H, _p(col))
