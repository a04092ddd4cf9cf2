"""ctypes binding of librtb200.so (include/rtb200.h) -- the product's only compute path.

There is no CPU fallback: if the shared library is missing (and cannot be built) or no CUDA device is present,
loading / context creation raises.  Record dtypes mirror the reference layouts byte for byte
(shaders/include/definitions.glsl:6-77 == VulkanWrapper/SceneTypes.hpp:32-123).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_DEFAULT_SO = os.path.join(_PKG, "librtb200.so")
_SO = _DEFAULT_SO
_HEADER = os.path.join(os.path.dirname(_PKG), "include", "rtb200.h")

MODEL = np.dtype([("m", "<f4", (16,))])
TRIANGLE = np.dtype([("v0", "<f4", (4,)), ("v1", "<f4", (4,)), ("v2", "<f4", (4,)),
                     ("materialIndex", "<u4"), ("modelIndex", "<u4"), ("_pad", "<u4", (2,))])
SPHERE = np.dtype([("center", "<f4", (4,)), ("radius", "<f4"), ("materialIndex", "<u4"),
                   ("modelIndex", "<u4"), ("_pad", "<u4")])
MATERIAL = np.dtype([("albedo", "<f4", (4,)), ("materialType", "<u4"), ("_pad", "<u4", (3,))])
BVH_NODE = np.dtype([("aabb", "<f4", (6,)), ("leftIndex", "<u4"), ("rightIndex", "<u4"),
                     ("primitiveIndex", "<u4"), ("primitiveType", "<u4")])
MORTON_PRIMITIVE = np.dtype([("code", "<u4"), ("primitiveIndex", "<u4"), ("primitiveType", "<u4")])
CONSTRUCTION_INFO = np.dtype([("parent", "<u4"), ("visitationCount", "<i4")])
ENCLOSING_BOX = np.dtype([("eMin", "<f4", (4,)), ("eMax", "<f4", (4,))])
UBO = np.dtype([("camPos", "<f4", (4,)), ("camLookAt", "<f4", (4,)), ("camUpDir", "<f4", (4,)),
                ("verticalFOV", "<f4"), ("numTriangles", "<u4"), ("numSpheres", "<u4"),
                ("numMaterials", "<u4"), ("numLights", "<u4"), ("maxRayTraceDepth", "<u4"),
                ("randomState", "<u4"), ("_pad", "<u4")])

LIGHT, DIFFUSE, METALLIC, DIELECTRIC = 0, 1, 2, 3
SPHERE_PRIMITIVE, TRIANGLE_PRIMITIVE = 0, 1
TRACE_COUNT, TRACE_EXT_MATERIALS, TRACE_ENCLOSING_INF, TRACE_SIMPLE_KERNEL, TRACE_LINEAR_SCAN, TRACE_CULLED = 1, 2, 4, 8, 16, 32
TRACE_STREAM_KERNEL, TRACE_COMPRESSED_NODES, TRACE_WIDE_NODES, TRACE_EXACT_NODES = 64, 128, 256, 512
TRACE_NO_PRIMARY_SHARING = 1024
TRACE_REFERENCE_ORDER = 2048
TRACE_WALK_COUNT = 4096

COUNTER_FIELDS = ("rays", "nodeVisits", "triTests", "sphTests", "matReads", "samples")
# rtb_walk_counters (16 x u64): what the production kernels fetch (RTB_TRACE_WALK_COUNT)
WALK_COUNTER_FIELDS = ("rays", "recordFetches", "leafBoxFetches", "triTests", "sphTests", "matReads", "items", "paths", "parked",
                       "laneSteps", "warpSteps", "tailRays", "tailTurns", "uniqueRecordFetches", "_r1", "_r2")


def walk_bytes(w: dict, primary_sharing: bool) -> dict:
    """Bytes the production trace kernels fetch / store for the counted work (DESIGN.md "roofline"): every figure is the size
    of the record the kernel loads, as laid out in HBM."""
    loads = (64 * w["recordFetches"] + 32 * w["leafBoxFetches"] + 64 * w["triTests"] + 20 * w["sphTests"] + 16 * w["matReads"]
             + (12 + (48 if primary_sharing else 0)) * w["items"] + 240 * w["parked"])
    stores = 12 * w["paths"] + 240 * w["parked"]
    return {"load_bytes": int(loads), "store_bytes": int(stores), "record_bytes": int(64 * w["recordFetches"]),
            "leaf_box_bytes": int(32 * w["leafBoxFetches"]), "primitive_bytes": int(64 * w["triTests"] + 20 * w["sphTests"])}


class TraceArgs(C.Structure):
    _fields_ = [("imageWidth", C.c_uint32), ("imageHeight", C.c_uint32), ("localRows", C.c_uint32),
                ("bandRows", C.c_uint32), ("bandFirst", C.c_uint32), ("bandStep", C.c_uint32),
                ("sampleSkip", C.c_uint32), ("sampleCount", C.c_uint32), ("flags", C.c_uint32), ("_pad", C.c_uint32),
                ("hitPrim", C.c_void_p), ("hitT", C.c_void_p), ("rngOut", C.c_void_p), ("counters", C.c_void_p),
                ("walkCounters", C.c_void_p)]


EXPORTS = [
    "rtb_last_error", "rtb_version", "rtb_device_count", "rtb_ctx_create", "rtb_ctx_destroy", "rtb_sync",
    "rtb_device_name", "rtb_sm_count", "rtb_alloc", "rtb_free", "rtb_upload", "rtb_download", "rtb_memset",
    "rtb_host_alloc", "rtb_host_free", "rtb_timer_start", "rtb_timer_stop_ms", "rtb_model_to_world",
    "rtb_enclosing_aabb", "rtb_morton_codes", "rtb_sort_morton", "rtb_build_hlbvh", "rtb_refit_aabbs",
    "rtb_build_bvh", "rtb_clear_image", "rtb_bind_trace_buffers", "rtb_raytrace", "rtb_resolve_rgba8",
    "rtb_launch_count", "rtb_logistic_step",
    "rtb_comm_unique_id", "rtb_comm_init_rank", "rtb_comm_destroy", "rtb_comm_info", "rtb_comm_all_gather", "rtb_comm_broadcast",
    "rtb_gather_tiles", "rtb_reduce_samples", "rtb_probe_gather", "rtb_download_async", "rtb_export_hit_slack",
]


class RtbError(RuntimeError):
    """Raised where the reference throws std::runtime_error("failed to ...") (e.g. RaytracerBVH.hpp:398-404)."""


def library_path() -> str:
    return _SO


def _stale() -> bool:
    srcs = [os.path.join(_PKG, "csrc", f) for f in os.listdir(os.path.join(_PKG, "csrc"))] + [_HEADER,
                                                                                               os.path.join(_PKG, "Makefile")]
    return (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)


def build(force: bool = False) -> str:
    """Compile librtb200.so for sm_100a in-tree (nvcc cross-compiles without a GPU).  Several processes may call this at once (the
    ranks of a torchrun launch): an exclusive file lock lets one of them build while the others wait and then find the library fresh;
    the Makefile links to a temporary name and renames, so a library that exists is always complete."""
    if not (force or _stale()):
        return _SO
    import fcntl
    with open(os.path.join(_PKG, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or _stale():
                subprocess.run(["make", "-C", _PKG, "-j8", "-s"], check=True)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if _SO == _DEFAULT_SO:
        try:
            build()      # mtime-gated: a no-op unless csrc/ is newer than the library (a stale .so must never be tested silently)
        except Exception as e:  # noqa: BLE001
            if not os.path.exists(_SO):
                raise RtbError(f"librtb200.so is missing and could not be built ({e}); there is no CPU fallback") from e
            raise RtbError(f"librtb200.so is older than its sources and could not be rebuilt ({e})") from e
    L = C.CDLL(_SO)
    vp, u32, sz = C.c_void_p, C.c_uint32, C.c_size_t
    L.rtb_last_error.restype = C.c_char_p; L.rtb_last_error.argtypes = []
    L.rtb_version.restype = C.c_int; L.rtb_version.argtypes = []
    sigs = {
        "rtb_device_count": [C.POINTER(C.c_int)],
        "rtb_ctx_create": [C.c_int, vp, C.POINTER(vp)],
        "rtb_ctx_destroy": [vp],
        "rtb_sync": [vp],
        "rtb_device_name": [vp, C.c_char_p, sz],
        "rtb_sm_count": [vp, C.POINTER(C.c_int)],
        "rtb_alloc": [vp, sz, C.POINTER(vp)],
        "rtb_free": [vp, vp],
        "rtb_upload": [vp, vp, vp, sz],
        "rtb_download": [vp, vp, vp, sz],
        "rtb_memset": [vp, vp, C.c_int, sz],
        "rtb_host_alloc": [sz, C.POINTER(vp)],
        "rtb_host_free": [vp],
        "rtb_timer_start": [vp],
        "rtb_timer_stop_ms": [vp, C.POINTER(C.c_float)],
        "rtb_model_to_world": [vp, vp, vp, vp, vp],
        "rtb_enclosing_aabb": [vp, vp, vp, vp, vp, u32],
        "rtb_morton_codes": [vp, vp, vp, vp, vp, vp],
        "rtb_sort_morton": [vp, vp, vp, vp],
        "rtb_build_hlbvh": [vp, vp, vp, vp, vp, vp, vp],
        "rtb_refit_aabbs": [vp, vp, vp, vp],
        "rtb_build_bvh": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, u32],
        "rtb_clear_image": [vp, vp, u32, u32],
        "rtb_bind_trace_buffers": [vp, vp, vp, vp, vp, vp],
        "rtb_raytrace": [vp, vp, vp, C.POINTER(TraceArgs)],
        "rtb_resolve_rgba8": [vp, vp, u32, u32, u32, vp],
        "rtb_launch_count": [vp, C.POINTER(C.c_uint64)],
        "rtb_logistic_step": [vp, vp, u32, vp, u32, u32, vp],
        "rtb_probe_gather": [vp, sz, C.POINTER(C.c_float)],
        "rtb_export_hit_slack": [vp, vp, sz, vp],
        "rtb_download_async": [vp, vp, vp, sz],
        "rtb_comm_unique_id": [vp],
        "rtb_comm_init_rank": [vp, C.c_int, C.c_int, vp],
        "rtb_comm_destroy": [vp],
        "rtb_comm_info": [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "rtb_comm_all_gather": [vp, vp, sz],
        "rtb_comm_broadcast": [vp, vp, sz, C.c_int],
        "rtb_gather_tiles": [vp, vp, u32, u32, u32, vp, u32, vp],
        "rtb_reduce_samples": [vp, vp, u32, u32, C.c_int, u32, vp],
    }
    for name, args in sigs.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise RtbError(lib().rtb_last_error().decode("utf-8", "replace"))
