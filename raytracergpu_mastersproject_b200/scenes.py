"""Scene arrays for Python callers (tests, bench.py), produced by the C++ host mirror (host/Scenes.cpp,
host/VulkanWrapper/RaytraceScene.cpp) through librtb200_host.so -- one implementation of the scene builders and of
the GameObject -> device-array flatten for both host languages.  Host-only: no device is touched here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import capi

_HOST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host")
_SO = os.path.join(_HOST, "librtb200_host.so")
_lib = None

NAMED = ("complexScene", "simpleScene", "cornellBoxScene", "cornellMixedScene", "randomSpheres")
# BASELINE.md configs -> scene spec, image size, samples per pixel (depth / fov come from the scene itself)
CONFIGS = {
    "C1": dict(spec="complexScene", width=800, height=800, spp=1, random_state=12345),
    "C2": dict(spec="meshRoom:660:1", width=1920, height=1080, spp=64, random_state=1),
    "C3": dict(spec="sphereField:100000:2", width=1920, height=1080, spp=256, random_state=2),
    "C4": dict(spec="heightField:3162:1581:3:0", width=3840, height=2160, spp=256, random_state=3),
    "C5": dict(spec="heightField:708:708:4:80", width=1920, height=1080, spp=1024, random_state=4),
}


def build(force: bool = False) -> str:
    capi.build()
    srcs = []
    for root, _, files in os.walk(_HOST):
        srcs += [os.path.join(root, f) for f in files if f.endswith((".cpp", ".hpp")) or f == "Makefile"]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HOST, "-j8", "-s"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        capi.lib()   # librtb200.so first: the host library links against it
        L = C.CDLL(_SO)
        L.rtbh_last_error.restype = C.c_char_p
        L.rtbh_scene_create.restype = C.c_void_p; L.rtbh_scene_create.argtypes = [C.c_char_p]
        L.rtbh_scene_destroy.restype = None; L.rtbh_scene_destroy.argtypes = [C.c_void_p]
        L.rtbh_scene_info.restype = None
        L.rtbh_scene_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_float)]
        L.rtbh_scene_copy.restype = None; L.rtbh_scene_copy.argtypes = [C.c_void_p] * 5
        L.rtbh_transform_mat4.restype = None; L.rtbh_transform_mat4.argtypes = [C.c_void_p] * 4
        L.rtbh_load_obj.restype = C.c_long; L.rtbh_load_obj.argtypes = [C.c_char_p, C.c_void_p, C.c_long]
        L.rtbh_set_model_dir.restype = None; L.rtbh_set_model_dir.argtypes = [C.c_char_p]
        L.rtbh_write_png.restype = C.c_int; L.rtbh_write_png.argtypes = [C.c_char_p, C.c_void_p, C.c_uint, C.c_uint]
        L.rtbh_set_model_dir(os.path.join(_HOST, "models").encode())
        _lib = L
    return _lib


def load_scene(spec: str) -> dict:
    """Build + flatten a scene.  Returns models / triangles / spheres / materials (reference record layouts, model
    space) and the scene's raysPerPixel, maxRaytraceDepth, verticalFOV."""
    L = lib()
    h = L.rtbh_scene_create(spec.encode())
    if not h:
        raise capi.RtbError(L.rtbh_last_error().decode())
    try:
        counts = (C.c_uint * 4)(); params = (C.c_uint * 2)(); fov = C.c_float()
        L.rtbh_scene_info(h, counts, params, C.byref(fov))
        models = np.zeros(counts[0], capi.MODEL); tris = np.zeros(counts[1], capi.TRIANGLE)
        sphs = np.zeros(counts[2], capi.SPHERE); mats = np.zeros(counts[3], capi.MATERIAL)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        L.rtbh_scene_copy(h, p(models), p(tris), p(sphs), p(mats))
    finally:
        L.rtbh_scene_destroy(h)
    return dict(models=models, triangles=tris, spheres=sphs, materials=mats, rays_per_pixel=int(params[0]),
                max_depth=int(params[1]), vfov=float(fov.value), spec=spec)


def transform_mat4(translation, scale, rotation) -> np.ndarray:
    t = np.asarray(translation, np.float32); s = np.asarray(scale, np.float32); r = np.asarray(rotation, np.float32)
    out = np.zeros(16, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib().rtbh_transform_mat4(p(t), p(s), p(r), p(out))
    return out


def write_png(path: str, rgba: np.ndarray):
    """RGBA8 [H, W, 4] -> PNG through the C++ hosts' own encoder (host/utils/Png.hpp)"""
    a = np.ascontiguousarray(rgba, np.uint8)
    if lib().rtbh_write_png(path.encode(), a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0]) != 0:
        raise capi.RtbError(lib().rtbh_last_error().decode())


def read_png_rgb(path: str) -> np.ndarray:
    """Decoder for the test suite (stdlib zlib): checks the signature, every chunk CRC, the zlib Adler-32, un-filters (type 0
    only, as the encoder writes) -> uint8 [H, W, 3]"""
    import struct
    import zlib
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n", "not a PNG"
    pos, idat, w = 8, b"", None
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert zlib.crc32(typ + body) == struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0], f"bad CRC in {typ}"
        if typ == b"IHDR":
            w, h, depth, ctype, comp, flt, inter = struct.unpack(">IIBBBBB", body)
            assert (depth, ctype, comp, flt, inter) == (8, 2, 0, 0, 0)
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
    assert np.all(raw[:, 0] == 0), "only filter type 0 is written"
    return raw[:, 1:].reshape(h, w, 3).copy()


def load_obj(path: str) -> np.ndarray:
    """Triangle positions [n,3,3] of an OBJ file through the host's loadModel (RTModel.cpp path)."""
    L = lib()
    n = L.rtbh_load_obj(path.encode(), None, 0)
    if n < 0:
        raise capi.RtbError(L.rtbh_last_error().decode())
    out = np.zeros((n, 3, 3), np.float32)
    L.rtbh_load_obj(path.encode(), out.ctypes.data_as(C.c_void_p), n)
    return out
