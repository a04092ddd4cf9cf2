"""Python host-side mirror of the reference's renderer host for the BVH program, over the C-ABI.

Names follow the reference: Device / Buffer (VulkanWrapper/Device.hpp, Buffer.hpp), and `Raytracer` whose
`doIteration` runs S1 (BVH build) then S2 (trace) like RaytracerBVHRenderer::Raytracer::doIteration
(RaytracerBVH.hpp:206-496).  The C++20 mirror of the same API lives in host/ ; this module exists so tests and
bench.py can drive the identical C-ABI entry points from Python.  No compute happens in Python.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import UBO, TraceArgs, check


class Device:
    """A CUDA device + one ordered compute queue (Device.hpp:27-113; Device::computeQueue)."""

    def __init__(self, index: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        self._lib = capi.lib()
        check(self._lib.rtb_ctx_create(index, C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.index = index

    @property
    def handle(self):
        return self._h

    def name(self) -> str:
        buf = C.create_string_buffer(256)
        check(self._lib.rtb_device_name(self._h, buf, 256))
        return buf.value.decode()

    def sm_count(self) -> int:
        n = C.c_int()
        check(self._lib.rtb_sm_count(self._h, C.byref(n)))
        return n.value

    def wait_idle(self):  # vkWaitForFences / vkDeviceWaitIdle
        check(self._lib.rtb_sync(self._h))

    def launch_count(self) -> int:
        n = C.c_uint64()
        check(self._lib.rtb_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def timer_start(self):
        check(self._lib.rtb_timer_start(self._h))

    def timer_stop_ms(self) -> float:
        ms = C.c_float()
        check(self._lib.rtb_timer_stop_ms(self._h, C.byref(ms)))
        return float(ms.value)

    # -- multi-GPU: one Device per rank; the exchange at the end of a frame (include/rtb200.h "multi-GPU")
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(capi.lib().rtb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes):
        check(self._lib.rtb_comm_init_rank(self._h, n_ranks, rank, C.c_char_p(unique_id)))

    def comm_info(self) -> tuple[int, int]:
        r, n = C.c_int(), C.c_int()
        check(self._lib.rtb_comm_info(self._h, C.byref(r), C.byref(n)))
        return r.value, n.value

    def comm_all_gather(self, ptr: int, bytes_per_rank: int):
        check(self._lib.rtb_comm_all_gather(self._h, C.c_void_p(ptr), bytes_per_rank))

    def gather_tiles(self, local_image_ptr: int, width: int, height: int, band_rows: int, frame_f32_ptr: int | None,
                     rays_per_pixel: int, frame_rgba8_ptr: int | None):
        check(self._lib.rtb_gather_tiles(self._h, C.c_void_p(local_image_ptr), width, height, band_rows,
                                         C.c_void_p(frame_f32_ptr) if frame_f32_ptr else None, rays_per_pixel,
                                         C.c_void_p(frame_rgba8_ptr) if frame_rgba8_ptr else None))

    def reduce_samples(self, image_ptr: int, width: int, height: int, root: int, rays_per_pixel: int, frame_rgba8_ptr: int | None):
        check(self._lib.rtb_reduce_samples(self._h, C.c_void_p(image_ptr), width, height, root, rays_per_pixel,
                                           C.c_void_p(frame_rgba8_ptr) if frame_rgba8_ptr else None))

    def close(self):
        if self._h:
            self._lib.rtb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class Buffer:
    """Device-local storage buffer (Buffer.hpp:23-57): instanceSize x instanceCount bytes, freed in the dtor."""

    def __init__(self, device: Device, instance_size: int, instance_count: int):
        self.device = device
        self.instance_size = int(instance_size)
        self.instance_count = int(instance_count)
        self._p = C.c_void_p()
        check(capi.lib().rtb_alloc(device.handle, self.size, C.byref(self._p)))

    @property
    def size(self) -> int:
        return self.instance_size * self.instance_count

    @property
    def ptr(self) -> int:
        return self._p.value or 0

    def write(self, array: np.ndarray):
        """staging map + writeToBuffer + Device::copyBuffer (RaytraceScene.hpp:139-178)"""
        a = np.ascontiguousarray(array)
        if a.nbytes > self.size:
            raise capi.RtbError("Buffer.write: source larger than buffer")
        check(capi.lib().rtb_upload(self.device.handle, self._p, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def read(self, dtype, count: int | None = None) -> np.ndarray:
        """DEBUGgetDeployedBufferAs<T> (RaytracerBVH.hpp:575-617)"""
        dtype = np.dtype(dtype)
        n = self.size // dtype.itemsize if count is None else count
        out = np.empty(n, dtype)
        check(capi.lib().rtb_download(self.device.handle, out.ctypes.data_as(C.c_void_p), self._p, out.nbytes))
        return out

    def zero(self):
        check(capi.lib().rtb_memset(self.device.handle, self._p, 0, self.size))

    def free(self):
        if self._p:
            capi.lib().rtb_free(self.device.handle, self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:  # noqa: BLE001
            pass


def make_ubo(num_triangles, num_spheres, num_materials, max_depth, random_state, vfov=40.0,
             cam_pos=(275.0, 275.0, -800.0), look_at=(275.0, 275.0, 0.0), up=(0.0, 1.0, 0.0), num_lights=20):
    """RaytracingUniformBufferObject as doIteration fills it (RaytracerBVH.hpp:340-352): hard-coded camera (D10)."""
    u = np.zeros(1, UBO)
    u["camPos"][0, :3] = cam_pos
    u["camLookAt"][0, :3] = look_at
    u["camUpDir"][0, :3] = up
    u["verticalFOV"] = vfov
    u["numTriangles"] = num_triangles
    u["numSpheres"] = num_spheres
    u["numMaterials"] = num_materials
    u["numLights"] = num_lights
    u["maxRayTraceDepth"] = max_depth
    u["randomState"] = random_state & 0xFFFFFFFF
    return u


class Raytracer:
    """RaytracerBVHRenderer::Raytracer for one GPU, headless (RaytracerBVH.hpp:51-573).

    Holds the scene SSBOs and the build buffers created in createScene (RaytracerBVH.cpp:479-550) and runs the
    S1 / S2 submissions of doIteration through the C-ABI.
    """

    def __init__(self, device: Device, width: int = 800, height: int = 800, keep_reference_buffers: bool = True):
        self.device = device
        self.width, self.height = int(width), int(height)   # window{800,800}: RaytracerBVH.cpp:8
        self.keep = keep_reference_buffers
        self._lib = capi.lib()
        self.models = self.triangles = self.spheres = self.materials = None
        self.enclosing = self.morton1 = self.morton2 = self.nodes = self.cinfo = None
        self.image = None
        self.counters = Buffer(device, 8, 6)
        self.T = self.S = self.M = 0
        self._host = None

    # -- RaytraceScene::prepForRender / updateScene: (re-)upload the model-space arrays (RaytraceScene.cpp:78-113)
    def update_scene(self, models, triangles, spheres, materials):
        models = np.ascontiguousarray(models); triangles = np.ascontiguousarray(triangles)
        spheres = np.ascontiguousarray(spheres); materials = np.ascontiguousarray(materials)
        T, S, M = len(triangles), len(spheres), len(materials)
        d = self.device
        if self.models is None or (T, S, M, len(models)) != (self.T, self.S, self.M, self.models.instance_count):
            self.models = Buffer(d, 64, max(len(models), 1))
            self.triangles = Buffer(d, 64, max(T, 1))
            self.spheres = Buffer(d, 32, max(S, 1))
            self.materials = Buffer(d, 32, max(M, 1))
            n = T + S
            if self.keep:
                self.enclosing = Buffer(d, 32, 1)
                self.morton1 = Buffer(d, 12, max(n, 1))
                self.morton2 = Buffer(d, 12, max(n, 1))
                self.nodes = Buffer(d, 40, max(2 * n - 1, 1))
                self.cinfo = Buffer(d, 8, max(2 * n - 1, 1))
            self.T, self.S, self.M = T, S, M
        self.models.write(models); self.triangles.write(triangles)
        self.spheres.write(spheres); self.materials.write(materials)

    def _p(self, b):
        return b._p if b is not None else None

    # -- S1: recordComputeS1CommandBuffer + submit (RaytracerBVH.cpp:734-997)
    def build_bvh(self, ubo, flags: int = 0):
        u = np.ascontiguousarray(ubo)
        check(self._lib.rtb_build_bvh(self.device.handle, u.ctypes.data_as(C.c_void_p), self._p(self.models),
                                      self._p(self.triangles), self._p(self.spheres), self._p(self.materials),
                                      self._p(self.enclosing), self._p(self.morton1), self._p(self.morton2),
                                      self._p(self.nodes), self._p(self.cinfo), flags))

    # -- the non-BVH program's frame preparation (Raytracer.cpp:394-538): K1 only, then bind without nodes
    def prepare_linear(self, ubo):
        u = np.ascontiguousarray(ubo)
        up = u.ctypes.data_as(C.c_void_p)
        check(self._lib.rtb_model_to_world(self.device.handle, up, self._p(self.models), self._p(self.triangles), self._p(self.spheres)))
        check(self._lib.rtb_bind_trace_buffers(self.device.handle, up, self._p(self.triangles), self._p(self.spheres),
                                               self._p(self.materials), None))

    def ensure_image(self, rows: int | None = None):
        rows = self.height if rows is None else rows
        if self.image is None or self.image.instance_count != rows * self.width:
            self.image = Buffer(self.device, 16, rows * self.width)
        return self.image

    def clear_image(self, rows: int | None = None):
        rows = self.height if rows is None else rows
        img = self.ensure_image(rows)
        check(self._lib.rtb_clear_image(self.device.handle, img._p, self.width, rows))

    # -- S2: recordComputeS2CommandBuffer + submit (RaytracerBVH.cpp:998-1050)
    def raytrace(self, ubo, sample_count: int, sample_skip: int = 0, flags: int = 0, rows: int | None = None,
                 band_rows: int | None = None, band_first: int = 0, band_step: int = 1,
                 hit_prim: Buffer | None = None, hit_t: Buffer | None = None, rng_out: Buffer | None = None,
                 image_ptr: int | None = None, walk_counters: Buffer | None = None):
        rows = self.height if rows is None else rows
        a = TraceArgs()
        a.imageWidth, a.imageHeight, a.localRows = self.width, self.height, rows
        a.bandRows = self.height if band_rows is None else band_rows
        a.bandFirst, a.bandStep = band_first, band_step
        a.sampleSkip, a.sampleCount, a.flags = sample_skip, sample_count, flags
        a.hitPrim = hit_prim.ptr if hit_prim else None
        a.hitT = hit_t.ptr if hit_t else None
        a.rngOut = rng_out.ptr if rng_out else None
        a.counters = self.counters.ptr if (flags & capi.TRACE_COUNT) else None
        a.walkCounters = walk_counters.ptr if walk_counters else None
        u = np.ascontiguousarray(ubo)
        img = C.c_void_p(image_ptr) if image_ptr else self.ensure_image(rows)._p
        check(self._lib.rtb_raytrace(self.device.handle, u.ctypes.data_as(C.c_void_p), img, C.byref(a)))

    def read_counters(self, reset: bool = True) -> dict:
        v = self.counters.read(np.uint64, 6)
        if reset:
            self.counters.zero()
        return {k: int(x) for k, x in zip(capi.COUNTER_FIELDS, v)}

    def read_image(self, rows: int | None = None) -> np.ndarray:
        rows = self.height if rows is None else rows
        return self.image.read(np.float32, rows * self.width * 4).reshape(rows, self.width, 4)

    def resolve_rgba8(self, rays_per_pixel: int, rows: int | None = None) -> np.ndarray:
        rows = self.height if rows is None else rows
        out = Buffer(self.device, 4, rows * self.width)
        check(self._lib.rtb_resolve_rgba8(self.device.handle, self.image._p, self.width, rows, rays_per_pixel, out._p))
        return out.read(np.uint8, rows * self.width * 4).reshape(rows, self.width, 4)

    # -- doIteration (RaytracerBVH.hpp:206-496): updateScene -> UBO -> S1 -> wait -> S2 -> wait
    def do_iteration(self, scene: dict, ubo, rays_per_pixel: int, flags: int = 0):
        self.update_scene(scene["models"], scene["triangles"], scene["spheres"], scene["materials"])
        self.build_bvh(ubo, flags)
        self.clear_image()
        self.device.wait_idle()
        self.raytrace(ubo, rays_per_pixel, flags=flags)
        self.device.wait_idle()
        return self.read_image()
