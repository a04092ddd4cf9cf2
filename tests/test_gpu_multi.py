"""GPU, >= 2 devices: the library's own multi-GPU entry points (rtb_comm_*, rtb_gather_tiles, rtb_reduce_samples -- NCCL over NVLink,
bound at run time) driven by one process per GPU, the id handed over through a file.  The frame the ranks assemble must be the
1-GPU frame: bit for bit in tile mode (fp32 image and the fused RGBA8 resolve, both exchange paths), within 1e-5 (1 + |x|) with a
bit-exact alpha chain in sample-range mode; the sliced scene upload (rtb_comm_all_gather) must reproduce the arrays.
Skipped on a 1-GPU box (NCCL refuses two ranks on one device); run with `gpurun --gpus 2`."""
import os
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _worker(rank, world, tmp):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    from raytracergpu_mastersproject_b200 import Buffer, Device, Raytracer, capi, make_ubo, scenes
    from raytracergpu_mastersproject_b200.sharding import BandLayout, sample_range
    dev = Device(rank)
    idf = os.path.join(tmp, "nccl_id")
    if rank == 0:
        open(idf + ".tmp", "wb").write(Device.comm_unique_id()); os.rename(idf + ".tmp", idf)
    while not os.path.exists(idf):
        time.sleep(0.01)
    dev.comm_init(world, rank, open(idf, "rb").read())
    assert dev.comm_info() == (rank, world)
    sc = scenes.load_scene("meshRoom:70:5")                      # 9 806 primitives: the 4-ary nearest-first production walk
    W, H, spp, band = 200, 116, 6, 8                             # 15 bands over 2 ranks: ragged tail
    ubo = make_ubo(len(sc["triangles"]), len(sc["spheres"]), len(sc["materials"]), sc["max_depth"], 9, sc["vfov"])
    rt = Raytracer(dev, W, H, keep_reference_buffers=False)
    # sliced upload: every rank writes only its slice of the triangle array, the rest arrives over NVLink
    tb = np.ascontiguousarray(sc["triangles"]).view(np.uint8).reshape(-1)
    per = ((tb.size + world * 256 - 1) // (world * 256)) * 256
    rt.update_scene(sc["models"], sc["triangles"], sc["spheres"], sc["materials"])
    big = Buffer(dev, 1, per * world); big.zero()
    lo, hi = min(rank * per, tb.size), min((rank + 1) * per, tb.size)
    if hi > lo:
        capi.check(capi.lib().rtb_upload(dev.handle, C.c_void_p(big.ptr + lo), tb[lo:hi].ctypes.data_as(C.c_void_p), hi - lo))
    dev.comm_all_gather(big.ptr, per)
    assert np.array_equal(big.read(np.uint8, tb.size), tb), "sliced upload + all-gather does not reproduce the triangle array"
    rt.build_bvh(ubo)
    # ---- tile mode
    lay = BandLayout(H, world, band)
    rows = lay.local_rows
    rt.clear_image(rows)
    rt.raytrace(ubo, spp, rows=rows, band_rows=band, band_first=rank, band_step=world)
    f32 = Buffer(dev, 16, W * H); u8a = Buffer(dev, 4, W * H); u8b = Buffer(dev, 4, W * H)
    dev.gather_tiles(rt.image.ptr, W, H, band, f32.ptr, spp, u8a.ptr)      # RGBA32F exchange + fused resolve
    dev.gather_tiles(rt.image.ptr, W, H, band, None, spp, u8b.ptr)         # resolve first, RGBA8 exchange
    dev.wait_idle()
    tiled = f32.read(np.float32).reshape(H, W, 4); ra = u8a.read(np.uint8).reshape(H, W, 4); rb = u8b.read(np.uint8).reshape(H, W, 4)
    # ---- sample-range mode
    first, count = sample_range(spp, world, rank)
    rt.image = None
    rt.clear_image()
    rt.raytrace(ubo, count, sample_skip=first)
    u8c = Buffer(dev, 4, W * H)
    dev.reduce_samples(rt.image.ptr, W, H, 0, spp, u8c.ptr if rank == 0 else None)
    dev.wait_idle()
    summed = rt.read_image()
    # ---- the 1-GPU frame (every rank renders it: the results must also agree between devices)
    rt.clear_image(); rt.raytrace(ubo, spp); dev.wait_idle()
    ref = rt.read_image(); ref8 = rt.resolve_rgba8(spp)
    assert np.array_equal(tiled.view(np.uint32), ref.view(np.uint32)), "tile mode: assembled fp32 frame differs from the 1-GPU frame"
    assert np.array_equal(ra, ref8) and np.array_equal(rb, ref8), "tile mode: fused resolve differs"
    if rank == 0:
        assert np.array_equal(summed[..., 3].view(np.uint32), ref[..., 3].view(np.uint32)), "sample ranges: alpha chain differs"
        assert np.all(np.abs(summed[..., :3] - ref[..., :3]) <= 1e-5 * (1 + np.abs(ref[..., :3]))), "sample ranges: radiance out of tolerance"
        assert np.abs(u8c.read(np.uint8).reshape(H, W, 4).astype(int) - ref8.astype(int)).max() <= 1
    np.save(os.path.join(tmp, f"ref{rank}.npy"), ref)
    dev.close()


@pytest.mark.timeout(600)
def test_two_gpus_assemble_the_one_gpu_frame(tmp_path):
    if _device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "ref0.npy"), np.load(tmp_path / "ref1.npy")
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "two devices render different 1-GPU frames"


def test_cpp_host_on_two_gpus_matches_oracle(tmp_path):
    """rtb200_main --gpus 2: one host thread and one RaytracerBVHRenderer::Raytracer per GPU, bands assembled by rtb_gather_tiles;
    the frame rank 0 writes must be the oracle's."""
    if _device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import subprocess
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import oracle as O
    from raytracergpu_mastersproject_b200 import make_ubo, scenes
    scenes.build()
    main = os.path.join(ROOT, "raytracergpu_mastersproject_b200", "host", "rtb200_main")
    w, h, spec = 160, 116, "complexScene"
    r = subprocess.run([main, "--gpus", "2", spec, str(w), str(h)], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = scenes.read_png_rgb(str(tmp_path / "frame.png"))
    random_state = int.from_bytes(np.random.RandomState(12345).bytes(4), "little")
    sc = scenes.load_scene(spec)
    ubo = make_ubo(len(sc["triangles"]), len(sc["spheres"]), len(sc["materials"]), sc["max_depth"], random_state, sc["vfov"])
    b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    img = O.raytrace(ubo, w, h, b["tris"], b["sphs"], sc["materials"], b["nodes"], sc["rays_per_pixel"], want_hits=False, want_rng=False)["image"]
    assert np.array_equal(got, O.resolve_rgba8(img, sc["rays_per_pixel"])[..., :3])


def test_comm_entry_points_fail_loudly_without_a_communicator(device):
    from raytracergpu_mastersproject_b200 import Buffer, RtbError
    b = Buffer(device, 16, 64)
    with pytest.raises(RtbError, match="communicator"):
        device.gather_tiles(b.ptr, 8, 8, 8, b.ptr, 1, None)
    with pytest.raises(RtbError, match="communicator"):
        device.reduce_samples(b.ptr, 8, 8, 0, 1, None)
    with pytest.raises(RtbError, match="communicator"):
        device.comm_all_gather(b.ptr, 16)
    assert device.comm_info() == (0, 1)
