"""CPU, world_size 2 over gloo: the N > 1 path's host logic -- band sharding, all-gather, re-assembly, sample-range
split -- with the CPU oracle standing in for the device renderer (the oracle renders exactly the rows / samples a rank
owns).  The device side of the same layouts is covered by tests/test_gpu_parity.py::test_tile_sharding_bit_identical."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scene_util as SU
    from oracle import oracle as O
    from raytracergpu_mastersproject_b200.sharding import BandLayout, assemble_gathered, sample_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H, spp, band = 48, 26, 4, 4                         # 26 rows / 4-row bands / 2 ranks -> ragged tail
    sc = SU.random_scene(31, n_tris=120, n_spheres=12)
    ubo = SU.make_ubo(sc, random_state=2024)
    b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])            # scene replicated: every rank builds
    opt = O.make_options(threads=1)
    lay = BandLayout(H, world, band)
    # --- tile mode: render the rows this rank owns, band by band, into its compact local buffer
    full = np.empty((H, W, 4), np.float32); O.lib().orc_clear_image(full.ctypes.data, W, H)
    for j0 in range(0, lay.local_rows, band):
        y0 = lay.global_row(rank, j0)
        if y0 < H:
            O.raytrace(ubo, W, H, b["tris"], b["sphs"], sc["materials"], b["nodes"], spp, rows=(y0, min(y0 + band, H)), image=full,
                       opt=opt, want_hits=False, want_rng=False)
    local = np.zeros((lay.local_rows, W, 4), np.float32)
    for j in range(lay.local_rows):
        y = lay.global_row(rank, j)
        if y < H:
            local[j] = full[y]
    gathered = [torch.zeros(lay.local_rows, W, 4) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(local))
    tiled = assemble_gathered(torch.stack(gathered), lay).numpy()
    # --- sample-range mode: this rank renders its sample range; partial sums are reduced
    first, count = sample_range(spp, world, rank)
    img = np.empty((H, W, 4), np.float32); O.lib().orc_clear_image(img.ctypes.data, W, H)
    if first:                                               # fast-forward the alpha chain by rendering nothing: emulate sampleSkip
        for y in range(H):
            for x in range(W):
                a = 1.0
                for _ in range(first):
                    a, _s = O.pcg_float((O.seed_base(x, y, 2024) + O.alpha_to_u32(a)) & 0xFFFFFFFF)
                img[y, x, 3] = a
    part = O.raytrace(ubo, W, H, b["tris"], b["sphs"], sc["materials"], b["nodes"], count, image=img, opt=opt,
                      want_hits=False, want_rng=False)["image"]
    rgb = torch.from_numpy(part[..., :3].copy())
    dist.reduce(rgb, dst=0, op=dist.ReduceOp.SUM)
    alpha_last = torch.from_numpy(part[..., 3].copy())
    if rank == world - 1 and world > 1:
        dist.send(alpha_last, dst=0)
    if rank == 0:
        if world > 1:
            dist.recv(alpha_last, src=world - 1)
        ref = O.raytrace(ubo, W, H, b["tris"], b["sphs"], sc["materials"], b["nodes"], spp, opt=opt, want_hits=False, want_rng=False)["image"]
        np.savez(out, tiled=tiled, ref=ref, rgb=rgb.numpy(), alpha=alpha_last.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_tile_and_sample_range(tmp_path):
    out = str(tmp_path / "r.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    d = np.load(out)
    # tile mode is bit-identical to the single-rank frame
    assert np.array_equal(d["tiled"].view(np.uint32), d["ref"].view(np.uint32))
    # sample-range mode: alpha chain bit-exact; radiance within the fp32 re-association tolerance 1e-5 * (1 + |x|)
    assert np.array_equal(d["alpha"].view(np.uint32), d["ref"][..., 3].view(np.uint32))
    assert np.all(np.abs(d["rgb"] - d["ref"][..., :3]) <= 1e-5 * (1 + np.abs(d["ref"][..., :3])))


def test_band_layout_covers_every_row_once():
    from raytracergpu_mastersproject_b200.sharding import BandLayout, sample_range, single_gpu_layout
    for H, world, band in [(1080, 8, 8), (2160, 8, 8), (26, 2, 4), (50, 4, 4), (7, 8, 8), (800, 3, 8)]:
        lay = BandLayout(H, world, band)
        rows = sorted(y for r in range(world) for y in lay.owned_rows(r))
        assert rows == list(range(H))
        assert lay.local_rows % band == 0
    assert single_gpu_layout(33).owned_rows(0) == list(range(33))
    for spp, world in [(1024, 8), (64, 8), (7, 4), (3, 8)]:
        parts = [sample_range(spp, world, r) for r in range(world)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == spp
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
