"""GPU: the LogisticMap demo program (Config::Programs::LogisticMap, logistic.comp) -- points and plotted image bit-exact
against the oracle, through the C-ABI and through the C++ host binary."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAIN = os.path.join(ROOT, "raytracergpu_mastersproject_b200", "host", "rtb200_main")


def test_logistic_steps_bit_exact(device):
    from raytracergpu_mastersproject_b200 import Buffer, capi
    rng = np.random.default_rng(5)
    n, W, H, steps = 100_000, 320, 200, 25
    pts = np.stack([rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32) * np.float32(4)], axis=1).astype(np.float32)
    pts[:4] = [[0.0, 3.9], [1.0, 4.0], [0.5, 4.0], [0.25, 0.0]]                  # edge points: y == H, x == W are discarded
    ref_pts = pts.copy(); ref_img = np.zeros((H, W, 4), np.uint8)
    bp = Buffer(device, 8, n); bp.write(pts)
    bi = Buffer(device, 4, W * H); bi.zero()
    col = (C.c_float * 4)(1.0, 0.5, 0.25, 1.0)
    for _ in range(steps):
        O.logistic_step(ref_pts, ref_img, (1.0, 0.5, 0.25, 1.0))
        capi.check(capi.lib().rtb_logistic_step(device.handle, bp._p, n, bi._p, W, H, col))
    device.wait_idle()
    assert np.array_equal(bp.read(np.float32, 2 * n).view(np.uint32), ref_pts.reshape(-1).view(np.uint32))
    assert np.array_equal(bi.read(np.uint8, W * H * 4).reshape(H, W, 4), ref_img)
    assert ref_img[..., 0].sum() > 0


def test_logistic_host_program(tmp_path):
    """rtb200_main --logistic = LogisticMapRenderer::LogisticMap{}.mainLoop(): 30 frames at 384x216; the host's own point
    set is read back implicitly: the plotted attractor must be non-trivial and confined to the r < 4 columns."""
    from raytracergpu_mastersproject_b200 import scenes
    scenes.build()
    W, H, frames = 384, 216, 30
    r = subprocess.run([MAIN, "--logistic", str(W), str(H), str(frames)], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    data = open(tmp_path / "frame.ppm", "rb").read().split(b"\n", 3)
    got = np.frombuffer(data[3], np.uint8).reshape(H, W, 3)
    lit = got[..., 0] == 255
    assert 0.02 < lit.mean() < 0.9
    # for r < 1 every orbit decays to x = 0 -> after 30 steps those columns are lit only near the bottom row band
    cols = lit[:, : W // 8]
    assert cols[: H // 2].mean() < cols[H // 2:].mean()
