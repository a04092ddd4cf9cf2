"""GPU: the C++20 host mirror (host/rtb200_main = the reference's main.cpp shape: Raytracer{} + mainLoop()) renders the
default scene headless; its frame must equal the oracle's resolve of the same frame bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAIN = os.path.join(ROOT, "raytracergpu_mastersproject_b200", "host", "rtb200_main")


def _read_ppm(path):
    data = open(path, "rb").read()
    parts = data.split(b"\n", 3)
    assert parts[0] == b"P6"
    w, h = map(int, parts[1].split())
    return np.frombuffer(parts[3], np.uint8).reshape(h, w, 3)


@pytest.mark.parametrize("spec,w,h", [("complexScene", 160, 120), ("cornellBoxScene", 96, 96)])
def test_cpp_host_frame_matches_oracle(tmp_path, spec, w, h):
    from raytracergpu_mastersproject_b200 import make_ubo, scenes
    scenes.build()
    r = subprocess.run([MAIN, spec, str(w), str(h)], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Total BVH Build Time" in r.stdout
    got = _read_ppm(tmp_path / "frame.ppm")
    assert np.array_equal(scenes.read_png_rgb(str(tmp_path / "frame.png")), got), "frame.png differs from frame.ppm"
    # the host seeds std::mt19937 with Config::Headless::RandomState = 12345 and draws randomState = gen() once per frame
    random_state = int.from_bytes(np.random.RandomState(12345).bytes(4), "little")
    sc = scenes.load_scene(spec)
    ubo = make_ubo(len(sc["triangles"]), len(sc["spheres"]), len(sc["materials"]), sc["max_depth"], random_state, sc["vfov"])
    b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    img = O.raytrace(ubo, w, h, b["tris"], b["sphs"], sc["materials"], b["nodes"], sc["rays_per_pixel"], want_hits=False, want_rng=False)["image"]
    ref = O.resolve_rgba8(img, sc["rays_per_pixel"])[..., :3]
    assert np.array_equal(got, ref)


def test_cpp_non_bvh_host_frame_matches_oracle(tmp_path):
    """Config::Programs::Raytracer host (K1 + raytrace.comp) on the small simpleScene."""
    from raytracergpu_mastersproject_b200 import make_ubo, scenes
    scenes.build()
    w, h = 64, 64
    r = subprocess.run([MAIN, "--non-bvh", "simpleScene", str(w), str(h)], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = _read_ppm(tmp_path / "frame.ppm")
    random_state = int.from_bytes(np.random.RandomState(12345).bytes(4), "little")
    sc = scenes.load_scene("simpleScene")
    ubo = make_ubo(len(sc["triangles"]), len(sc["spheres"]), len(sc["materials"]), sc["max_depth"], random_state, sc["vfov"])
    tw, sw = O.model_to_world(sc["models"], sc["triangles"], sc["spheres"])
    img = O.raytrace(ubo, w, h, tw, sw, sc["materials"], None, sc["rays_per_pixel"], opt=O.make_options(linear_scan=True),
                     want_hits=False, want_rng=False)["image"]
    assert np.array_equal(got, O.resolve_rgba8(img, sc["rays_per_pixel"])[..., :3])


def test_rays_per_pixel_sweep_writes_runtimes_csv(tmp_path):
    """Config::RunRayPerPixelIncreasingDemo (RaytracerBVH.hpp:522-570; here the rtb200_rppdemo build): 100..200 rays per pixel in steps
    of 5, 4 frames each, mean frame time per step -> runtimes.csv, the reference's only benchmark artefact."""
    from raytracergpu_mastersproject_b200 import scenes
    scenes.build()
    demo = os.path.join(os.path.dirname(MAIN), "rtb200_rppdemo")
    r = subprocess.run([demo, "simpleScene", "96", "64"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-800:] + r.stderr
    rows = [l.split(",") for l in open(tmp_path / "runtimes.csv").read().splitlines() if l.strip()]
    assert [int(x[0]) for x in rows] == list(range(1, 22)), "one row per rays-per-pixel step 100, 105, ..., 200"
    us = [int(x[1]) for x in rows]
    assert all(u > 0 for u in us)
    assert r.stdout.count("RaysPerPixel: 100 ") >= 4 and "RaysPerPixel: 200 " in r.stdout and "RaysPerPixel: 210" not in r.stdout
    assert os.path.exists(tmp_path / "frame.png")


def test_cpp_host_reports_errors(tmp_path):
    r = subprocess.run([MAIN, "noSuchScene"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "unknown scene" in r.stderr
