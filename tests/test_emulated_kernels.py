"""CPU tier: the CUDA kernel SOURCES, executed under SIMT emulation, against the oracle and the reference binaries' vectors.

tests/emu/ compiles raytracergpu_mastersproject_b200/csrc/*.cu unmodified (only the `<<<...>>>` launch syntax is rewritten) with g++
against an emulation of the CUDA runtime in which every thread is a fiber and the fibers of a block meet at the warp collectives
(__ballot_sync, __shfl_sync, __match_any_sync, ...) and at __syncthreads -- so the kernels' warp-level control flow (phase
votes, ballot-ranked queue appends, the one-ray-per-warp tail kernel, the radix sort's peer ranking) runs as written.  This test
then runs a selection of the `-m gpu` parity tests THEMSELVES over that library (tests/conftest.py points the ctypes binding at it) in a
subprocess.  It is a check of the kernel code in the tier that has no GPU and a development aid; it is not a product path: the
emulated library is built into a scratch directory outside the repository, only this test loads it, and the device it reports
is called "SIMT-EMU".  The GPU tier runs the same tests (and all the others) on the real device.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import pytest

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="the emulated build needs g++")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

# (test files, -k expression): nearly the whole `-m gpu` suite; left out are only the tests at sizes the emulation is too slow for
# (full-size configs, multi-pass, mid-size meshes, 5-sample frames of every seed, the 800x800 / 1920x1080 pixel-sampled fixtures)
SELECTION = [
    ("tests/test_gpu_parity.py", "not multi_pass and not mid_size and not (test_frame_bit_exact and 5-seed)"),
    ("tests/test_spirv_golden.py tests/test_golden.py tests/test_gpu_logistic.py",
     "test_cuda_matches_reference_binaries or test_cuda_logistic or test_logistic_steps or (test_cuda_reproduces_golden and not complexScene)"),
    # the C++20 host binary (reference main.cpp shape) over the emulated library: its PPM frame == the oracle's resolved frame
    ("tests/test_gpu_host_main.py", "complexScene or non_bvh or reports_errors"),
]


@pytest.fixture(scope="module")
def emulated_library():
    sys.path.insert(0, os.path.join(HERE, "emu"))
    import build_emu
    return build_emu.build(os.path.join(tempfile.gettempdir(), "rtb200_emu"))


def test_emulated_library_is_not_the_product(emulated_library):
    pkg = os.path.join(ROOT, "raytracergpu_mastersproject_b200")
    assert not os.path.abspath(emulated_library).startswith(os.path.abspath(ROOT) + os.sep), "the emulated build must stay outside the repository"
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")) or f == "Makefile":
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "librtb200_emu" not in text and "build_emu" not in text, f"{f} refers to the emulated build"


def test_tail_kernel_exact_fallback_over_the_emulated_kernels():
    """trace_tail_kernel sends a ray to hit_bvh() -- the reference's own walk on the exact records -- when its pending set
    overflows or a NaN hit turns up; neither happens on ordinary scenes.  A test build (-DRTB_TAIL_TEST_FALLBACK) forces every
    third ray down that path; with the hand-over thresholds forced low as well, the frames must still be the oracle's."""
    sys.path.insert(0, os.path.join(HERE, "emu"))
    import build_emu
    lib = build_emu.build(os.path.join(tempfile.gettempdir(), "rtb200_emu_fallback"), defines=("RTB_TAIL_TEST_FALLBACK",))
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-x", "-q", "-m", "gpu", "-k", "test_tail_handover_forced and 8-3",
                        "-p", "no:cacheprovider"], cwd=ROOT, env=dict(os.environ, RTB_EMU_LIB=lib, RTB_TEST_EMULATION="1"), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and " passed" in r.stdout, (r.stdout + r.stderr)[-1500:]


def test_package_has_no_hook_for_the_emulated_library(emulated_library):
    """Outside pytest's conftest the package loads only its own in-tree librtb200.so, whatever the environment says: on this
    GPU-less tier that ends in 'no CUDA device', never in an emulated frame."""
    code = ("from raytracergpu_mastersproject_b200 import Device, capi; assert capi.library_path().endswith('raytracergpu_mastersproject_b200/librtb200.so');"
            "Device(0)")
    env = dict(os.environ, RTB_LIB=emulated_library, RTB_EMU_LIB=emulated_library, RTB_TEST_EMULATION="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    import torch
    if not torch.cuda.is_available():
        assert r.returncode != 0 and "no CUDA device" in r.stderr, r.stderr[-800:]
    for f in ("capi.py", "renderer.py", "scenes.py", "sharding.py", "__init__.py"):
        text = open(os.path.join(ROOT, "raytracergpu_mastersproject_b200", f)).read()
        assert "RTB_LIB" not in text and "RTB_EMU_LIB" not in text and "EMULATION" not in text, f


@pytest.mark.parametrize("path,expr", SELECTION, ids=["parity", "fixtures", "cpp_host"])
def test_gpu_parity_tests_pass_over_the_emulated_kernels(emulated_library, path, expr):
    # rtb200_main / librtb200_host.so name librtb200.so in their NEEDED entries: a directory with that name pointing at the emulated
    # build, searched before their RUNPATH, binds them to it for this subprocess only
    ld = os.path.join(os.path.dirname(emulated_library), "ld")
    os.makedirs(ld, exist_ok=True)
    link = os.path.join(ld, "librtb200.so")
    if not os.path.islink(link):
        os.symlink(emulated_library, link)
    env = dict(os.environ, RTB_EMU_LIB=emulated_library, RTB_TEST_EMULATION="1", LD_LIBRARY_PATH=ld + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([sys.executable, "-m", "pytest", *path.split(), "-x", "-q", "-m", "gpu", "-k", expr, "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "SIMT-EMU" not in r.stderr, tail
