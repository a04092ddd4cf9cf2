"""CPU-only: the C-ABI library loads and exports every symbol include/rtb200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rtb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rtb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from raytracergpu_mastersproject_b200 import capi
    capi.build()
    L = ctypes.CDLL(capi.library_path())
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in rtb200.h but not exported"
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS out of sync with rtb200.h"


def test_version_and_loud_failure_without_gpu():
    import torch
    from raytracergpu_mastersproject_b200 import Device, RtbError, capi
    assert capi.lib().rtb_version() == 200
    if not torch.cuda.is_available():
        try:
            Device(0)
        except RtbError as e:
            assert "no CUDA device" in str(e) or "cuda" in str(e).lower()
        else:
            raise AssertionError("Device(0) must fail loudly without a GPU (no CPU fallback)")


def test_record_sizes_match_reference_layouts():
    from raytracergpu_mastersproject_b200 import capi
    sizes = [capi.MODEL.itemsize, capi.TRIANGLE.itemsize, capi.SPHERE.itemsize, capi.MATERIAL.itemsize,
             capi.BVH_NODE.itemsize, capi.MORTON_PRIMITIVE.itemsize, capi.CONSTRUCTION_INFO.itemsize,
             capi.ENCLOSING_BOX.itemsize, capi.UBO.itemsize]
    assert sizes == [64, 64, 32, 32, 40, 12, 8, 32, 80]
    assert ctypes.sizeof(capi.TraceArgs) == 80


def test_concurrent_builds_are_serialised(tmp_path):
    """The ranks of a torchrun launch all call capi.lib() -> build(): with a stale library they must not race (one builds under a
    file lock and renames the finished library into place, the others wait and load it)."""
    import subprocess
    import sys
    import time
    from raytracergpu_mastersproject_b200 import capi
    capi.build()
    src = os.path.join(os.path.dirname(capi.library_path()), "csrc", "probe.cu")
    time.sleep(0.05); os.utime(src)                              # the smallest translation unit: stale by mtime only
    code = ("from raytracergpu_mastersproject_b200 import capi; import ctypes; capi.build(); L = ctypes.CDLL(capi.library_path()); "
            "assert L.rtb_version() >= 200; assert not capi._stale()")
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    procs = [subprocess.Popen([sys.executable, "-c", code], env=env, stderr=subprocess.PIPE) for _ in range(4)]
    errs = [(p.wait(), p.stderr.read().decode()[-400:]) for p in procs]
    assert all(rc == 0 for rc, _ in errs), errs
