"""Builds the SIMT-emulated library (TEST INFRASTRUCTURE ONLY, see tests/emu/cuda_runtime.h): the kernel sources of
raytracergpu_mastersproject_b200/csrc/*.cu, unmodified except for the `kernel<<<grid, block, smem, stream>>>(args)` launch syntax
(rewritten to emu::launch), compiled with g++ against the emulation header.  Output: <out_dir>/librtb200_emu.so with the C-ABI of
include/rtb200.h.  The emulated device has one multiprocessor, so the streaming A/B kernel's cooperative launch is a single block.

    python tests/emu/build_emu.py [out_dir]        (default: $TMPDIR/rtb200_emu -- outside the repository on purpose)
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "raytracergpu_mastersproject_b200", "csrc")
UNITS = ["capi.cu", "comm.cu", "probe.cu", "bvh_build.cu", "traversal_tree.cu", "radix_sort.cu", "trace.cu", "trace_wave.cu"]


def _match_back(s, end):
    """start index of the kernel expression that ends at `end` (an identifier with optional namespace and template arguments)"""
    i = end
    while i > 0 and s[i - 1].isspace():
        i -= 1
    if s[i - 1] == ">":                       # template arguments: skip back to the matching '<'
        depth = 0
        while i > 0:
            i -= 1
            if s[i] == ">":
                depth += 1
            elif s[i] == "<":
                depth -= 1
                if depth == 0:
                    break
    while i > 0 and (s[i - 1].isalnum() or s[i - 1] in "_:"):
        i -= 1
    return i


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


COOP = re.compile(r"cudaLaunchCooperativeKernel\(\(const void\*\)(.+?), (dim3\(.+?\)), (dim3\(.+?\)), args, 0, st\);")


def rewrite_launches(src):
    # the cooperative launch of the streaming kernel (its single argument is the TraceParams `p` the args array points at)
    src = COOP.sub(lambda m: f"emu::launch({m.group(2)}, {m.group(3)}, [=]() {{ ({m.group(1)})(p); }});", src)
    out, pos = "", 0
    while True:
        a = src.find("<<<", pos)
        if a < 0:
            return out + src[pos:]
        b = src.index(">>>", a)
        k0 = _match_back(src, a)
        kernel = src[k0:a].strip()
        cfg = _split_top(src[a + 3:b])
        j = b + 3
        while src[j].isspace():
            j += 1
        assert src[j] == "(", src[a - 40:b + 40]
        depth, e = 0, j
        while True:
            if src[e] == "(":
                depth += 1
            elif src[e] == ")":
                depth -= 1
                if depth == 0:
                    break
            e += 1
        args = src[j + 1:e]
        out += src[pos:k0] + f"emu::launch(dim3({cfg[0]}), dim3({cfg[1]}), [=]() {{ ({kernel})({args}); }})"
        pos = e + 1




def build(out_dir=None, defines=()):
    out_dir = out_dir or os.path.join(tempfile.gettempdir(), "rtb200_emu")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "librtb200_emu.so")
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in os.listdir(HERE)] + [os.path.join(ROOT, "include", "rtb200.h")]
    if os.path.exists(so) and all(os.path.getmtime(s) <= os.path.getmtime(so) for s in srcs):
        return so
    flags = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-frounding-math", "-fno-strict-aliasing", "-w", "-Wno-psabi", "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0",
             "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include")] + ["-D" + d for d in defines]
    objs, procs = [], []
    for u in UNITS:
        cpp = os.path.join(out_dir, u.replace(".cu", "_emu.cpp"))
        text = open(os.path.join(CSRC, u)).read()
        open(cpp, "w").write('#include <cuda_runtime.h>\n#line 1 "%s"\n' % os.path.join(CSRC, u) + rewrite_launches(text))
        objs.append(cpp.replace(".cpp", ".o"))
        procs.append(subprocess.Popen(flags + ["-c", cpp, "-o", objs[-1]]))
    for src in (os.path.join(HERE, "emu_runtime.cpp"),):
        objs.append(os.path.join(out_dir, os.path.basename(src).replace(".cpp", ".o")))
        procs.append(subprocess.Popen(flags + ["-c", src, "-o", objs[-1]]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("emu build failed")
    subprocess.run(["g++", "-shared", "-o", so] + objs + ["-ldl"], check=True)
    return so


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else None))
