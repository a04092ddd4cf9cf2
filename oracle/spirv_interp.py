"""A small SPIR-V interpreter -- TEST INFRASTRUCTURE ONLY (golden-vector generation, this container only).

Purpose: the reference ships its compute shaders pre-compiled (`shaders/compiled/*.comp.spv`, built by `compile.bat` with
glslc, unoptimised, SPIR-V 1.0 / 1.3).  No Vulkan driver exists in this image, so those binaries cannot run on a device --
but they are plain SPIR-V, and this module executes them instruction by instruction on the CPU.  `tests/golden/make_spirv_golden.py`
uses it to run THE REFERENCE'S OWN BINARIES on small seeded inputs and commits the outputs as golden vectors; the C oracle
(`oracle/rt_oracle.c`) and the CUDA path are then checked against them.  Everything the binary defines is therefore taken from the
reference itself: control flow, evaluation order of every expression, constants, struct layouts (Offset / ArrayStride
decorations), RNG, traversal order, tie-breaks.

What a SPIR-V binary does NOT define is the arithmetic of the driver's built-ins.  Those follow the pins of SURVEY.md
Appendix B (the same ones the oracle uses) and are listed here once:
  * + - * / sqrt and int<->float conversions: IEEE-754 binary32, round to nearest even, no contraction (computed in binary64
    and rounded once -- exact for these operations); denormals kept;
  * OpDot: ((a0*b0 + a1*b1) + a2*b2) + ...; OpMatrixTimesVector: ((c0*x + c1*y) + c2*z) + c3*w; OpVectorTimesScalar per lane;
  * GLSL.std.450 Cross = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x); Normalize = v / sqrt(dot(v, v));
    FMin(x, y) = y < x ? y : x; FMax(x, y) = x < y ? y : x; FClamp = FMin(FMax(x, lo), hi); Radians = x * fl(pi / 180);
  * Sin / Cos: caller-supplied (`sincos`), the golden generator passes the oracle's pinned `orc_pin_sincos`; Tan: libm tanf;
    Acos / Atan2: libm (only reached by dead code in the reference);
  * OpConvertFToU / FToS saturate, NaN -> 0 (pin U1 / U5: NVIDIA F2I behaviour);
  * variables without initialiser read as zero (pin U4); an image declared `rgba8` but created RGBA32F is fp32 (pin U3).

Invocations run one after the other (workgroup by workgroup, local index ascending).  Shaders with barriers / subgroup
operations run their workgroup in lock-step (every invocation is a generator that yields at those instructions); the
subgroup size is 32, as the reference assumes (RadixSortSimple.comp:10).

Nothing under raytracergpu_mastersproject_b200/ imports this module, and neither does anything that runs on the GPU box:
the reference tree (`/root/reference`) only exists in the build container.
"""
from __future__ import annotations

import ctypes
import math
import struct

_cf = ctypes.c_float()
_libm = ctypes.CDLL("libm.so.6")
for _n in ("tanf", "acosf", "sinf", "cosf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
_libm.atan2f.restype = ctypes.c_float
_libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]

M32 = 0xFFFFFFFF
INF = float("inf")
NAN = float("nan")


def f32(x):
    """round a binary64 value to binary32 (RNE), keep it as a Python float"""
    _cf.value = x
    return _cf.value


def fdiv(a, b):
    try:
        return f32(a / b)
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return NAN
        return -INF if (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0) else INF


def fsqrt(a):
    if a != a or a < 0.0:
        return NAN
    if a == INF:
        return INF
    return f32(math.sqrt(a))


def s32(u):
    return u - 0x100000000 if u & 0x80000000 else u


def f2u(f):
    if not (f > 0.0):
        return 0
    if f >= 4294967296.0:
        return M32
    return int(f)


def f2s(f):
    if f != f:
        return 0
    if f >= 2147483648.0:
        return 0x7FFFFFFF
    if f <= -2147483648.0:
        return 0x80000000
    return int(f) & M32


def bits_f(u):
    return struct.unpack("<f", struct.pack("<I", u & M32))[0]


def f_bits(f):
    return struct.unpack("<I", struct.pack("<f", f))[0]


def cp(v):
    return [cp(x) for x in v] if type(v) is list else v


def assign(dst, k, v):
    """store v into dst[k]; composites are copied element-wise into the existing storage (pointers stay valid)"""
    if type(v) is list:
        cur = dst[k]
        if type(cur) is list and len(cur) == len(v):
            for i, x in enumerate(v):
                assign(cur, i, x)
        else:
            dst[k] = cp(v)
    else:
        dst[k] = v


# decorations / builtins / storage classes used
DEC_BLOCK, DEC_BUFFER_BLOCK, DEC_ARRAY_STRIDE, DEC_MATRIX_STRIDE, DEC_BUILTIN, DEC_BINDING, DEC_SET, DEC_OFFSET = 2, 3, 6, 7, 11, 33, 34, 35
BI_NUM_WG, BI_WG_SIZE, BI_WG_ID, BI_LOCAL_ID, BI_GLOBAL_ID, BI_LOCAL_INDEX = 24, 25, 26, 27, 28, 29
BI_SG_SIZE, BI_NUM_SG, BI_SG_ID, BI_SG_LOCAL_ID = 36, 38, 40, 41
SC_UNIFORM_CONSTANT, SC_INPUT, SC_UNIFORM, SC_OUTPUT, SC_WORKGROUP, SC_PRIVATE, SC_FUNCTION, SC_STORAGE_BUFFER = 0, 1, 2, 3, 4, 6, 7, 12


class Image:
    """a storage image: float32 texels [H][W][4] kept as nested lists"""

    def __init__(self, array_hw4):
        self.h = len(array_hw4)
        self.w = len(array_hw4[0])
        self.px = [[[f32(float(c)) for c in t] for t in row] for row in array_hw4]


class Function:
    __slots__ = ("id", "rtype", "params", "blocks", "entry", "vars")

    def __init__(self, fid, rtype):
        self.id, self.rtype, self.params, self.blocks, self.entry, self.vars = fid, rtype, [], {}, None, []


class Module:
    def __init__(self, path, sincos=None):
        data = open(path, "rb").read()
        w = struct.unpack("<%dI" % (len(data) // 4), data)
        if w[0] != 0x07230203:
            raise ValueError("not a SPIR-V module: " + path)
        self.version = w[1]
        self.bound = w[3]
        self.names, self.member_names = {}, {}
        self.decor, self.mdecor = {}, {}
        self.types, self.consts, self.globals = {}, {}, {}
        self.functions = {}
        self.entry = None
        self.local_size = (1, 1, 1)
        self.glsl_ext = None
        self.sincos = sincos or (lambda x: (f32(_libm.sinf(x)), f32(_libm.cosf(x))))
        self.n_executed = 0
        self.calls = {}                 # function id -> number of OpFunctionCall executions (work counters of the reference binary)
        cur_fn, cur_block = None, None
        i = 5
        while i < len(w):
            wc, op = w[i] >> 16, w[i] & 0xFFFF
            a = w[i + 1:i + wc]
            i += wc
            if op == 5:
                self.names[a[0]] = self._str(a[1:])
            elif op == 6:
                self.member_names[(a[0], a[1])] = self._str(a[2:])
            elif op == 11:
                if self._str(a[1:]) == "GLSL.std.450":
                    self.glsl_ext = a[0]
            elif op == 15:
                self.entry = a[1]
            elif op == 16:
                if a[1] == 17:
                    self.local_size = (a[2], a[3], a[4])
            elif op == 71:
                self.decor.setdefault(a[0], {})[a[1]] = list(a[2:])
            elif op == 72:
                self.mdecor.setdefault((a[0], a[1]), {})[a[2]] = list(a[3:])
            elif op == 19:
                self.types[a[0]] = ("void",)
            elif op == 20:
                self.types[a[0]] = ("bool",)
            elif op == 21:
                self.types[a[0]] = ("int", a[1], a[2])
            elif op == 22:
                self.types[a[0]] = ("float", a[1])
            elif op == 23:
                self.types[a[0]] = ("vec", a[1], a[2])
            elif op == 24:
                self.types[a[0]] = ("mat", a[1], a[2])
            elif op == 25:
                self.types[a[0]] = ("image",) + tuple(a[1:])
            elif op == 27:
                self.types[a[0]] = ("sampled_image", a[1])
            elif op == 28:
                self.types[a[0]] = ("array", a[1], a[2])
            elif op == 29:
                self.types[a[0]] = ("rtarray", a[1])
            elif op == 30:
                self.types[a[0]] = ("struct", list(a[1:]))
            elif op == 32:
                self.types[a[0]] = ("ptr", a[1], a[2])
            elif op == 33:
                self.types[a[0]] = ("func", a[1], list(a[2:]))
            elif op == 41:
                self.consts[a[1]] = True
            elif op == 42:
                self.consts[a[1]] = False
            elif op == 43:
                t = self.types[a[0]]
                self.consts[a[1]] = bits_f(a[2]) if t[0] == "float" else a[2] & M32
            elif op == 44:
                self.consts[a[1]] = [cp(self.consts[x]) for x in a[2:]]
            elif op == 46:
                self.consts[a[1]] = self.zero(a[0])
            elif op == 54:
                cur_fn = Function(a[1], a[0])
                self.functions[a[1]] = cur_fn
            elif op == 55:
                cur_fn.params.append(a[1])
            elif op == 56:
                cur_fn = None
            elif op == 59 and cur_fn is None:
                self.globals[a[1]] = (a[0], a[2], a[3] if len(a) > 3 else None)
            elif op == 248:
                cur_block = []
                cur_fn.blocks[a[0]] = cur_block
                if cur_fn.entry is None:
                    cur_fn.entry = a[0]
            elif cur_fn is not None:
                if op in (246, 247, 8, 317):        # merge hints / OpLine / OpNoLine carry no semantics
                    continue
                cur_block.append((op,) + tuple(a))
        # ---- handler table
        self.H = {}
        for name in dir(self):
            if name.startswith("op_"):
                self.H[int(name.split("_")[1])] = getattr(self, name)

    @staticmethod
    def _str(words):
        b = b"".join(struct.pack("<I", x) for x in words)
        return b.split(b"\0", 1)[0].decode()

    # ------------------------------------------------------------------------------------------------ types / layout
    def zero(self, tid):
        t = self.types[tid]
        k = t[0]
        if k == "bool":
            return False
        if k == "int":
            return 0
        if k == "float":
            return 0.0
        if k == "vec" or k == "mat":
            return [self.zero(t[1]) for _ in range(t[2])]
        if k == "array":
            return [self.zero(t[1]) for _ in range(self.consts[t[2]])]
        if k == "struct":
            return [self.zero(m) for m in t[1]]
        if k == "rtarray":
            return []
        raise ValueError("zero of " + k)

    def read(self, tid, buf, off, mstride=None):
        """deserialise a value of type tid from bytes at off using the module's explicit layout decorations"""
        t = self.types[tid]
        k = t[0]
        if k == "int":
            return struct.unpack_from("<I", buf, off)[0]
        if k == "float":
            return struct.unpack_from("<f", buf, off)[0]
        if k == "vec":
            return [self.read(t[1], buf, off + 4 * j) for j in range(t[2])]
        if k == "mat":
            return [self.read(t[1], buf, off + mstride * j) for j in range(t[2])]
        if k == "array":
            st = self.decor[tid][DEC_ARRAY_STRIDE][0]
            return [self.read(t[1], buf, off + st * j, mstride) for j in range(self.consts[t[2]])]
        if k == "rtarray":
            st = self.decor[tid][DEC_ARRAY_STRIDE][0]
            n = (len(buf) - off) // st
            return [self.read(t[1], buf, off + st * j, mstride) for j in range(n)]
        if k == "struct":
            out = []
            for m, mt in enumerate(t[1]):
                d = self.mdecor.get((tid, m), {})
                out.append(self.read(mt, buf, off + d[DEC_OFFSET][0], d.get(DEC_MATRIX_STRIDE, [None])[0]))
            return out
        raise ValueError("read of " + k)

    def write(self, tid, val, buf, off, mstride=None):
        t = self.types[tid]
        k = t[0]
        if k == "int":
            struct.pack_into("<I", buf, off, val & M32)
        elif k == "float":
            struct.pack_into("<f", buf, off, val)
        elif k == "vec":
            for j in range(t[2]):
                self.write(t[1], val[j], buf, off + 4 * j)
        elif k == "mat":
            for j in range(t[2]):
                self.write(t[1], val[j], buf, off + mstride * j)
        elif k in ("array", "rtarray"):
            st = self.decor[tid][DEC_ARRAY_STRIDE][0]
            for j, x in enumerate(val):
                self.write(t[1], x, buf, off + st * j, mstride)
        elif k == "struct":
            for m, mt in enumerate(t[1]):
                d = self.mdecor.get((tid, m), {})
                self.write(mt, val[m], buf, off + d[DEC_OFFSET][0], d.get(DEC_MATRIX_STRIDE, [None])[0])
        else:
            raise ValueError("write of " + k)

    def call_count(self, name):
        """executions of OpFunctionCall whose callee's debug name starts with `name(`"""
        return sum(n for fid, n in self.calls.items() if self.names.get(fid, "").startswith(name + "("))

    def bindings(self):
        """{binding: (variable id, pointee type id, storage class, name)} of the module's descriptor-backed variables"""
        out = {}
        for vid, (ptid, sc, _) in self.globals.items():
            d = self.decor.get(vid, {})
            if DEC_BINDING in d:
                pointee = self.types[ptid][2]
                out[d[DEC_BINDING][0]] = (vid, pointee, sc, self.names.get(vid) or self.names.get(pointee, ""))
        return out

    # ------------------------------------------------------------------------------------------------ dispatch
    def dispatch(self, groups, resources, lockstep=False, only=None):
        """Run the entry point for groups = (gx, gy, gz) workgroups.
        resources: {binding: bytearray (buffer) | Image}.  Buffers are deserialised into nested lists before the dispatch and
        written back afterwards.  only: optional predicate(global_id) -> bool to skip invocations (they would not touch memory
        anyway, e.g. out-of-image threads)."""
        base = dict(self.consts)
        bound = {}
        for b, (vid, pointee, sc, _) in self.bindings().items():
            res = resources[b]
            if isinstance(res, Image):
                base[vid] = ([res], 0)
            else:
                val = self.read(pointee, res, 0)
                bound[b] = (pointee, val, res)
                base[vid] = ([val], 0)
        lx, ly, lz = self.local_size
        gx, gy, gz = groups
        builtin_vars = {}
        for vid, (ptid, sc, _) in self.globals.items():
            d = self.decor.get(vid, {})
            if DEC_BUILTIN in d:
                builtin_vars[d[DEC_BUILTIN][0]] = vid
        for wz in range(gz):
            for wy in range(gy):
                for wx in range(gx):
                    shared = {}
                    for vid, (ptid, sc, init) in self.globals.items():
                        if sc == SC_WORKGROUP:
                            shared[vid] = ([self.zero(self.types[ptid][2])], 0)
                    gens = []
                    for lzv in range(lz):
                        for lyv in range(ly):
                            for lxv in range(lx):
                                gid = [wx * lx + lxv, wy * ly + lyv, wz * lz + lzv]
                                if only is not None and not only(gid):
                                    continue
                                lindex = (lzv * ly + lyv) * lx + lxv
                                env = dict(base)
                                env.update(shared)
                                bi = {BI_NUM_WG: [gx, gy, gz], BI_WG_SIZE: [lx, ly, lz], BI_WG_ID: [wx, wy, wz],
                                      BI_LOCAL_ID: [lxv, lyv, lzv], BI_GLOBAL_ID: gid, BI_LOCAL_INDEX: lindex,
                                      BI_SG_SIZE: 32, BI_NUM_SG: (lx * ly * lz + 31) // 32, BI_SG_ID: lindex // 32,
                                      BI_SG_LOCAL_ID: lindex % 32}
                                for b, vid in builtin_vars.items():
                                    env[vid] = ([cp(bi[b])], 0)
                                for vid, (ptid, sc, init) in self.globals.items():
                                    if sc == SC_PRIVATE or (sc in (SC_INPUT, SC_OUTPUT) and vid not in env):
                                        v0 = cp(self.consts[init]) if init is not None else self.zero(self.types[ptid][2])
                                        env[vid] = ([v0], 0)
                                g = self.call(self.entry, [], env)
                                if lockstep:
                                    gens.append((lindex, g))
                                else:
                                    for y in g:
                                        raise RuntimeError("shader yielded %r outside lock-step mode" % (y,))
                    if lockstep:
                        self._run_lockstep(gens)
        for b, (pointee, val, res) in bound.items():
            self.write(pointee, val, res, 0)

    def run_stage(self, inputs, resources):
        """One invocation of a non-compute entry point (the fullscreen fragment shader): inputs {location: value};
        returns {location: value} of the Output variables."""
        env = dict(self.consts)
        outs = {}
        for vid, (ptid, sc, init) in self.globals.items():
            d = self.decor.get(vid, {})
            pointee = self.types[ptid][2]
            if DEC_BINDING in d:
                res = resources[d[DEC_BINDING][0]]
                env[vid] = ([res if isinstance(res, Image) else self.read(pointee, res, 0)], 0)
            elif sc == SC_INPUT:
                env[vid] = ([cp(inputs[d[30][0]])], 0)                   # decoration 30 = Location
            else:
                env[vid] = ([self.zero(pointee)], 0)
                if sc == SC_OUTPUT and 30 in d:
                    outs[d[30][0]] = env[vid]
        for y in self.call(self.entry, [], env):
            raise RuntimeError("unexpected yield %r" % (y,))
        return {loc: cp(cell[0][0]) for loc, cell in outs.items()}

    def _run_lockstep(self, gens):
        """advance every invocation of a workgroup to its next barrier / subgroup instruction, resolve, repeat"""
        live = {li: g for li, g in gens}
        pending = {li: None for li in live}
        while live:
            waits = {}
            for li in sorted(live):
                try:
                    waits[li] = live[li].send(pending[li])
                except StopIteration:
                    del live[li]
            pending = {li: None for li in live}
            if not waits:
                break
            kinds = {wv[0] for wv in waits.values()}
            if kinds == {"barrier"}:
                if len(waits) != len(live):
                    raise RuntimeError("barrier not reached by every live invocation")
                continue
            # subgroup operations: group the waiting invocations by subgroup; invocations at a barrier keep waiting there
            by_sg = {}
            for li, wv in waits.items():
                if wv[0] == "barrier":
                    raise RuntimeError("mixed barrier / subgroup wait: divergent workgroup")
                by_sg.setdefault(li // 32, []).append(li)
            for sg, lanes in by_sg.items():
                lanes.sort()
                kind = {waits[li][:3] for li in lanes}
                if len(kind) != 1:
                    raise RuntimeError("subgroup lanes wait at different operations: %r" % (kind,))
                _, opname, gop = waits[lanes[0]][:3]
                vals = [waits[li][3] for li in lanes]
                res = self._subgroup(opname, gop, vals)
                for li, r in zip(lanes, res):
                    pending[li] = r

    @staticmethod
    def _subgroup(opname, gop, vals):
        if opname == "elect":
            return [i == 0 for i in range(len(vals))]
        if opname == "iadd":
            f1, ident = (lambda a, b: (a + b) & M32), 0
        elif opname == "fmin":
            f1, ident = (lambda a, b: b if b < a else a), None
        elif opname == "fmax":
            f1, ident = (lambda a, b: b if a < b else a), None
        else:
            raise NotImplementedError(opname)

        def f(a, b):      # vector operands reduce per component
            return [f1(x, y) for x, y in zip(a, b)] if type(a) is list else f1(a, b)
        if gop == 0:        # Reduce
            acc = vals[0]
            for v in vals[1:]:
                acc = f(acc, v)
            return [acc] * len(vals)
        if gop == 2:        # ExclusiveScan
            out, acc = [], ident
            for v in vals:
                out.append(acc)
                acc = f(acc, v)
            return out
        if gop == 1:        # InclusiveScan
            out, acc = [], None
            for v in vals:
                acc = v if acc is None else f(acc, v)
                out.append(acc)
            return out
        raise NotImplementedError("group operation %d" % gop)

    # ------------------------------------------------------------------------------------------------ execution
    def call(self, fid, args, v):
        """generator: executes function fid with argument values args in value environment v; returns the result value"""
        fn = self.functions[fid]
        for p, a in zip(fn.params, args):
            v[p] = a
        label, prev = fn.entry, None
        H = self.H
        types = self.types
        while True:
            block = fn.blocks[label]
            nxt = None
            # OpPhi instructions read the values of the edge taken: evaluate them together first
            j = 0
            if block and block[0][0] == 245:
                phis = []
                while j < len(block) and block[j][0] == 245:
                    ins = block[j]
                    for q in range(3, len(ins), 2):
                        if ins[q + 1] == prev:
                            phis.append((ins[2], v[ins[q]]))
                            break
                    else:
                        raise RuntimeError("phi without matching predecessor")
                    j += 1
                for rid, val in phis:
                    v[rid] = val
            n = len(block)
            self.n_executed += n
            while j < n:
                ins = block[j]
                j += 1
                op = ins[0]
                if op == 61:                                    # OpLoad
                    c, k = v[ins[3]]
                    x = c[k]
                    v[ins[2]] = cp(x) if type(x) is list else x
                elif op == 62:                                  # OpStore
                    c, k = v[ins[1]]
                    assign(c, k, v[ins[2]])
                elif op == 65 or op == 66:                      # OpAccessChain
                    c, k = v[ins[3]]
                    for q in range(4, len(ins)):
                        c = c[k]
                        k = v[ins[q]]
                        if k >= len(c):
                            raise IndexError("access chain index %d out of range (%d) in %s" % (k, len(c), self.names.get(fid, fid)))
                    v[ins[2]] = (c, k)
                elif op == 59:                                  # OpVariable (Function storage)
                    init = v[ins[4]] if len(ins) > 4 else self.zero(types[ins[1]][2])
                    v[ins[2]] = ([cp(init)], 0)
                elif op == 249:                                 # OpBranch
                    nxt = ins[1]
                elif op == 250:                                 # OpBranchConditional
                    nxt = ins[2] if v[ins[1]] else ins[3]
                elif op == 253:                                 # OpReturn
                    return None
                elif op == 254:                                 # OpReturnValue
                    return v[ins[1]]
                elif op == 57:                                  # OpFunctionCall
                    self.calls[ins[3]] = self.calls.get(ins[3], 0) + 1
                    v[ins[2]] = yield from self.call(ins[3], [v[x] for x in ins[4:]], v)
                elif op == 224:                                 # OpControlBarrier
                    yield ("barrier",)
                elif op == 333:                                 # OpGroupNonUniformElect
                    v[ins[2]] = yield ("sg", "elect", 0, None)
                elif op == 349:                                 # OpGroupNonUniformIAdd
                    v[ins[2]] = yield ("sg", "iadd", ins[4], v[ins[5]])
                elif op == 355:                                 # OpGroupNonUniformFMin
                    v[ins[2]] = yield ("sg", "fmin", ins[4], v[ins[5]])
                elif op == 358:                                 # OpGroupNonUniformFMax
                    v[ins[2]] = yield ("sg", "fmax", ins[4], v[ins[5]])
                elif op == 255:
                    raise RuntimeError("OpUnreachable executed")
                else:
                    h = H.get(op)
                    if h is None:
                        raise NotImplementedError("SPIR-V opcode %d" % op)
                    h(ins, v)
            if nxt is None:
                raise RuntimeError("block %d fell through" % label)
            prev, label = label, nxt

    # ---- helpers
    def _is_float(self, tid):
        t = self.types[tid]
        return t[0] == "float" or (t[0] == "vec" and self.types[t[1]][0] == "float")

    @staticmethod
    def _map2(f, a, b):
        if type(a) is list:
            return [f(x, y) for x, y in zip(a, b)]
        return f(a, b)

    @staticmethod
    def _map1(f, a):
        if type(a) is list:
            return [f(x) for x in a]
        return f(a)

    def _bin(fn):     # noqa: N805 -- decorator factory used while the class body is built
        def h(self, ins, v):
            v[ins[2]] = self._map2(fn, v[ins[3]], v[ins[4]])
        return h

    def _un(fn):      # noqa: N805
        def h(self, ins, v):
            v[ins[2]] = self._map1(fn, v[ins[3]])
        return h

    # ---- composites
    def op_79(self, ins, v):                                    # OpVectorShuffle
        src = v[ins[3]] + v[ins[4]]
        v[ins[2]] = [src[c] if c != M32 else 0.0 for c in ins[5:]]

    def op_80(self, ins, v):                                    # OpCompositeConstruct
        t = self.types[ins[1]]
        parts = [v[x] for x in ins[3:]]
        if t[0] == "vec":
            out = []
            for p in parts:
                if type(p) is list:
                    out.extend(p)
                else:
                    out.append(p)
            v[ins[2]] = out
        else:
            v[ins[2]] = [cp(p) for p in parts]

    def op_81(self, ins, v):                                    # OpCompositeExtract
        x = v[ins[3]]
        for q in ins[4:]:
            x = x[q]
        v[ins[2]] = cp(x)

    def op_82(self, ins, v):                                    # OpCompositeInsert
        comp = cp(v[ins[4]])
        c = comp
        for q in ins[5:-1]:
            c = c[q]
        c[ins[-1]] = cp(v[ins[3]])
        v[ins[2]] = comp

    def op_83(self, ins, v):                                    # OpCopyObject
        v[ins[2]] = cp(v[ins[3]])

    # ---- images
    def op_98(self, ins, v):                                    # OpImageRead (outside the image: zero)
        img = v[ins[3]]
        x, y = s32(v[ins[4]][0]), s32(v[ins[4]][1])
        v[ins[2]] = list(img.px[y][x]) if 0 <= x < img.w and 0 <= y < img.h else [0.0, 0.0, 0.0, 0.0]

    def op_99(self, ins, v):                                    # OpImageWrite (outside the image: discarded)
        img = v[ins[1]]
        x, y = s32(v[ins[2]][0]), s32(v[ins[2]][1])
        if 0 <= x < img.w and 0 <= y < img.h:
            img.px[y][x] = list(v[ins[3]])

    def op_87(self, ins, v):                                    # OpImageSampleImplicitLod: nearest texel, clamp to edge
        img = v[ins[3]]
        u, w = v[ins[4]][0], v[ins[4]][1]
        x = min(max(int(math.floor(f32(u * img.w))), 0), img.w - 1)
        y = min(max(int(math.floor(f32(w * img.h))), 0), img.h - 1)
        v[ins[2]] = list(img.px[y][x])

    def op_104(self, ins, v):                                   # OpImageQuerySize
        img = v[ins[3]]
        v[ins[2]] = [img.w, img.h]

    # ---- conversions
    op_109 = _un(f2u)                                           # OpConvertFToU (pin U1 / U5)
    op_110 = _un(f2s)                                           # OpConvertFToS
    op_111 = _un(lambda a: f32(float(s32(a))))                  # OpConvertSToF
    op_112 = _un(lambda a: f32(float(a)))                       # OpConvertUToF

    def op_124(self, ins, v):                                   # OpBitcast
        to_float = self._is_float(ins[1])
        x = v[ins[3]]

        def one(a):
            if to_float:
                return bits_f(a) if type(a) is int else a
            return f_bits(a) if type(a) is float else a
        v[ins[2]] = self._map1(one, x)

    # ---- arithmetic
    op_126 = _un(lambda a: (-a) & M32)                          # OpSNegate
    op_127 = _un(lambda a: -a)                                  # OpFNegate
    op_128 = _bin(lambda a, b: (a + b) & M32)                   # OpIAdd
    op_129 = _bin(lambda a, b: f32(a + b))                      # OpFAdd
    op_130 = _bin(lambda a, b: (a - b) & M32)                   # OpISub
    op_131 = _bin(lambda a, b: f32(a - b))                      # OpFSub
    op_132 = _bin(lambda a, b: (a * b) & M32)                   # OpIMul
    op_133 = _bin(lambda a, b: f32(a * b))                      # OpFMul
    op_134 = _bin(lambda a, b: (a // b) if b else M32)          # OpUDiv
    op_136 = _bin(fdiv)                                         # OpFDiv
    op_137 = _bin(lambda a, b: (a % b) if b else 0)             # OpUMod

    def op_135(self, ins, v):                                   # OpSDiv (truncating)
        def sdiv(a, b):
            a, b = s32(a), s32(b)
            if b == 0:
                return M32
            q = abs(a) // abs(b)
            return (q if (a < 0) == (b < 0) else -q) & M32
        v[ins[2]] = self._map2(sdiv, v[ins[3]], v[ins[4]])

    def op_142(self, ins, v):                                   # OpVectorTimesScalar
        s = v[ins[4]]
        v[ins[2]] = [f32(x * s) for x in v[ins[3]]]

    def op_145(self, ins, v):                                   # OpMatrixTimesVector: ((c0*x + c1*y) + c2*z) + c3*w
        m, x = v[ins[3]], v[ins[4]]
        rows = len(m[0])
        out = []
        for r in range(rows):
            acc = f32(m[0][r] * x[0])
            for c in range(1, len(m)):
                acc = f32(acc + f32(m[c][r] * x[c]))
            out.append(acc)
        v[ins[2]] = out

    def op_148(self, ins, v):                                   # OpDot: ((a0*b0 + a1*b1) + a2*b2) + ...
        a, b = v[ins[3]], v[ins[4]]
        acc = f32(a[0] * b[0])
        for q in range(1, len(a)):
            acc = f32(acc + f32(a[q] * b[q]))
        v[ins[2]] = acc

    # ---- logic / comparisons
    op_164 = _bin(lambda a, b: a == b)                          # OpLogicalEqual
    op_165 = _bin(lambda a, b: a != b)                          # OpLogicalNotEqual
    op_166 = _bin(lambda a, b: a or b)                          # OpLogicalOr
    op_167 = _bin(lambda a, b: a and b)                         # OpLogicalAnd
    op_168 = _un(lambda a: not a)                               # OpLogicalNot

    def op_169(self, ins, v):                                   # OpSelect
        c, a, b = v[ins[3]], v[ins[4]], v[ins[5]]
        if type(c) is list:
            v[ins[2]] = [x if cc else y for cc, x, y in zip(c, a, b)]
        else:
            v[ins[2]] = cp(a) if c else cp(b)

    op_170 = _bin(lambda a, b: a == b)                          # OpIEqual
    op_171 = _bin(lambda a, b: a != b)                          # OpINotEqual
    op_172 = _bin(lambda a, b: a > b)                           # OpUGreaterThan
    op_173 = _bin(lambda a, b: s32(a) > s32(b))                 # OpSGreaterThan
    op_174 = _bin(lambda a, b: a >= b)                          # OpUGreaterThanEqual
    op_175 = _bin(lambda a, b: s32(a) >= s32(b))                # OpSGreaterThanEqual
    op_176 = _bin(lambda a, b: a < b)                           # OpULessThan
    op_177 = _bin(lambda a, b: s32(a) < s32(b))                 # OpSLessThan
    op_178 = _bin(lambda a, b: a <= b)                          # OpULessThanEqual
    op_179 = _bin(lambda a, b: s32(a) <= s32(b))                # OpSLessThanEqual
    op_180 = _bin(lambda a, b: a == b)                          # OpFOrdEqual
    op_182 = _bin(lambda a, b: a < b or a > b)                  # OpFOrdNotEqual
    op_183 = _bin(lambda a, b: a != b)                          # OpFUnordNotEqual
    op_184 = _bin(lambda a, b: a < b)                           # OpFOrdLessThan
    op_186 = _bin(lambda a, b: a > b)                           # OpFOrdGreaterThan
    op_188 = _bin(lambda a, b: a <= b)                          # OpFOrdLessThanEqual
    op_190 = _bin(lambda a, b: a >= b)                          # OpFOrdGreaterThanEqual

    # ---- bits
    op_194 = _bin(lambda a, b: (a >> (b & 31)) & M32)           # OpShiftRightLogical
    op_195 = _bin(lambda a, b: (s32(a) >> (b & 31)) & M32)      # OpShiftRightArithmetic
    op_196 = _bin(lambda a, b: (a << (b & 31)) & M32)           # OpShiftLeftLogical
    op_197 = _bin(lambda a, b: a | b)                           # OpBitwiseOr
    op_198 = _bin(lambda a, b: a ^ b)                           # OpBitwiseXor
    op_199 = _bin(lambda a, b: a & b)                           # OpBitwiseAnd
    op_200 = _un(lambda a: (~a) & M32)                          # OpNot
    op_205 = _un(lambda a: bin(a).count("1"))                   # OpBitCount

    # ---- atomics (invocations run one after the other, so read-modify-write is trivially atomic)
    def _atomic(fn):  # noqa: N805
        def h(self, ins, v):
            c, k = v[ins[3]]
            old = c[k]
            c[k] = fn(old, v[ins[6]]) & M32
            v[ins[2]] = old
        return h

    op_234 = _atomic(lambda a, b: a + b)                        # OpAtomicIAdd
    op_235 = _atomic(lambda a, b: a - b)                        # OpAtomicISub
    op_240 = _atomic(lambda a, b: a & b)                        # OpAtomicAnd
    op_241 = _atomic(lambda a, b: a | b)                        # OpAtomicOr

    def op_227(self, ins, v):                                   # OpAtomicLoad
        c, k = v[ins[3]]
        v[ins[2]] = c[k]

    def op_228(self, ins, v):                                   # OpAtomicStore
        c, k = v[ins[1]]
        c[k] = v[ins[4]]

    def op_225(self, ins, v):                                   # OpMemoryBarrier
        pass

    # ---- GLSL.std.450
    def op_12(self, ins, v):                                    # OpExtInst
        if ins[3] != self.glsl_ext:
            raise NotImplementedError("extended instruction set")
        e = ins[4]
        a = [v[x] for x in ins[5:]]
        m1, m2 = self._map1, self._map2
        if e == 4:
            r = m1(abs, a[0])                                                   # FAbs
        elif e == 11:
            r = m1(lambda x: f32(x * 0.01745329238474369), a[0])               # Radians: x * fl(pi/180)
        elif e == 13:
            r = m1(lambda x: self.sincos(x)[0], a[0])                           # Sin
        elif e == 14:
            r = m1(lambda x: self.sincos(x)[1], a[0])                           # Cos
        elif e == 15:
            r = m1(lambda x: f32(_libm.tanf(x)), a[0])                          # Tan
        elif e == 17:
            r = m1(lambda x: f32(_libm.acosf(x)), a[0])                         # Acos
        elif e == 25:
            r = m2(lambda y, x: f32(_libm.atan2f(y, x)), a[0], a[1])            # Atan2
        elif e == 31:
            r = m1(fsqrt, a[0])                                                 # Sqrt
        elif e == 37:
            r = m2(lambda x, y: y if y < x else x, a[0], a[1])                  # FMin
        elif e == 40:
            r = m2(lambda x, y: y if x < y else x, a[0], a[1])                  # FMax
        elif e == 38:
            r = m2(min, a[0], a[1])                                             # UMin
        elif e == 41:
            r = m2(max, a[0], a[1])                                             # UMax
        elif e == 39:
            r = m2(lambda x, y: x if s32(x) <= s32(y) else y, a[0], a[1])       # SMin
        elif e == 42:
            r = m2(lambda x, y: x if s32(x) >= s32(y) else y, a[0], a[1])       # SMax
        elif e == 43:                                                           # FClamp = FMin(FMax(x, lo), hi)
            def clamp(x, lo, hi):
                t = lo if x < lo else x
                return hi if hi < t else t
            if type(a[0]) is list:
                r = [clamp(x, lo, hi) for x, lo, hi in zip(a[0], a[1], a[2])]
            else:
                r = clamp(a[0], a[1], a[2])
        elif e == 68:                                                           # Cross
            p, q = a[0], a[1]
            r = [f32(f32(p[1] * q[2]) - f32(p[2] * q[1])), f32(f32(p[2] * q[0]) - f32(p[0] * q[2])),
                 f32(f32(p[0] * q[1]) - f32(p[1] * q[0]))]
        elif e == 69:                                                           # Normalize = v / sqrt(dot(v, v))
            p = a[0]
            acc = f32(p[0] * p[0])
            for q in range(1, len(p)):
                acc = f32(acc + f32(p[q] * p[q]))
            ln = fsqrt(acc)
            r = [fdiv(x, ln) for x in p]
        elif e == 75:
            r = m1(lambda x: (x.bit_length() - 1) & M32, a[0])                  # FindUMsb (0 -> -1)
        elif e == 74:
            r = m1(lambda x: ((x if not x & 0x80000000 else (~x) & M32).bit_length() - 1) & M32, a[0])   # FindSMsb
        else:
            raise NotImplementedError("GLSL.std.450 instruction %d" % e)
        v[ins[2]] = r
