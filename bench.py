#!/usr/bin/env python
"""bench.py -- Mrays/s of the path-tracer hot path (S1 BVH build + S2 trace + resolve) on N B200s.

    python bench.py --gpus N --steps K --warmup W [--config C2] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

One "step" = one frame of the named config, exactly what RaytracerBVHRenderer::Raytracer::doIteration does per
frame (RaytracerBVH.hpp:206-496): fresh model-space scene arrays -> clear -> K1..K6 BVH build -> all samples of the
frame -> resolve to RGBA8.  N > 1: the frame's 8-row bands are interleaved over the ranks (scene replicated, BVH built
per rank, no data-path collective while tracing); ONE exchange through the library's own collective entry point
(rtb_gather_tiles: NCCL all-gather over NVLink + fused re-assembly / resolve) ends the frame (strong scaling).

`value`   : REFERENCE-EQUIVALENT rays per second -- the hitBVH calls the reference shader executes for this frame (counted on
            the device by the instrumented reference-order walk, RTB_TRACE_COUNT) / device time of the step, scene arrays
            resident in HBM.  The frame is bit-identical to the reference's, but the kernels walk FEWER rays than that (a pixel's
            un-jittered primary ray is traced once, not once per sample): `mrays_traversed_per_s` next to it is the rate of rays
            really walked (RTB_TRACE_WALK_COUNT: the production kernels' own counters).
`e2e`     : the same frame through the C-ABI with HOST buffers: pinned host scene arrays uploaded (each rank 1/N of them over its
            own PCIe link, the rest over NVLink: rtb_comm_all_gather) and the resolved RGBA8 frame read back, every step, inside
            the timed region; uploads / read-backs run on their own streams and overlap the neighbouring frames.
`roofline`: the trace launches only.  achieved = bytes the production kernels fetch + store (their own counters x the record sizes
            in HBM) / CUDA-event time, against the measured HBM copy bandwidth; `node_fetch` compares the same fetch rate with a
            live micro-benchmark of divergent 64-byte record fetches at the scene's footprint (rtb_probe_gather); the
            reference-equivalent figure of SURVEY.md 8d is kept under `reference_equivalent`.
`frame_check`: every measured config renders its frame twice more, untimed -- the production walk and the exact-record walk in the
            reference's order without primary-hit sharing -- and compares the fp32 images bit for bit on the device; at N > 1 the
            assembled frame is also compared with a 1-GPU render on rank 0.
`sustained`: the K-step timed region is a burst (< 1 s); the same step is then repeated for >= --min-seconds and reported too.
`breakdown.configs`: compact results of the other BASELINE configs (C1, C3, C4, C5; C3 / C5 also with the material extension).
`--impl reference`: the reference's own algorithm on the host cores = the CPU oracle (the reference's GLSL cannot run
            here: no Vulkan ICD / lavapipe in the image), timed on a bounded sample of the same workload.
"""
import argparse
import datetime
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BAND_ROWS = 8
FRAMES_IN_FLIGHT = 1
METRIC = "Mrays/s per scene (path segments = hitBVH calls per second)"
VALUE_COUNTS = ("reference-equivalent rays: the hitBVH calls the reference shader executes for this frame (bit-identical output); "
                "rays really walked are reported as mrays_traversed_per_s")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C2", help="C1..C5 of BASELINE.md (default: C2, the metric's single-GPU config)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (parity / debugging only)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-seconds", type=float, default=3.0, help="length of the sustained leg (0 = skip)")
    ap.add_argument("--breakdown", default="auto", help="comma list of extra configs measured compactly (e.g. C1,C3,C3x,C4,C5,C5x; x = material "
                                                        "extension), 'auto' = all of them at N = 1 / C4 tiles + C5 sample ranges at N > 1, 'none'")
    ap.add_argument("--ext", action="store_true", help="main config in extension mode (RTB_TRACE_EXT_MATERIALS: metal / dielectric scatter)")
    ap.add_argument("--kernel", default="wave", choices=["wave", "simple", "stream"], help="trace kernel (A/B switch; simple / stream need a -DRTB_AB_KERNELS build)")
    ap.add_argument("--reference-order", action="store_true",
                    help="walk the tree in the reference's visiting order with no t-interval, like the shader (A/B switch, same results)")
    ap.add_argument("--no-primary-sharing", action="store_true",
                    help="trace every sample's (identical, un-jittered) primary ray separately like the shader does (A/B switch, same results)")
    ap.add_argument("--nodes", default="auto", choices=["auto", "exact", "wide"],
                    help="traversal records: auto (library default: 4-ary for >= 512 primitives), exact 64-byte child pairs, "
                         "64-byte 4-ary (A/B switch, same results)")
    ap.add_argument("--emulate-rank", default="", help="debugging: 'r/n' renders only the bands rank r of n would own, on one GPU, "
                                                         "without the collective (the per-rank workload of a tile-mode run, e.g. for ncu)")
    ap.add_argument("--shard", default="tiles", choices=["tiles", "samples"],
                    help="N > 1: tiles = 8-row bands + all-gather (bit-identical, default); samples = sample ranges + sum-reduce (C5)")
    ap.add_argument("--mode", default="exact", choices=["exact", "culled"],
                    help="exact = the reference's traversal (default, the headline); culled = extension RTB_TRACE_CULLED")
    ap.add_argument("--no-frame-check", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true", help="experiment: skip the 256 MiB write between steps (a C2 frame's own working set, "
                                                                "2.3 GB of sample slots + records, is 18x the L2 in any case)")
    ap.add_argument("--frames-in-flight", type=int, default=FRAMES_IN_FLIGHT,
                    help="timed / sustained / e2e legs: this many frames are in flight, each on its own library context + stream + buffers "
                         "(frame k+1's BVH build runs while frame k's trace launch drains); 1 = strictly one frame after the other")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload(cfg_name, spp_override=0, ext=False):
    from raytracergpu_mastersproject_b200 import make_ubo, scenes
    cfg = dict(scenes.CONFIGS[cfg_name])
    sc = scenes.load_scene(cfg["spec"])
    spp = spp_override or cfg["spp"]
    ubo = make_ubo(len(sc["triangles"]), len(sc["spheres"]), len(sc["materials"]), sc["max_depth"], cfg["random_state"], sc["vfov"])
    desc = {
        "workload": f"{cfg_name}: {cfg['spec']} ({len(sc['triangles'])} triangles, {len(sc['spheres'])} spheres), "
                    f"{cfg['width']}x{cfg['height']}, {spp} spp, max depth {sc['max_depth']}, randomState {cfg['random_state']}",
        "scene": cfg["spec"], "triangles": len(sc["triangles"]), "spheres": len(sc["spheres"]),
        "width": cfg["width"], "height": cfg["height"], "spp": spp, "max_depth": sc["max_depth"],
        "mode": ("EXTENSION RTB_TRACE_EXT_MATERIALS: metal / dielectric scatter (no reference behaviour exists; frame bit-identical to the oracle's extension)"
                 if ext else "reference-parity (metal/dielectric absorb like the shader; frame bit-identical to the oracle)"),
    }
    return cfg, sc, ubo, spp, desc


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 20 ms from well before the warm-up (it needs a few hundred ms to start); only the samples whose
    own timestamp falls inside a queried window (+- one period) are reported."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.rows = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    @staticmethod
    def now():
        return datetime.datetime.now()

    def stop(self):
        if self.p is None:
            self.rows = []
            return
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        rows = []
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(parts[2]), float(parts[3]), float(parts[4]), [n for n, v in zip(self.NAMES, parts[6:10]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        self.rows = rows

    def window(self, t0, t1, fallback_from=None):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        pad = datetime.timedelta(milliseconds=20)
        window, scope = [r for r in self.rows if t0 - pad <= r[0] <= t1 + pad], "timed region"
        if not window and fallback_from is not None:
            window, scope = [r for r in self.rows if fallback_from <= r[0] <= t1 + pad], "warm-up + timed region"
        if not window:      # region shorter than the sampling period: the samples right around it
            near = datetime.timedelta(milliseconds=500)
            window, scope = [r for r in self.rows if t0 - near <= r[0] <= t1 + near], "within 0.5 s of the timed region (shorter than the sampling period)"
        if not window:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}
        reasons = sorted({n for r in window for n in r[4]})
        return {"sm_mhz": float(np.median([r[1] for r in window])), "sm_max_mhz": float(max(r[2] for r in window)), "reasons": reasons,
                "power_w_median": float(np.median([r[3] for r in window])), "samples": len(window), "scope": scope}


# ----------------------------------------------------------------------------------------------------------------
# CPU oracle legs (the ONLY place bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------------------------
def cpu_oracle_sample(sc, ubo, W, H, seconds, max_spp=64, ext=False):
    """Times the oracle (C restatement of the reference shaders, OpenMP over rows, ALL host cores whatever OMP_NUM_THREADS says)
    on a bounded sample of the workload: the full frame at 1 spp, repeated (as further samples of the same frame) until
    `seconds` is used up.  The BVH is built once, as for a frame, and timed separately."""
    from oracle import oracle as O
    threads = O.host_threads()
    opt = O.make_options(threads=threads, ext_materials=ext)
    t0 = time.time()
    b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    t_build = time.time() - t0
    img = None
    rays = 0; t_trace = 0.0; spp = 0
    while spp < max_spp and (spp == 0 or t_trace + t_trace / spp <= seconds):
        t0 = time.time()
        r = O.raytrace(ubo, W, H, b["tris"], b["sphs"], sc["materials"], b["nodes"], 1, image=img, opt=opt, want_hits=False, want_rng=False)
        t_trace += time.time() - t0
        img = r["image"]; rays += r["counters"]["rays"]; spp += 1
    return dict(mrays=rays / t_trace / 1e6, rays=rays, t_trace=t_trace, t_build=t_build, spp=spp, cores=threads)


def run_reference(args):
    """--impl reference: the reference's own CPU-runnable implementation of the path.  Its GLSL cannot execute in this
    image (no Vulkan loader / lavapipe ICD), so this is the oracle port on ALL host cores (the thread count is
    set explicitly: launchers such as torchrun export OMP_NUM_THREADS=1).  Like the B200 arm it builds the BVH once per frame:
    one step = one of the frame's `spp` samples per pixel over the whole image (a bounded sample of the frame); the build is timed
    once and charged to every step as build / spp, so value = rays of a frame / (build + spp x step) -- the rate of whole frames."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, sc, ubo, spp, desc = workload(args.config, args.spp, args.ext)
    from oracle import oracle as O
    W, H = cfg["width"], cfg["height"]
    threads = O.host_threads()
    opt = O.make_options(threads=threads, ext_materials=args.ext)
    t0 = time.time()
    b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    t_build = time.time() - t0
    times, rays = [], 0
    img = None
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        r = O.raytrace(ubo, W, H, b["tris"], b["sphs"], sc["materials"], b["nodes"], 1, image=img, opt=opt, want_hits=False, want_rng=False)
        dt = time.time() - t0
        img = r["image"]
        if i >= args.warmup:
            times.append(dt); rays += r["counters"]["rays"]
    n = max(len(times), 1)
    step = sum(times) / n + t_build / spp              # one sample of the frame + its share of the frame's one BVH build
    value = (rays / n) / step / 1e6
    sample = (f"full {W}x{H} frame, 1 of {spp} spp per step (consecutive samples of the same frame), BVH built once "
              f"({t_build:.3f} s, charged as build/{spp} per step); trace alone {rays / sum(times) / 1e6:.3f} Mrays/s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": desc,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample,
                         "bvh_build_s": t_build, "trace_mrays_per_s": rays / sum(times) / 1e6,
                         "note": "CPU oracle = C restatement of the reference shaders (gcc -O3 -march=native -fopenmp); lavapipe/Vulkan unavailable in image"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------------------------------
class Session:
    """One rank: its GPU, the stream torch and librtb200 share, the library context and (N > 1) its communicator."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from raytracergpu_mastersproject_b200 import Device, capi
        self.torch, self.dist, self.capi = torch, dist, capi
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # NCCL prints its version banner on stdout; keep stdout to the one JSON line the contract asks for
            if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"
            os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join(tempfile.gettempdir(), "rtb200_nccl_%h_%p.log"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback (use --impl reference for the CPU oracle)")
        torch.cuda.set_device(self.local_rank)
        self.tdev = torch.device("cuda", self.local_rank)
        if self.world > 1:      # torch.distributed: rendezvous, barriers and max-over-ranks of the timings (plumbing); the data path uses rtb_comm_*
            dist.init_process_group("nccl", device_id=self.tdev)
        # a non-default stream shared by torch (buffer copies, events) and librtb200 (all kernels and collectives)
        self.stream = torch.cuda.Stream(device=self.tdev)
        torch.cuda.set_stream(self.stream)
        self.L = capi.lib()
        self.dev = Device(self.local_rank, stream=self.stream.cuda_stream)
        self.h = self.dev.handle
        if self.world > 1:
            ids = [Device.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            self.dev.comm_init(self.world, self.rank, ids[0])
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.tdev)      # > 126 MB L2
        self._aux = {}
        self._lanes = [self]

    def lanes(self, n):
        """n views of this rank, one per frame in flight: lane 0 is the session itself, the others own a further library context
        (and communicator) on a stream of their own.  Collective: every rank calls it with the same n."""
        import copy
        from raytracergpu_mastersproject_b200 import Device
        while len(self._lanes) < n:
            ln = copy.copy(self)
            ln.stream = self.torch.cuda.Stream(device=self.tdev)
            ln.dev = Device(self.local_rank, stream=ln.stream.cuda_stream)
            ln.h = ln.dev.handle
            if self.world > 1:
                ids = [Device.comm_unique_id() if self.rank == 0 else None]
                self.dist.broadcast_object_list(ids, src=0)
                ln.dev.comm_init(self.world, self.rank, ids[0])
            self._lanes.append(ln)
        return self._lanes[:n]

    def aux_device(self, name, comm=False):
        """a second context of the same GPU on its own stream (H2D / D2H copy engines overlap the compute stream); comm: with a
        communicator of its own (collective: every rank asks for it at the same point)"""
        if name not in self._aux:
            from raytracergpu_mastersproject_b200 import Device
            st = self.torch.cuda.Stream(device=self.tdev)
            d = Device(self.local_rank, stream=st.cuda_stream)
            if comm and self.world > 1:
                ids = [Device.comm_unique_id() if self.rank == 0 else None]
                self.dist.broadcast_object_list(ids, src=0)
                d.comm_init(self.world, self.rank, ids[0])
            self._aux[name] = (d, st)
        return self._aux[name]

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.tdev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def sum_over_ranks(self, vals):
        if self.world == 1:
            return [int(v) for v in vals]
        t = self.torch.tensor([int(v) for v in vals], dtype=self.torch.int64, device=self.tdev)
        self.dist.all_reduce(t)
        return [int(x) for x in t]

    def all_ok(self, ok):
        return bool(self.sum_over_ranks([0 if ok else 1])[0] == 0)

    def close(self):
        for d, _ in self._aux.values():
            d.close()
        for ln in self._lanes[1:]:
            ln.dev.close()
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()
        self.dev.close()


class Frame:
    """One BASELINE config resident on this rank's GPU: scene arrays, buffers, and the frame as doIteration submits it."""

    def __init__(self, S, args, cfg_name, ext=False, shard="tiles", main=False):
        import ctypes as C
        from raytracergpu_mastersproject_b200.sharding import BandLayout, sample_range, single_gpu_layout
        torch, capi = S.torch, S.capi
        self.S, self.C, self.args, self.name, self.ext, self.main = S, C, args, cfg_name, ext, main
        cfg, sc, ubo, spp, desc = workload(cfg_name, args.spp if main else 0, ext)
        self.cfg, self.sc, self.ubo, self.spp, self.desc = cfg, sc, ubo, spp, desc
        self.W, self.H = cfg["width"], cfg["height"]
        self.T, self.Sn, self.M = len(sc["triangles"]), len(sc["spheres"]), len(sc["materials"])
        self.ubo_p = ubo.ctypes.data_as(C.c_void_p)
        world, rank = S.world, S.rank
        self.by_samples = world > 1 and shard == "samples"
        self.tiled = world > 1 and not self.by_samples
        emu = tuple(int(x) for x in args.emulate_rank.split("/")) if (main and args.emulate_rank and world == 1) else None
        self.emu = emu
        self.layout = (BandLayout(self.H, world, BAND_ROWS) if self.tiled else BandLayout(self.H, emu[1], BAND_ROWS) if emu else single_gpu_layout(self.H))
        self.rows = self.layout.local_rows
        self.my_first, self.my_count = sample_range(spp, world, rank) if self.by_samples else (0, spp)
        tdev = S.tdev

        def to_dev(a, pad_to=0):
            b = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
            t = torch.zeros(max(pad_to, b.size, 64), dtype=torch.uint8, device=tdev)
            t[: b.size].copy_(torch.from_numpy(b.copy()))
            return t
        # resident inputs: pristine model-space arrays + working copies (K1 transforms in place, so every frame starts
        # from the model-space data -- the reference re-uploads it, RaytraceScene.cpp:78-113)
        self.host = {k: np.ascontiguousarray(sc[k]).view(np.uint8).reshape(-1) for k in ("models", "materials", "triangles", "spheres")}
        # slices for the sliced upload of the e2e leg: n x bytes_per_rank, 256-byte granules
        self.slice = {k: ((self.host[k].size + world * 256 - 1) // (world * 256)) * 256 for k in ("triangles", "spheres")}
        self.d_models = to_dev(sc["models"]); self.d_mats = to_dev(sc["materials"])
        self.d_tris0 = to_dev(sc["triangles"], self.slice["triangles"] * world)
        self.d_sphs0 = to_dev(sc["spheres"], self.slice["spheres"] * world)
        self.d_tris = torch.empty_like(self.d_tris0); self.d_sphs = torch.empty_like(self.d_sphs0)
        self.image = torch.empty((self.rows, self.W, 4), dtype=torch.float32, device=tdev)
        self.final = torch.empty((self.H, self.W, 4), dtype=torch.float32, device=tdev) if world > 1 else self.image
        self.rgba8 = torch.empty((self.H, self.W, 4), dtype=torch.uint8, device=tdev)
        self.counters = torch.zeros(6, dtype=torch.int64, device=tdev)
        self.walk = torch.zeros(16, dtype=torch.int64, device=tdev)
        a = capi.TraceArgs()
        a.imageWidth, a.imageHeight, a.localRows = self.W, self.H, self.rows
        a.bandRows, a.bandFirst, a.bandStep = self.layout.band_rows, (rank if self.tiled else 0), (world if self.tiled else 1)
        if emu:
            a.bandFirst, a.bandStep = emu[0], emu[1]
        a.sampleSkip, a.sampleCount, a.flags = self.my_first, self.my_count, 0
        self.targs = a
        self.base_flags = ((capi.TRACE_EXT_MATERIALS if ext else 0)
                           | ({"simple": capi.TRACE_SIMPLE_KERNEL, "stream": capi.TRACE_STREAM_KERNEL}.get(args.kernel, 0) if main else 0)
                           | ((capi.TRACE_CULLED if args.mode == "culled" else 0) if main else 0)
                           | ((capi.TRACE_NO_PRIMARY_SHARING if args.no_primary_sharing else 0) if main else 0)
                           | ((capi.TRACE_REFERENCE_ORDER if args.reference_order else 0) if main else 0)
                           | ({"wide": capi.TRACE_WIDE_NODES, "exact": capi.TRACE_EXACT_NODES}.get(args.nodes, 0) if main else 0))
        self.production = main is False or (args.kernel == "wave" and args.mode == "exact")

    def vp(self, t):
        return self.C.c_void_p(t.data_ptr())

    def twin(self, lane):
        """the same frame for another lane (Session.lanes): shares the read-only inputs, owns what a frame writes"""
        import copy
        torch, t = self.S.torch, copy.copy(self)
        t.S = lane
        t.d_tris, t.d_sphs = torch.empty_like(self.d_tris0), torch.empty_like(self.d_sphs0)
        t.image = torch.empty_like(self.image)
        t.final = torch.empty_like(self.final) if self.S.world > 1 else t.image
        t.rgba8 = torch.empty_like(self.rgba8)
        t.targs = type(self.targs).from_buffer_copy(self.targs)
        t._twins = [t]
        return t

    def in_flight(self, n):
        """[self, twins...]: the frames that alternate in the timed legs"""
        if not hasattr(self, "_twins"):
            self._twins = [self]
        lanes = self.S.lanes(n)
        while len(self._twins) < n:
            self._twins.append(self.twin(lanes[len(self._twins)]))
        return self._twins[:n]

    def flushed_frame(self, **kw):
        """one timed step on this frame's lane: L2 flush, then the frame"""
        S = self.S
        with S.torch.cuda.stream(S.stream):
            if not self.args.no_l2_flush:
                S.flush.zero_()
            self.frame(**kw)

    # -- the frame: restore -> clear -> S1 -> S2 -> (exchange) -> resolve ------------------------------------------------
    def build(self, tris=None, sphs=None, models=None, mats=None):
        S, capi = self.S, self.S.capi
        if tris is None:
            self.d_tris.copy_(self.d_tris0); self.d_sphs.copy_(self.d_sphs0)
            tris, sphs = self.d_tris, self.d_sphs
        models = self.d_models if models is None else models
        mats = self.d_mats if mats is None else mats
        capi.check(S.L.rtb_clear_image(S.h, self.vp(self.image), self.W, self.rows))
        capi.check(S.L.rtb_build_bvh(S.h, self.ubo_p, self.vp(models), self.vp(tris), self.vp(sphs), self.vp(mats), None, None, None, None, None, 0))

    def trace(self, extra_flags=0, counters=None, walk=None, events=None):
        S, capi, a = self.S, self.S.capi, self.targs
        a.flags = self.base_flags | extra_flags
        a.counters = counters.data_ptr() if counters is not None else None
        a.walkCounters = walk.data_ptr() if walk is not None else None
        if events:
            events[0].record(S.stream)
        capi.check(S.L.rtb_raytrace(S.h, self.ubo_p, self.vp(self.image), self.C.byref(a)))
        if events:
            events[1].record(S.stream)
        a.counters = None; a.walkCounters = None

    def finish(self, want_f32=False, rgba8=None):
        """the exchange that ends a multi-GPU frame (library entry points, in-stream) + resolve"""
        S, capi = self.S, self.S.capi
        out = self.rgba8 if rgba8 is None else rgba8
        if self.by_samples:
            S.dev.reduce_samples(self.image.data_ptr(), self.W, self.H, 0, self.spp, out.data_ptr())
        elif self.tiled:
            S.dev.gather_tiles(self.image.data_ptr(), self.W, self.H, BAND_ROWS, self.final.data_ptr() if want_f32 else None, self.spp, out.data_ptr())
        else:
            capi.check(S.L.rtb_resolve_rgba8(S.h, self.vp(self.image), self.W, self.rows if self.emu else self.H, self.spp, self.vp(out)))

    def frame(self, events=None, tris=None, sphs=None, rgba8=None, models=None, mats=None):
        self.build(tris, sphs, models, mats)
        self.trace(events=events)
        self.finish(rgba8=rgba8)

    # -- counters + frame check (untimed) ---------------------------------------------------------------------------------
    def count_and_check(self):
        """Two untimed frames: the instrumented reference-order walk over the exact records without primary-hit sharing
        (RTB_TRACE_COUNT: the reference's work, and its image) and the instrumented production walk (RTB_TRACE_WALK_COUNT: what the
        timed kernels fetch, and their image).  The two fp32 images must be equal bit for bit (this rank's part of the frame)."""
        S, capi, torch = self.S, self.S.capi, self.S.torch
        self.counters.zero_(); self.walk.zero_()
        self.build()
        self.trace(extra_flags=capi.TRACE_COUNT, counters=self.counters)
        ref_img = self.image.clone()
        torch.cuda.synchronize()
        self.local_cnt = dict(zip(capi.COUNTER_FIELDS, [int(x) for x in self.counters.tolist()]))
        if self.production:
            self.build()
            self.trace(extra_flags=capi.TRACE_WALK_COUNT, walk=self.walk)
        else:
            self.build(); self.trace()
        torch.cuda.synchronize()
        self.local_walk = dict(zip(capi.WALK_COUNTER_FIELDS, [int(x) for x in self.walk.tolist()])) if self.production else None
        same = bool(torch.equal(ref_img.view(torch.int32), self.image.view(torch.int32)))
        del ref_img
        tot = S.sum_over_ranks(list(self.local_cnt.values()))
        self.cnt = dict(zip(capi.COUNTER_FIELDS, tot))
        self.rays = self.cnt["rays"]
        if self.production:
            self.walk_tot = dict(zip(capi.WALK_COUNTER_FIELDS, S.sum_over_ranks(list(self.local_walk.values()))))
            self.rays_traversed = self.walk_tot["rays"]
        else:
            self.walk_tot = None
            self.rays_traversed = self.rays
        self.check = {"status": "ok" if S.all_ok(same) else "MISMATCH",
                      "production_vs_exact_reference_order_walk": "fp32 accumulation image equal bit for bit on every rank" if S.all_ok(same) else "DIFFERENT"}
        return self.check

    def check_assembled(self):
        """N > 1: the frame the ranks assembled through the library's collective against a 1-GPU render of the whole frame on rank 0
        (tile mode: bit for bit; sample-range mode: |a - b| <= 1e-5 (1 + |b|), fp32 summation order differs by construction)."""
        S, capi, torch = self.S, self.S.capi, self.S.torch
        if S.world == 1:
            return
        self.build(); self.trace(); self.finish(want_f32=True)
        torch.cuda.synchronize()
        assembled = self.image if self.by_samples else self.final
        ok, how = True, ""
        if S.rank == 0:
            a = capi.TraceArgs()
            a.imageWidth, a.imageHeight, a.localRows = self.W, self.H, self.H
            a.bandRows, a.bandFirst, a.bandStep = self.H, 0, 1
            a.sampleSkip, a.sampleCount, a.flags = 0, self.spp, self.base_flags
            full = torch.empty((self.H, self.W, 4), dtype=torch.float32, device=S.tdev)
            rg = torch.empty((self.H, self.W, 4), dtype=torch.uint8, device=S.tdev)
            capi.check(S.L.rtb_clear_image(S.h, self.vp(full), self.W, self.H))
            capi.check(S.L.rtb_raytrace(S.h, self.ubo_p, self.vp(full), self.C.byref(a)))
            capi.check(S.L.rtb_resolve_rgba8(S.h, self.vp(full), self.W, self.H, self.spp, self.vp(rg)))
            torch.cuda.synchronize()
            if self.by_samples:
                err = (assembled[..., :3] - full[..., :3]).abs() / (1.0 + full[..., :3].abs())
                worst = float(err.max())
                ok = worst <= 1e-5 and bool(torch.equal(assembled[..., 3].view(torch.int32), full[..., 3].view(torch.int32)))
                how = f"sum-reduced sample ranges vs 1-GPU frame on rank 0: max |a-b|/(1+|b|) = {worst:.2e} (tolerance 1e-5), alpha (seed chain) bit-exact"
            else:
                ok = bool(torch.equal(assembled.view(torch.int32), full.view(torch.int32))) and bool(torch.equal(rg, self.rgba8))
                how = "assembled fp32 frame and RGBA8 frame equal the 1-GPU render on rank 0 bit for bit"
            del full, rg
        ok = S.all_ok(ok)
        self.check["assembled_vs_1gpu"] = how if ok else "MISMATCH: " + how
        if not ok:
            self.check["status"] = "MISMATCH"

    def frame_sha(self):
        self.S.torch.cuda.synchronize()
        return hashlib.sha256(self.rgba8.cpu().numpy().tobytes()).hexdigest()

    # -- timed legs ---------------------------------------------------------------------------------------------------------
    def timed(self, steps, warmup, in_flight=1):
        """K steps.  in_flight == 1: one frame after the other on one stream, each step and each rtb_raytrace call bracketed by events.
        in_flight > 1: the frames alternate over that many lanes; the K steps are bracketed as a whole (first lane's start -> last
        lane's end) -- every frame is complete inside the region, none is skipped or shared."""
        S = self.S
        fl = self.in_flight(in_flight)
        for i in range(max(warmup, 0) * len(fl)):
            fl[i % len(fl)].flushed_frame()
        S.barrier()
        t0 = ClockSampler.now()
        launches0 = sum(f.S.dev.launch_count() for f in fl)
        if len(fl) == 1:
            step_ev, trace_ev = [], []
            for _ in range(steps):
                if not self.args.no_l2_flush:
                    S.flush.zero_()
                e0, e1, a, b = S.ev(), S.ev(), S.ev(), S.ev()
                e0.record(S.stream)
                self.frame(events=(a, b))
                e1.record(S.stream)
                step_ev.append((e0, e1)); trace_ev.append((a, b))
            S.barrier()
            step_ms = sum(a.elapsed_time(b) for a, b in step_ev)
            trace_ms = sum(a.elapsed_time(b) for a, b in trace_ev) / steps
        else:
            e0 = S.ev(); e0.record(S.stream)
            for f in fl[1:]:
                f.S.stream.wait_event(e0)
            for i in range(steps):
                fl[i % len(fl)].flushed_frame()
            ends = []
            for f in fl:
                e = S.ev(); e.record(f.S.stream); ends.append(e)
            S.barrier()
            step_ms = max(e0.elapsed_time(e) for e in ends)
            trace_ms = 0.0
        t1 = ClockSampler.now()
        launches = sum(f.S.dev.launch_count() for f in fl) - launches0
        step_ms, trace_ms = S.max_over_ranks(step_ms, trace_ms)
        return dict(ms_per_step=step_ms / steps, trace_ms=trace_ms, launches=launches, t0=t0, t1=t1)

    def sustained(self, seconds, in_flight=1):
        """the same step, back to back, for at least `seconds` (clocks settle well below the burst clock under a long load)"""
        S = self.S
        fl = self.in_flight(in_flight)
        S.barrier()
        t0 = ClockSampler.now()
        e0 = S.ev()
        w0 = time.perf_counter()
        e0.record(S.stream)
        for f in fl[1:]:
            f.S.stream.wait_event(e0)
        n = 0
        while True:
            for _ in range(4 * len(fl)):
                fl[n % len(fl)].flushed_frame(); n += 1
            S.torch.cuda.synchronize()
            go = time.perf_counter() - w0 < seconds
            if S.world > 1:
                go = not S.all_ok(not go)          # everybody continues while anybody has to
            if not go:
                break
        ends = []
        for f in fl:
            e = S.ev(); e.record(f.S.stream); ends.append(e)
        S.barrier()
        t1 = ClockSampler.now()
        ms = S.max_over_ranks(max(e0.elapsed_time(e) for e in ends))[0]
        return dict(steps=n, ms_per_step=ms / n, seconds=ms * 1e-3, t0=t0, t1=t1)

    def e2e(self, steps, in_flight=1):
        """Host buffers in, RGBA8 frame out, through the C-ABI, pipelined over three streams: H2D uploads (rtb_upload on a copy
        context; each rank its 1/N slice of the primitive arrays, completed over NVLink by rtb_comm_all_gather on the same stream), the frame, and the
        D2H read-back of the resolved frame (rtb_download_async on a second copy context).  Two buffer sets alternate (in_flight > 1:
        one per lane, and the frames alternate over the lanes like in the timed leg)."""
        S, capi, torch, C = self.S, self.S.capi, self.S.torch, self.C
        fl = self.in_flight(in_flight)
        nset = max(2, len(fl))
        world, rank = S.world, S.rank
        up, up_st = S.aux_device("h2d", comm=True)
        dn, dn_st = S.aux_device("d2h")
        pin = lambda a: torch.from_numpy(a.copy()).pin_memory()  # noqa: E731
        hm, hmat = pin(self.host["models"]), pin(self.host["materials"])
        parts = {}
        for k in ("triangles", "spheres"):
            lo = min(rank * self.slice[k], self.host[k].size); hi = min(lo + self.slice[k], self.host[k].size)
            parts[k] = (lo, pin(self.host[k][lo:hi]) if hi > lo else None)
        sets = [dict(tris=torch.zeros_like(self.d_tris0), sphs=torch.zeros_like(self.d_sphs0), rgba8=torch.empty_like(self.rgba8),
                     models=torch.zeros_like(self.d_models), mats=torch.zeros_like(self.d_mats),
                     out=torch.empty((self.H, self.W, 4), dtype=torch.uint8).pin_memory(), done=None, read=None) for _ in range(nset)]
        h2d = hm.numel() + hmat.numel() + sum(p[1].numel() for p in parts.values() if p[1] is not None)
        d2h = sets[0]["out"].numel() if rank == 0 else 0

        def step(i):
            s = sets[i % nset]
            f = fl[i % len(fl)]
            st = f.S.stream
            if s["done"] is not None:
                up_st.wait_event(s["done"])                      # the frame that used this set two steps ago has finished
            capi.check(S.L.rtb_upload(up.handle, self.vp(s["models"]), C.c_void_p(hm.data_ptr()), hm.numel()))
            capi.check(S.L.rtb_upload(up.handle, self.vp(s["mats"]), C.c_void_p(hmat.data_ptr()), hmat.numel()))
            for k, buf in (("triangles", s["tris"]), ("spheres", s["sphs"])):
                lo, hp = parts[k]
                if hp is not None:
                    capi.check(S.L.rtb_upload(up.handle, C.c_void_p(buf.data_ptr() + lo), C.c_void_p(hp.data_ptr()), hp.numel()))
            if world > 1:                                        # the other ranks' slices arrive over NVLink, still on the upload stream
                up.comm_all_gather(s["tris"].data_ptr(), self.slice["triangles"])
                up.comm_all_gather(s["sphs"].data_ptr(), self.slice["spheres"])
            st.wait_event(up_st.record_event())
            if s["read"] is not None:
                st.wait_event(s["read"])                         # this set's RGBA8 frame has been read back
            f.flushed_frame(tris=s["tris"], sphs=s["sphs"], rgba8=s["rgba8"], models=s["models"], mats=s["mats"])
            s["done"] = st.record_event()
            if rank == 0:
                dn_st.wait_event(s["done"])
                capi.check(S.L.rtb_download_async(dn.handle, C.c_void_p(s["out"].data_ptr()), self.vp(s["rgba8"]), s["out"].numel()))
                s["read"] = dn_st.record_event()

        for i in range(nset):                                     # warm every buffer set
            step(i)
        S.barrier()
        w0 = time.perf_counter()
        for i in range(steps):
            step(i)
        up.wait_idle(); dn.wait_idle()
        S.barrier()
        wall_ms = (time.perf_counter() - w0) * 1e3
        wall_ms = S.max_over_ranks(wall_ms)[0]
        h2d_all, d2h_all = S.sum_over_ranks([h2d, d2h])
        self.e2e_frame = sets[(steps - 1) % nset]["out"].numpy().copy() if rank == 0 else None
        return dict(ms_per_step=wall_ms / steps, h2d=h2d_all, d2h=d2h_all)

    # -- roofline ------------------------------------------------------------------------------------------------------------
    def roofline(self, trace_ms, probe=True):
        S, capi = self.S, self.S.capi
        peak, peak_src = peaks()
        lw, lc = self.local_walk, self.local_cnt
        pix_local = len(self.layout.owned_rows(S.rank if self.tiled else (self.emu[0] if self.emu else 0))) * self.W
        ref_bytes = 40 * lc["nodeVisits"] + 40 * lc["triTests"] + 20 * lc["sphTests"] + 20 * lc["matReads"] + 32 * pix_local
        out = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "kernel_ms": trace_ms,
               "kernel": "rtb::trace_wave_kernel (+ rtb::trace_tail_kernel, pre-pass, accumulate: the launches of one rtb_raytrace call)"}
        if lw is None:
            out.update({"achieved": None, "frac": None, "traffic": None, "note": "A/B kernel: no walk counters; see reference_equivalent"})
        else:
            sharing = self.my_count > 1 and not (self.base_flags & capi.TRACE_NO_PRIMARY_SHARING)
            wb = capi.walk_bytes(lw, sharing)
            wb["coalesced_record_bytes"] = int(64 * lw["uniqueRecordFetches"])     # lanes of a warp step on the same record share the fetch
            # the accumulate pass: one 16-byte slot read per (pixel, sample) + one 16-byte image read and write per active pixel;
            # the pre-pass: a 16-byte slot write per (pixel, sample) + the image read
            active = lw["items"] // (self.my_count + 1) if sharing else lw["items"] // max(self.my_count, 1)
            wb["accumulate_bytes"] = int(16 * lw["paths"] + 32 * active)
            wb["prepass_bytes"] = int(16 * lw["paths"] + 16 * pix_local + 8 * active + (48 * active if sharing else 0))
            total = wb["load_bytes"] + wb["store_bytes"] + wb["accumulate_bytes"] + wb["prepass_bytes"]
            achieved = total / (trace_ms * 1e-3) / 1e9
            fetch = (wb["record_bytes"] + wb["leaf_box_bytes"] + wb["primitive_bytes"]) / (trace_ms * 1e-3) / 1e9
            out.update({"achieved": achieved, "frac": achieved / peak, "algorithmic_bytes_per_launch": int(total), "bytes": wb,
                        "walk_counters": {k: v for k, v in lw.items() if not k.startswith("_")},
                        "per_ray": {"record_fetches": lw["recordFetches"] / max(lw["rays"], 1), "leaf_box_fetches": lw["leafBoxFetches"] / max(lw["rays"], 1),
                                    "primitive_tests": (lw["triTests"] + lw["sphTests"]) / max(lw["rays"], 1)},
                        "lanes_per_traverse_step": lw["laneSteps"] / max(lw["warpSteps"], 1)})
            footprint = 64 * max(self.T + self.Sn - 1, 1) + 32 * (self.T + self.Sn) + 64 * self.T + 20 * self.Sn
            if probe:
                g, g2 = self.C.c_float(), self.C.c_float()
                capi.check(S.L.rtb_probe_gather(S.h, 32 << 20, self.C.byref(g)))
                capi.check(S.L.rtb_probe_gather(S.h, footprint, self.C.byref(g2)))
                out["node_fetch"] = {"bound": "l2: divergent 64-byte record fetches served by L2 (rtb_probe_gather, measured in this run: independent "
                                              "random fetches from every lane of a trace-shaped grid over an L2-resident 32 MiB buffer)",
                                     "peak": float(g.value), "achieved": fetch, "unit": "GB/s", "frac": fetch / float(g.value) if g.value > 0 else None,
                                     "footprint_bytes": int(footprint), "uniform_random_at_footprint": float(g2.value),
                                     "note": "achieved = record + leaf-box + primitive bytes the lanes request / trace time; requests of a warp for the same "
                                             "record coalesce in L1 (see bytes.coalesced_record_bytes) and part of them hit L1, so this is the demand on the "
                                             "fetch path, not L2 traffic; uniform_random_at_footprint is the same probe over a buffer of this scene's record "
                                             "footprint -- a traversal has locality (tree tops stay cached) and beats it when the footprint exceeds L2"}
            tp = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
            traffic, lim = None, None
            if os.path.exists(tp):
                try:
                    prof = json.load(open(tp))
                    key = self.name + ("x" if self.ext else "")
                    ent = prof.get(key)
                    if ent and S.world == 1 and not self.emu:
                        traffic, lim = ent.get("dram_bytes_per_launch"), ent.get("limiter")
                        out["traffic_source"] = ent.get("source")
                        # ncu's count of distinct 32-byte sectors the global loads of the trace launches requested, beside this run's
                        # own estimate of the same quantity: coalesced records + the other loads (agreement: the counters count what runs)
                        out["l1tex_global_load_sector_bytes_ncu"] = ent.get("l1tex_global_load_sector_bytes")
                        out["l1tex_global_load_sector_bytes_counted"] = int(wb["load_bytes"] - wb["record_bytes"] + wb["coalesced_record_bytes"])
                except Exception:  # noqa: BLE001
                    pass
            out["traffic"] = traffic
            out["limiter_from_ncu"] = lim
        out["reference_equivalent"] = {
            "bytes_per_launch": int(ref_bytes), "gbs": ref_bytes / (trace_ms * 1e-3) / 1e9,
            "note": "SURVEY.md 8d formula 40 V + 40 Tt + 20 St + 20 H + 32 B/pixel on the REFERENCE's work counters (reference-order walk, one traversal "
                    "per sample): the rate at which the reference's algorithmic bytes are retired, not bytes this kernel moves -- it may exceed the HBM peak",
            "per_ray": {"node_visits": lc["nodeVisits"] / max(lc["rays"], 1), "tri_tests": lc["triTests"] / max(lc["rays"], 1),
                        "sphere_tests": lc["sphTests"] / max(lc["rays"], 1)}}
        return out

    def close(self):
        for t in getattr(self, "_twins", [self])[1:]:
            t.close()
        for k in list(self.__dict__):
            if k.startswith("d_") or k in ("image", "final", "rgba8"):
                delattr(self, k)
        self.S.torch.cuda.empty_cache()


def compact(S, args, name, ext, shard):
    """breakdown.configs entry: one other BASELINE config, measured the same way over a short timed region"""
    f = Frame(S, args, name, ext=ext, shard=shard)
    chk = f.count_and_check()
    f.check_assembled()
    S.flush.zero_(); e0, e1 = S.ev(), S.ev()
    e0.record(S.stream); f.frame(); e1.record(S.stream)
    S.torch.cuda.synchronize()
    first = S.max_over_ranks(e0.elapsed_time(e1))[0]
    steps = int(min(50, max(2, 1500.0 / max(first, 1e-3))))
    t = f.timed(steps, 1)
    ms = t["ms_per_step"]
    rf = f.roofline(t["trace_ms"], probe=True)
    out = {"workload": f.desc["workload"], "mode": "extension (metal / dielectric scatter)" if ext else "reference-parity",
           "n_gpus": S.world, "sharding": ("samples" if f.by_samples else "tiles") if S.world > 1 else "single GPU",
           "ms_per_step": ms, "steps": steps, "value": f.rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s (reference-equivalent)",
           "mrays_traversed_per_s": f.rays_traversed / (ms * 1e-3) / 1e6, "rays_per_step": f.rays, "rays_traversed_per_step": f.rays_traversed,
           "trace_ms": t["trace_ms"], "gpu_launches": t["launches"], "frame_check": chk, "rgba8_sha256": f.frame_sha() if S.rank == 0 else None,
           "roofline": {k: rf.get(k) for k in ("achieved", "frac", "algorithmic_bytes_per_launch", "traffic", "node_fetch", "per_ray", "lanes_per_traverse_step")},
           "reference_equivalent_per_ray": rf["reference_equivalent"]["per_ray"]}
    f.close()
    return out


def run_b200(args):
    S = Session()
    sampler = ClockSampler(S.local_rank)     # started early: nvidia-smi needs a few hundred ms before its first sample
    sampler.start()
    world, rank = S.world, S.rank
    f = Frame(S, args, args.config, ext=args.ext, shard=args.shard, main=True)
    chk = f.count_and_check()                # the counters are needed in any case; the comparison costs nothing extra
    if not args.no_frame_check:
        f.check_assembled()
    t_load = ClockSampler.now()
    fif = max(1, args.frames_in_flight)
    serial = f.timed(args.steps if fif == 1 else min(args.steps, 10), args.warmup)      # one frame after the other: frame latency + the trace launches' duration
    t = serial if fif == 1 else f.timed(args.steps, args.warmup, in_flight=fif)
    ms_per_step, trace_ms = t["ms_per_step"], serial["trace_ms"]
    value = f.rays / (ms_per_step * 1e-3) / 1e6
    S.torch.cuda.synchronize()
    sha_timed = f.frame_sha() if rank == 0 else None
    if fif > 1 and rank == 0:            # every lane renders the same frame
        shas = {hashlib.sha256(x.rgba8.cpu().numpy().tobytes()).hexdigest() for x in f.in_flight(fif)}
        chk["frames_in_flight"] = "the RGBA8 frames of all lanes are identical" if shas == {sha_timed} else "MISMATCH between lanes"
        if shas != {sha_timed}:
            chk["status"] = "MISMATCH"

    e = f.e2e(args.steps, in_flight=fif)
    e2e_value = f.rays / (e["ms_per_step"] * 1e-3) / 1e6
    if rank == 0 and f.e2e_frame is not None:
        chk["e2e_frame"] = ("RGBA8 frame read back by the e2e leg == the timed leg's frame" if hashlib.sha256(f.e2e_frame.tobytes()).hexdigest() == sha_timed
                            else "MISMATCH: e2e frame differs")
        if chk["e2e_frame"].startswith("MISMATCH"):
            chk["status"] = "MISMATCH"

    sus = f.sustained(args.min_seconds, in_flight=fif) if args.min_seconds > 0 else None

    # build-only timing (reported, explains the step)
    b0, b1 = S.ev(), S.ev()
    f.d_tris.copy_(f.d_tris0); f.d_sphs.copy_(f.d_sphs0)
    b0.record(S.stream); f.build(f.d_tris, f.d_sphs); b1.record(S.stream)
    S.torch.cuda.synchronize()
    build_ms = b0.elapsed_time(b1)
    roofline = f.roofline(trace_ms)

    names = []
    if args.breakdown == "auto":
        if args.emulate_rank:
            names = []
        elif world == 1:
            names = [n for n in ("C1", "C2", "C3", "C3x", "C4", "C5", "C5x") if n != args.config + ("x" if args.ext else "")]
        else:
            names = ["C4", "C5s"]
    elif args.breakdown != "none":
        names = [n for n in args.breakdown.split(",") if n]
    configs = {}
    for n in names:
        base = n.rstrip("xs")
        shard = "samples" if n.endswith("s") else "tiles"
        try:
            configs[n] = compact(S, args, base, ext=n.endswith("x"), shard=shard)
        except Exception as ex:  # noqa: BLE001
            configs[n] = {"error": f"{type(ex).__name__}: {ex}"}

    sampler.stop()
    clocks = sampler.window(t["t0"], t["t1"], fallback_from=t_load)
    if sus:
        sus_clocks = sampler.window(sus["t0"], sus["t1"])

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            s = cpu_oracle_sample(f.sc, f.ubo, f.W, f.H, args.cpu_seconds, ext=args.ext)
            cpu = {"value": s["mrays"], "unit": "Mrays/s", "cores": s["cores"], "kind": "port",
                   "sample": f"full {f.W}x{f.H} frame, {s['spp']} of {f.spp} spp ({s['rays']} rays in {s['t_trace']:.1f} s trace; BVH build "
                             f"{s['t_build']:.2f} s not included); CPU oracle = C restatement of the reference shaders, OpenMP over rows on all "
                             f"{s['cores']} host cores; lavapipe unavailable in image"}
        desc = f.desc
        desc["parallelism"] = ("single GPU" if world == 1 else
                               f"samples{world}: sample ranges of {f.spp // world} spp per rank, scene replicated, rtb_reduce_samples (NCCL sum-reduce + fused resolve)" if f.by_samples else
                               f"tile{world}: 8-row bands interleaved over {world} rank(s), scene replicated, rtb_gather_tiles (resolve -> NCCL all-gather of RGBA8 -> re-assembly)")
        desc["l2"] = "256 MiB buffer written between timed steps (L2 flush)"
        desc["frames_in_flight"] = (f"{fif}: consecutive frames alternate over {fif} library contexts (own stream, BVH and frame buffers), so a frame's BVH build "
                                    "overlaps the previous frame's draining trace launch; ms_per_step = the K frames' bracket / K, frame_latency_ms = one frame alone"
                                    if fif > 1 else "1: one frame after the other")
        desc["traversal"] = ("reference visiting order, no t-interval (--reference-order)" if args.reference_order or args.kernel != "wave" else
                             "library default: 4-ary records walked nearest-first with t-culling (scenes of >= 512 primitives), else the exact child pairs in the reference's order")
        desc["primary_sharing"] = ("off" if (args.no_primary_sharing or args.kernel != "wave") else
                                   "on: the samples of a pixel share one traversal of their identical (un-jittered) primary ray")
        if args.mode == "culled":
            desc["mode"] = "EXTENSION RTB_TRACE_CULLED: segment-box culling on top of the reference traversal (not the headline mode)"
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "value_counts": VALUE_COUNTS,
            "mrays_traversed_per_s": f.rays_traversed / (ms_per_step * 1e-3) / 1e6,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "frame_latency_ms": serial["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": desc, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(e["h2d"]), "d2h_bytes_per_step": int(e["d2h"]),
                    "ms_per_step": e["ms_per_step"],
                    "how": "wall clock over K pipelined steps (upload / frame / read-back on three streams, two buffer sets), max over ranks"},
            "gpu_launches": int(t["launches"]),
            "frame_check": chk, "rgba8_sha256": sha_timed,
            "sustained": None if not sus else {"value": f.rays / (sus["ms_per_step"] * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": sus["ms_per_step"],
                                               "steps": sus["steps"], "seconds": sus["seconds"], "clocks": sus_clocks,
                                               "mrays_traversed_per_s": f.rays_traversed / (sus["ms_per_step"] * 1e-3) / 1e6},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "breakdown": {"rays_per_step": f.rays, "rays_traversed_per_step": f.rays_traversed,
                          "samples_per_step": f.cnt["samples"], "msamples_per_s": f.cnt["samples"] / (ms_per_step * 1e-3) / 1e6,
                          "bvh_build_ms": build_ms, "trace_ms": trace_ms, "serial_ms_per_step": serial["ms_per_step"], "counters": f.cnt, "configs": configs},
        }
        print(json.dumps(line), flush=True)
    f.close()
    S.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
