#!/usr/bin/env python
"""bench.py -- Mrays/s of the path-tracer hot path (S1 BVH build + S2 trace + resolve) on N B200s.

    python bench.py --gpus N --steps K --warmup W [--config C2] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

One "step" = one frame of the named config, exactly what RaytracerBVHRenderer::Raytracer::doIteration does per
frame (RaytracerBVH.hpp:206-496): fresh model-space scene arrays -> clear -> K1..K6 BVH build -> all samples of the
frame -> resolve to RGBA8.  N > 1: the frame's 8-row bands are interleaved over the ranks (scene replicated, BVH built
per rank, no data-path collective while tracing), then ONE NCCL all-gather assembles the image (strong scaling).

`value`  : rays (hitBVH calls, counted on the device by the instrumented variant of the same kernel) / device time of
           the step with the scene arrays already resident in HBM.
`e2e`    : the same frame through the reference-facing C-ABI with HOST buffers: pinned host scene arrays are uploaded
           and the resolved RGBA8 frame is read back inside the timed region.
`roofline`: trace kernel only: algorithmic bytes 40*V + 40*Tt + 20*St + 20*H (+32 B per pixel for the accumulator)
           over its CUDA-event duration, against the measured HBM copy bandwidth (DESIGN.md "roofline").
`--impl reference`: the reference's own algorithm on the host cores = the CPU oracle (the reference's GLSL cannot run
           here: no Vulkan ICD / lavapipe in the image), timed on a bounded sample of the same workload.
"""
import argparse
import datetime
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
METRIC = "Mrays/s per scene (path segments = hitBVH calls per second)"


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
    ap.add_argument("--kernel", default="wave", choices=["wave", "simple", "stream"], help="trace kernel (A/B switch)")
    ap.add_argument("--reference-order", action="store_true",
                    help="walk the tree in the reference's visiting order with no t-interval, like the shader (A/B switch, same results)")
    ap.add_argument("--no-primary-sharing", action="store_true",
                    help="trace every sample's (identical, un-jittered) primary ray separately like the shader does (A/B switch, same results)")
    ap.add_argument("--nodes", default="auto", choices=["auto", "exact", "compressed", "wide"],
                    help="traversal records: auto (library default: 4-ary for >= 8192 primitives), exact 64-byte child pairs, "
                         "32-byte compressed, 64-byte 4-ary (A/B switch, same results)")
    ap.add_argument("--emulate-rank", default="", help="debugging: 'r/n' renders only the bands rank r of n would own, on one GPU, "
                                                         "without the collective (the per-rank workload of a tile-mode run, e.g. for ncu)")
    ap.add_argument("--shard", default="tiles", choices=["tiles", "samples"],
                    help="N > 1: tiles = 8-row bands + all-gather (bit-identical, default); samples = sample ranges + sum-reduce (C5)")
    ap.add_argument("--mode", default="exact", choices=["exact", "culled"],
                    help="exact = the reference's traversal (default, the headline); culled = extension RTB_TRACE_CULLED")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload(cfg_name, spp_override=0):
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
        "mode": "reference-parity (metal/dielectric absorb like the shader; frame bit-identical to the oracle)",
    }
    return cfg, sc, ubo, spp, desc


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 20 ms from well before the warm-up (it needs a few hundred ms to start); only the samples whose
    own timestamp falls inside the timed region (+- one period) are reported, falling back to the warm-up + timed span when a
    very short timed region caught none."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t_load = self.t0 = self.t1 = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def mark_load(self):      # the GPU is under the benchmark's load from here (warm-up)
        self.t_load = datetime.datetime.now()

    def mark_begin(self):
        self.t0 = datetime.datetime.now()

    def mark_end(self):
        self.t1 = datetime.datetime.now()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(parts[2]), float(parts[3]), [n for n, v in zip(names, parts[6:10]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        pad = datetime.timedelta(milliseconds=20)
        window, scope = [r for r in rows if self.t0 and self.t1 and self.t0 - pad <= r[0] <= self.t1 + pad], "timed region"
        if not window:
            window, scope = [r for r in rows if self.t_load and self.t1 and self.t_load <= r[0] <= self.t1 + pad], "warm-up + timed region"
        if not window and self.t0 and self.t1:      # region shorter than the sampling period: the samples right around it
            near = datetime.timedelta(milliseconds=500)
            window, scope = [r for r in rows if self.t0 - near <= r[0] <= self.t1 + near], "within 0.5 s of the timed region (shorter than the sampling period)"
        if not window:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}
        reasons = sorted({n for r in window for n in r[3]})
        return {"sm_mhz": float(np.median([r[1] for r in window])), "sm_max_mhz": float(max(r[2] for r in window)), "reasons": reasons,
                "samples": len(window), "scope": scope}


# ----------------------------------------------------------------------------------------------------------------
# CPU oracle legs (the ONLY place bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------------------------
def cpu_oracle_sample(sc, ubo, W, H, seconds, max_spp=64):
    """Times the oracle (C restatement of the reference shaders, OpenMP over rows) on a bounded sample of the workload:
    the full frame at 1 spp, repeated (as further samples of the same frame) until `seconds` is used up."""
    from oracle import oracle as O
    t0 = time.time()
    b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    t_build = time.time() - t0
    img = None
    rays = 0; t_trace = 0.0; spp = 0
    while spp < max_spp and (spp == 0 or t_trace + t_trace / spp <= seconds):
        t0 = time.time()
        r = O.raytrace(ubo, W, H, b["tris"], b["sphs"], sc["materials"], b["nodes"], 1, image=img, want_hits=False, want_rng=False)
        t_trace += time.time() - t0
        img = r["image"]; rays += r["counters"]["rays"]; spp += 1
    return dict(mrays=rays / t_trace / 1e6, rays=rays, t_trace=t_trace, t_build=t_build, spp=spp, cores=O.max_threads())


def run_reference(args):
    """--impl reference: the reference's own CPU-runnable implementation of the path.  Its GLSL cannot execute in this
    image (no Vulkan loader / lavapipe ICD), so this is the oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, sc, ubo, spp, desc = workload(args.config, args.spp)
    from oracle import oracle as O
    W, H = cfg["width"], cfg["height"]
    b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
    times, rays = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        # one step = a bounded sample of the frame: BVH build + 1 of the `spp` samples per pixel
        b = O.build_bvh(sc["models"], sc["triangles"], sc["spheres"])
        r = O.raytrace(ubo, W, H, b["tris"], b["sphs"], sc["materials"], b["nodes"], 1, want_hits=False, want_rng=False)
        dt = time.time() - t0
        if i >= args.warmup:
            times.append(dt); rays += r["counters"]["rays"]
    total = sum(times)
    value = rays / total / 1e6
    sample = f"full {W}x{H} frame, BVH build + 1 of {spp} spp per step (rate)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(len(times), 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": desc,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": O.max_threads(), "kind": "port", "sample": sample,
                         "note": "CPU oracle = C restatement of the reference shaders; lavapipe/Vulkan unavailable in image"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from raytracergpu_mastersproject_b200 import Device, capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout; keep stdout to the one JSON line the contract asks for
        # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, so VERSION is raised to WARN)
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join(tempfile.gettempdir(), "rtb200_nccl_%h_%p.log"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    tdev = torch.device("cuda", local_rank)
    # a non-default stream shared by torch (copies, NCCL ordering, events) and librtb200 (all kernels): the default
    # stream's handle is NULL, which rtb_ctx_create reads as "create a private stream"
    stream = torch.cuda.Stream(device=tdev)
    torch.cuda.set_stream(stream)

    cfg, sc, ubo, spp, desc = workload(args.config, args.spp)
    W, H = cfg["width"], cfg["height"]
    T, S, M = len(sc["triangles"]), len(sc["spheres"]), len(sc["materials"])
    L = capi.lib()
    dev = Device(local_rank, stream=stream.cuda_stream)
    h = dev.handle
    import ctypes as C
    ubo_p = ubo.ctypes.data_as(C.c_void_p)

    # band sharding of the frame (bit-identical to 1 GPU, tests/test_gpu_parity.py::test_tile_sharding_bit_identical)
    from raytracergpu_mastersproject_b200.sharding import BandLayout, assemble_gathered, single_gpu_layout
    by_samples = world > 1 and args.shard == "samples"
    emu = tuple(int(x) for x in args.emulate_rank.split("/")) if args.emulate_rank and world == 1 else None
    layout = (BandLayout(H, world, BAND_ROWS) if (world > 1 and not by_samples) else
              BandLayout(H, emu[1], BAND_ROWS) if emu else single_gpu_layout(H))
    rows, band_rows = layout.local_rows, layout.band_rows
    from raytracergpu_mastersproject_b200.sharding import sample_range
    my_first, my_count = sample_range(spp, world, rank) if by_samples else (0, spp)

    # resident inputs: pristine model-space arrays + working copies (K1 transforms in place, so every frame starts
    # from the model-space data -- the reference re-uploads it, RaytraceScene.cpp:78-113)
    def to_dev(a):
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy())
        return t.to(tdev)
    d_models = to_dev(sc["models"]); d_mats = to_dev(sc["materials"])
    d_tris0 = to_dev(sc["triangles"]) if T else torch.zeros(64, dtype=torch.uint8, device=tdev)
    d_sphs0 = to_dev(sc["spheres"]) if S else torch.zeros(32, dtype=torch.uint8, device=tdev)
    d_tris = torch.empty_like(d_tris0); d_sphs = torch.empty_like(d_sphs0)
    image = torch.empty((rows, W, 4), dtype=torch.float32, device=tdev)
    gathered = torch.empty((world, rows, W, 4), dtype=torch.float32, device=tdev) if (world > 1 and not by_samples) else None
    final = torch.empty((H, W, 4), dtype=torch.float32, device=tdev) if world > 1 else image
    rgba8 = torch.empty((H, W, 4), dtype=torch.uint8, device=tdev)
    counters = torch.zeros(6, dtype=torch.int64, device=tdev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=tdev)      # > 126 MB L2

    targs = capi.TraceArgs()
    targs.imageWidth, targs.imageHeight, targs.localRows = W, H, rows
    tiled = world > 1 and not by_samples
    targs.bandRows, targs.bandFirst, targs.bandStep = band_rows, (rank if tiled else 0), (world if tiled else 1)
    if emu:
        targs.bandFirst, targs.bandStep = emu[0], emu[1]
    targs.sampleSkip, targs.sampleCount, targs.flags = my_first, my_count, 0

    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def frame(trace_events=None, count=False):
        d_tris.copy_(d_tris0); d_sphs.copy_(d_sphs0)
        capi.check(L.rtb_clear_image(h, vp(image), W, rows))
        capi.check(L.rtb_build_bvh(h, ubo_p, vp(d_models), vp(d_tris), vp(d_sphs), vp(d_mats), None, None, None, None, None, 0))
        targs.flags = ((capi.TRACE_COUNT if count else 0) | {"simple": capi.TRACE_SIMPLE_KERNEL, "stream": capi.TRACE_STREAM_KERNEL}.get(args.kernel, 0)
                       | (capi.TRACE_CULLED if args.mode == "culled" else 0) | (capi.TRACE_NO_PRIMARY_SHARING if args.no_primary_sharing else 0) | (capi.TRACE_REFERENCE_ORDER if args.reference_order else 0) | {"compressed": capi.TRACE_COMPRESSED_NODES, "wide": capi.TRACE_WIDE_NODES, "exact": capi.TRACE_EXACT_NODES}.get(args.nodes, 0))
        targs.counters = counters.data_ptr() if count else None
        if trace_events:
            trace_events[0].record(stream)
        capi.check(L.rtb_raytrace(h, ubo_p, vp(image), C.byref(targs)))
        if trace_events:
            trace_events[1].record(stream)
        if by_samples:
            # every rank holds a full-frame partial sum of its sample range: sum the rgb planes onto rank 0 (fp32
            # re-association -> tolerance, not bit-equality); alpha = end of the seed chain = the last rank's
            final.copy_(image)
            dist.reduce(final, dst=0, op=dist.ReduceOp.SUM)
        elif world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), image.view(-1))
            final.copy_(assemble_gathered(gathered, layout))      # rank r, local band b -> global band b*world + r
        capi.check(L.rtb_resolve_rgba8(h, vp(final), W, rows if emu else H, spp, vp(rgba8)))

    sampler = ClockSampler(local_rank)     # started early: nvidia-smi needs a few hundred ms before its first sample
    sampler.start()
    # ---- work counters (deterministic; one instrumented, untimed frame) ----
    counters.zero_()
    frame(count=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.all_reduce(counters)
    cnt = dict(zip(capi.COUNTER_FIELDS, [int(x) for x in counters.tolist()]))
    local_cnt = cnt
    rays = cnt["rays"]
    # Rays the kernel really walks: the reference's camera rays are not jittered, so the wave kernel traces each active pixel's
    # primary ray once per submission and its samples start from that hit (RTB_TRACE_NO_PRIMARY_SHARING switches it off).
    # Active pixels = rays of an instrumented 1-sample, depth-1 submission.
    sharing = args.kernel == "wave" and not args.no_primary_sharing and my_count > 1
    u1 = ubo.copy(); u1["maxRayTraceDepth"] = 1
    counters.zero_()
    targs.sampleSkip, targs.sampleCount, targs.flags, targs.counters = 0, 1, capi.TRACE_COUNT, counters.data_ptr()
    capi.check(L.rtb_raytrace(h, u1.ctypes.data_as(C.c_void_p), vp(image), C.byref(targs)))
    torch.cuda.synchronize()
    active_px = int(counters[0])
    targs.sampleSkip, targs.sampleCount, targs.flags, targs.counters = my_first, my_count, 0, None
    saved = torch.tensor([active_px * (my_count - 1) if sharing else 0], dtype=torch.int64, device=tdev)
    if world > 1:
        dist.all_reduce(saved)
    rays_traversed = rays - int(saved[0])

    # ---- warm-up ----
    sampler.mark_load()
    for _ in range(max(args.warmup, 0)):
        flush.zero_()
        frame()
    torch.cuda.synchronize()

    # ---- timed: exactly K steps, L2 flushed between steps, device-timed ----
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_begin()
    launches0 = dev.launch_count()
    step_ev, trace_ev = [], []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1, t0, t1 = ev(), ev(), ev(), ev()
        e0.record(stream)
        frame(trace_events=(t0, t1))
        e1.record(stream)
        step_ev.append((e0, e1)); trace_ev.append((t0, t1))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_end()
    launches = dev.launch_count() - launches0
    clocks = sampler.stop()
    step_ms = sum(a.elapsed_time(b) for a, b in step_ev)
    trace_ms = sum(a.elapsed_time(b) for a, b in trace_ev) / args.steps
    if world > 1:
        t = torch.tensor([step_ms, trace_ms], dtype=torch.float64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, trace_ms = float(t[0]), float(t[1])
    ms_per_step = step_ms / args.steps
    value = rays / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host buffers in, RGBA8 frame out, through the C-ABI (N = 1: whole frame; N > 1: each rank uploads
    # its replica and rank 0 reads the assembled frame) ----
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).pin_memory()  # noqa: E731
    h_models, h_mats = pin(sc["models"]), pin(sc["materials"])
    h_tris = pin(sc["triangles"]) if T else None
    h_sphs = pin(sc["spheres"]) if S else None
    h_out = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    h2d = h_models.numel() + h_mats.numel() + (h_tris.numel() if T else 0) + (h_sphs.numel() if S else 0)
    d2h = h_out.numel()

    def frame_e2e():
        capi.check(L.rtb_upload(h, vp(d_models), C.c_void_p(h_models.data_ptr()), h_models.numel()))
        capi.check(L.rtb_upload(h, vp(d_mats), C.c_void_p(h_mats.data_ptr()), h_mats.numel()))
        if T:
            capi.check(L.rtb_upload(h, vp(d_tris0), C.c_void_p(h_tris.data_ptr()), h_tris.numel()))
        if S:
            capi.check(L.rtb_upload(h, vp(d_sphs0), C.c_void_p(h_sphs.data_ptr()), h_sphs.numel()))
        frame()
        if rank == 0:
            capi.check(L.rtb_download(h, C.c_void_p(h_out.data_ptr()), vp(rgba8), h_out.numel()))   # synchronises

    frame_e2e()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = ev(), ev()
    e0.record(stream)
    for _ in range(args.steps):
        flush.zero_()
        frame_e2e()
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e2e_ms, wall_ms) if world == 1 else e2e_ms
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    e2e_value = rays / (e2e_ms / args.steps * 1e-3) / 1e6

    # ---- build-only timing (reported, explains the step) ----
    b0, b1 = ev(), ev()
    d_tris.copy_(d_tris0); d_sphs.copy_(d_sphs0)
    b0.record(stream)
    capi.check(L.rtb_build_bvh(h, ubo_p, vp(d_models), vp(d_tris), vp(d_sphs), vp(d_mats), None, None, None, None, None, 0))
    b1.record(stream)
    torch.cuda.synchronize()
    build_ms = b0.elapsed_time(b1)

    # ---- roofline of the dominant kernel (trace) ----
    peak, peak_src = peaks()
    pix_local = len(layout.owned_rows(rank if tiled else 0)) * W
    # per launch (this rank): counters of this rank's launch; at N = 1 the all-reduced counters are this rank's
    if world > 1:
        lc = torch.zeros(6, dtype=torch.int64, device=tdev)
        counters.zero_(); frame(count=True); torch.cuda.synchronize()
        lc.copy_(counters)
        local_cnt = dict(zip(capi.COUNTER_FIELDS, [int(x) for x in lc.tolist()]))
    alg_bytes = 40 * local_cnt["nodeVisits"] + 40 * local_cnt["triTests"] + 20 * local_cnt["sphTests"] + 20 * local_cnt["matReads"] + 32 * pix_local
    achieved = alg_bytes / (trace_ms * 1e-3) / 1e9
    traffic, limiter = None, None
    tp = os.path.join(ROOT, "profiles", "trace_kernel_dram_bytes.json")
    if os.path.exists(tp):
        try:
            prof = json.load(open(tp))
            traffic = prof.get(args.config)
            limiter = prof.get("_limiter", {}).get(args.config)     # what ncu says bounds the kernel (not HBM): reported, not measured live
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": {"wave": "rtb::trace_wave_kernel", "stream": "rtb::trace_stream_kernel"}.get(args.kernel, "rtb::trace_kernel"), "kernel_ms": trace_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "peak_source": peak_src, "limiter_from_ncu": limiter,
                "per_ray": {"node_visits": local_cnt["nodeVisits"] / max(local_cnt["rays"], 1),
                            "tri_tests": local_cnt["triTests"] / max(local_cnt["rays"], 1),
                            "sphere_tests": local_cnt["sphTests"] / max(local_cnt["rays"], 1)}}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            s = cpu_oracle_sample(sc, ubo, W, H, args.cpu_seconds)
            cpu = {"value": s["mrays"], "unit": "Mrays/s", "cores": s["cores"], "kind": "port",
                   "sample": f"full {W}x{H} frame, {s['spp']} of {spp} spp ({s['rays']} rays in {s['t_trace']:.1f} s trace; BVH build "
                             f"{s['t_build']:.2f} s not included); CPU oracle = C restatement of the reference shaders, OpenMP over rows; "
                             f"lavapipe unavailable in image"}
        desc["parallelism"] = ("single GPU" if world == 1 else
                               f"samples{world}: sample ranges of {spp // world} spp per rank, scene replicated, NCCL sum-reduce" if by_samples else
                               f"tile{world}: 8-row bands interleaved over {world} rank(s), scene replicated, NCCL all-gather")
        desc["l2"] = "256 MiB buffer written between timed steps (L2 flush)"
        desc["traversal"] = ("reference visiting order, no t-interval (--reference-order)" if args.reference_order or args.kernel != "wave" else
                             "library default: 4-ary records walked nearest-first with t-culling when the scene qualifies, else the reference's order")
        desc["primary_sharing"] = ("on: the samples of a pixel share one traversal of their identical (un-jittered) primary ray; value counts "
                                   "the reference's rays, breakdown.rays_traversed_per_step the rays walked" if sharing else "off")
        if args.mode == "culled":
            desc["mode"] = "EXTENSION RTB_TRACE_CULLED: segment-box culling on top of the reference traversal (not the headline mode)"
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": desc, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "breakdown": {"rays_per_step": rays, "rays_traversed_per_step": rays_traversed,
                          "mrays_traversed_per_s": rays_traversed / (ms_per_step * 1e-3) / 1e6, "samples_per_step": cnt["samples"], "msamples_per_s": cnt["samples"] / (ms_per_step * 1e-3) / 1e6,
                          "bvh_build_ms": build_ms, "trace_ms": trace_ms, "counters": cnt},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    dev.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
