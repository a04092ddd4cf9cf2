"""Turns an .ncu-rep (ncu --set full --import-source on) into the text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/wave3_r01.ncu-rep profiles/r01_wave_ncu_summary.txt "title" [launch index in the report]

Reads the raw page (kernel-level counters) and the source page (per-SASS-instruction execution counts) through
`ncu -i ... --page raw|source --csv` and prints: duration, DRAM / L2 / L1 traffic and hit rates, issue utilisation, pipe
utilisation, stall reasons, lanes per instruction, and the hot SASS regions with their lane efficiency.
"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "sm__sass_thread_inst_executed_op_fadd_pred_on.sum", "sm__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
]


SKIP = 0


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv", "-s", str(SKIP), "-c", "1"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    global SKIP
    rep, dst, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    SKIP = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    lines = [f"# {title}", f"# source: {rep} (ncu --set full --clock-control none --import-source on)", ""]
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    lines.append(f"kernel: {kname}")
    table = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    for k in KEYS:
        if k in table:
            lines.append(f"{k:75s} {table[k][1]:>18s} {table[k][0]}")
    lines.append("")
    lines.append("stall reasons (warps per issue-active cycle):")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            v = float(table[h][1] or 0)
            if v >= 0.05:
                lines.append(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:6.2f}")
    src = page(rep, "source")
    second = [i for i, r in enumerate(src) if i > 0 and r and r[0] == "Kernel Name"]
    if second:                          # with a launch filter ncu prints the kernel's table twice
        src = src[:second[0]]
    sh = src[1]
    ia, isrc, ie, it, isamp = sh.index("Address"), sh.index("Source"), sh.index("Instructions Executed"), sh.index("Thread Instructions Executed"), sh.index("# Samples")
    data = []
    for r in src[2:]:
        try:
            data.append((r[isrc].strip(), int(r[ie]), int(r[it]), int(r[isamp])))
        except (ValueError, IndexError):
            pass
    tot = sum(d[1] for d in data); tott = sum(d[2] for d in data); tots = max(sum(d[3] for d in data), 1)
    lines += ["", f"SASS: {len(data)} instructions, {tot} warp-instructions executed, {tott / max(tot, 1):.2f} active lanes per instruction",
              "hot regions (contiguous SASS with similar execution counts; share of warp-instructions, of stall samples, lanes):"]
    seg, cur = [], None
    for i, d in enumerate(data):
        if cur is None or abs(d[1] - cur["e"]) > 0.15 * max(cur["e"], 1):
            if cur:
                seg.append(cur)
            cur = dict(start=i, e=d[1], n=0, inst=0, tinst=0, samp=0, ops=[])
        cur["n"] += 1; cur["inst"] += d[1]; cur["tinst"] += d[2]; cur["samp"] += d[3]
        cur["ops"].append(d[0].split()[0] if d[0] else "")
    if cur:
        seg.append(cur)
    for s in seg:
        if s["inst"] < 0.01 * tot:
            continue
        c = Counter(o.split(".")[0] for o in s["ops"] if o.startswith(("LDG", "MUFU", "STS", "LDS", "STG", "ATOM", "VOTE", "SHFL", "CALL", "RET", "LDL", "STL")))
        lines.append(f"  sass[{s['start']:4d}..{s['start'] + s['n']:4d}) exec/inst {s['e']:>11d}  share {100 * s['inst'] / tot:5.1f}%  samples {100 * s['samp'] / tots:5.1f}%"
                     f"  lanes {s['tinst'] / max(s['inst'], 1):5.1f}  {dict(c)}")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
