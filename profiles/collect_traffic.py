"""Per-config DRAM traffic of the production trace launches, from an ncu metrics pass, into profiles/r02_dram_traffic.json
(the `roofline.traffic` source of bench.py).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum \\
        --clock-control none -k regex:trace_ --csv --log-file gpurun_out/traffic_C2.csv python bench.py --config C2 --steps 1 --warmup 1 ...
    python profiles/collect_traffic.py gpurun_out/traffic_C2.csv C2

One frame of the production path = [primary-hit launch of trace_wave_kernel, trace_tail_kernel,] then per pass the main launch of
trace_wave_kernel + trace_tail_kernel; the launches of the LAST frame of un-instrumented kernels (template argument COUNT = 0) are summed
("per launch" = per rtb_raytrace call, like bench.py's roofline.achieved).
"""
import csv
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    path, key = sys.argv[1], sys.argv[2]
    lines = [l for l in open(path) if not l.startswith("==")]
    launches = {}
    order = []
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        if i not in launches:
            launches[i] = {"name": row["Kernel Name"]}
            order.append(i)
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
        launches[i][row["Metric Name"]] = v * scale
    prod = [i for i in order if re.search(r"trace_wave_kernel<0, [01], 0, [034]>|trace_tail_kernel<[01], 0>", launches[i]["name"])]
    # bench.py --steps 1 --warmup 1 renders five un-instrumented frames (warm-up, timed step, two e2e warm-ups, e2e step), each
    # with the same launches (a frame may take several passes): the last fifth of the production launches is one frame
    assert prod and len(prod) % 5 == 0, (len(prod), "expected 5 identical frames")
    frame = prod[-(len(prod) // 5):]
    tot = lambda m: sum(launches[i].get(m, 0.0) for i in frame)  # noqa: E731
    main_launch = max(frame, key=lambda i: launches[i].get("gpu__time_duration.sum", 0.0))
    ent = {
        "dram_bytes_per_launch": int(tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum")),
        "dram_read_bytes": int(tot("dram__bytes_read.sum")), "dram_write_bytes": int(tot("dram__bytes_write.sum")),
        "l1tex_global_load_sector_bytes": int(32 * tot("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")),
        "l2_read_sector_bytes_from_l1": int(32 * tot("lts__t_sectors_srcunit_tex_op_read.sum")),
        "kernels_ms_under_ncu": round(tot("gpu__time_duration.sum"), 4),
        "main_launch": launches[main_launch]["name"], "launches": [launches[i]["name"] for i in frame],
        "source": os.path.relpath(path, os.path.dirname(HERE)),
        "limiter": "instruction issue under divergence (see profiles/r02_*_ncu.txt): DRAM traffic is a few per cent of the HBM peak",
    }
    out = os.path.join(HERE, "r02_dram_traffic.json")
    data = json.load(open(out)) if os.path.exists(out) else {}
    data[key] = ent
    json.dump(data, open(out, "w"), indent=1, sort_keys=True)
    print(key, json.dumps(ent))


if __name__ == "__main__":
    main()
