# final N=1: default bench line, reference arm, launch lists of the final kernels, smoke
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02v_smoke.txt 2>&1
tail -2 gpurun_out/r02v_smoke.txt
timeout -k 5 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_bench_c2_n1.json 2> gpurun_out/r02v_err.txt
timeout -k 5 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r02_bench_c2_reference_arm_n1.json 2>> gpurun_out/r02v_err.txt
BA="--breakdown none --min-seconds 0 --no-cpu-baseline --no-frame-check"
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches_c2.csv python bench.py --steps 2 --warmup 1 $BA > gpurun_out/r02v_l.log 2>&1
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches_rank0of8.csv python bench.py --steps 2 --warmup 1 --emulate-rank 0/8 $BA > gpurun_out/r02v_l8.log 2>&1
timeout -k 5 300 python bench.py --frames-in-flight 2 --steps 20 --warmup 3 --breakdown none --no-cpu-baseline > gpurun_out/r02_bench_c2_n1_fif2.json 2>> gpurun_out/r02v_err.txt
tail -5 gpurun_out/r02v_err.txt
python - <<'PY'
import json
for f in ("r02_bench_c2_n1", "r02_bench_c2_reference_arm_n1", "r02_bench_c2_n1_fif2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("sustained") or {}).get("ms_per_step"), (d.get("frame_check") or {}).get("status"), d.get("clocks"))
        for k, v in ((d.get("breakdown") or {}).get("configs") or {}).items():
            print("   ", k, v.get("ms_per_step"), v.get("value"), (v.get("frame_check") or {}).get("status"), v.get("error"))
    except Exception as e:
        print(f, "ERR", e)
PY
echo done
