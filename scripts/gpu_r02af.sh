# refresh of the one-GPU default line with the final tree
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_bench_c2_n1.json 2> gpurun_out/r02af_err.txt
tail -3 gpurun_out/r02af_err.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_c2_n1.json").read().strip().splitlines()[-1])
print(d.get("value"), d.get("ms_per_step"), d["e2e"]["value"], d["e2e"]["ms_per_step"], d["sustained"]["ms_per_step"], d["frame_check"]["status"], d["clocks"], d["breakdown"]["bvh_build_ms"])
for k, v in d["breakdown"]["configs"].items():
    print("   ", k, v.get("ms_per_step"), v.get("value"), v.get("mrays_traversed_per_s"), (v.get("frame_check") or {}).get("status"), v.get("error"))
PY
echo done
