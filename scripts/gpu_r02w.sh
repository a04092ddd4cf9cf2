# final N = $1: default bench line under torchrun, reference arm, (N = 8: frames-in-flight variant, C++ host)
set -x
N=$1
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -k 5 900 $T --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/r02w_err_n$N.txt | grep '^{' > gpurun_out/r02_bench_c2_n$N.json
timeout -k 5 400 $T --master-port 29522 bench.py --gpus $N --steps 20 --warmup 3 --frames-in-flight 2 --breakdown none 2>> gpurun_out/r02w_err_n$N.txt | grep '^{' > gpurun_out/r02_bench_c2_n${N}_fif2.json
if [ "$N" = "2" ]; then
  timeout -k 5 600 $T --master-port 29523 bench.py --impl reference --gpus $N --steps 5 --warmup 1 2>> gpurun_out/r02w_err_n$N.txt | grep '^{' > gpurun_out/r02_bench_c2_reference_arm_n$N.json
fi
tail -5 gpurun_out/r02w_err_n$N.txt
python - <<PY
import json
for f in ("r02_bench_c2_n$N", "r02_bench_c2_n${N}_fif2", "r02_bench_c2_reference_arm_n$N"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("frame_latency_ms"), (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("ms_per_step"), (d.get("sustained") or {}).get("ms_per_step"), (d.get("frame_check") or {}).get("status"), (d.get("cpu_baseline") or {}).get("cores"))
        for k, v in ((d.get("breakdown") or {}).get("configs") or {}).items():
            print("   ", k, v.get("ms_per_step"), v.get("value"), (v.get("frame_check") or {}).get("status"), v.get("error"))
    except Exception as e:
        print(f, "ERR", e)
PY
echo done
