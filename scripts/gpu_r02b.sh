set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02b_pytest_gpu.txt
cat gpurun_out/r02b_pytest_gpu.txt
B="python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --steps 10 --warmup 2 --no-frame-check"
for t in 12 16 24 32; do for c in 8 16 24; do
  echo "turns=$t coop=$c rank0of8" >> gpurun_out/r02b_sweep.txt
  RTB_WAVE_COOP_TURNS=$t RTB_WAVE_COOP=$c $B --emulate-rank 0/8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown']['trace_ms'], d['frame_check']['status'])" >> gpurun_out/r02b_sweep.txt
done; done
for t in 16 24 32 48; do for c in 8 16; do
  echo "turns=$t coop=$c full" >> gpurun_out/r02b_sweep.txt
  RTB_WAVE_COOP_TURNS=$t RTB_WAVE_COOP=$c $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['breakdown']['trace_ms'], d['frame_check']['status'])" >> gpurun_out/r02b_sweep.txt
done; done
cat gpurun_out/r02b_sweep.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02b_launches_rank0of8.csv python bench.py --steps 2 --warmup 1 --emulate-rank 0/8 --breakdown none --min-seconds 0 --no-cpu-baseline > gpurun_out/r02b_ncu_bench.log 2>&1
echo done
