# S-phase gating sweep + hand-over thresholds on the final kernels; tests incl. the new fuzz
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02p_pytest_gpu.txt
cat gpurun_out/r02p_pytest_gpu.txt
B="timeout -k 5 200 python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --no-frame-check --warmup 2 --steps 8"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["breakdown"]["trace_ms"])'
for s in 1 6 10 14 18; do
  echo "C2 smin=$s" >> gpurun_out/r02p_sweep.txt
  RTB_WAVE_SMIN=$s $B 2>/dev/null | python -c "$J" >> gpurun_out/r02p_sweep.txt
done
for s in 1 10 14; do
  echo "C3 smin=$s" >> gpurun_out/r02p_sweep.txt
  RTB_WAVE_SMIN=$s $B --config C3 --steps 3 2>/dev/null | python -c "$J" >> gpurun_out/r02p_sweep.txt
  echo "C4 smin=$s" >> gpurun_out/r02p_sweep.txt
  RTB_WAVE_SMIN=$s $B --config C4 --steps 3 2>/dev/null | python -c "$J" >> gpurun_out/r02p_sweep.txt
  echo "C5 smin=$s" >> gpurun_out/r02p_sweep.txt
  RTB_WAVE_SMIN=$s $B --config C5 --steps 3 2>/dev/null | python -c "$J" >> gpurun_out/r02p_sweep.txt
  echo "C2 rank0of8 smin=$s" >> gpurun_out/r02p_sweep.txt
  RTB_WAVE_SMIN=$s $B --emulate-rank 0/8 2>/dev/null | python -c "$J" >> gpurun_out/r02p_sweep.txt
done
for t in 12 20 32 48; do for c in 8 16; do
  echo "C2 rank0of8 turns=$t coop=$c" >> gpurun_out/r02p_sweep.txt
  RTB_WAVE_COOP_TURNS=$t RTB_WAVE_COOP=$c $B --emulate-rank 0/8 2>/dev/null | python -c "$J" >> gpurun_out/r02p_sweep.txt
done; done
for t in 12 20 48; do
  echo "C2 turns=$t" >> gpurun_out/r02p_sweep.txt
  RTB_WAVE_COOP_TURNS=$t $B 2>/dev/null | python -c "$J" >> gpurun_out/r02p_sweep.txt
done
cat gpurun_out/r02p_sweep.txt | paste - -
echo done
