# knob re-sweep after hoisting (C2; C3 for the stacking order)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="timeout -k 5 200 python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --no-frame-check --warmup 2 --steps 8"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["breakdown"]["trace_ms"])'
for q in 2 4 6; do for t in 16 20 24; do
  echo "C2 qgate=$q tmin=$t" >> gpurun_out/r02m_sweep.txt
  RTB_WAVE_QGATE=$q RTB_WAVE_TMIN=$t $B 2>/dev/null | python -c "$J" >> gpurun_out/r02m_sweep.txt
done; done
for sp in 0 1; do
  echo "C2 sorted_push=$sp" >> gpurun_out/r02m_sweep.txt
  RTB_WAVE_SORTED_PUSH=$sp $B 2>/dev/null | python -c "$J" >> gpurun_out/r02m_sweep.txt
  echo "C3 sorted_push=$sp" >> gpurun_out/r02m_sweep.txt
  RTB_WAVE_SORTED_PUSH=$sp $B --config C3 --steps 3 2>/dev/null | python -c "$J" >> gpurun_out/r02m_sweep.txt
  echo "C4 sorted_push=$sp" >> gpurun_out/r02m_sweep.txt
  RTB_WAVE_SORTED_PUSH=$sp $B --config C4 --steps 3 2>/dev/null | python -c "$J" >> gpurun_out/r02m_sweep.txt
done
cat gpurun_out/r02m_sweep.txt | paste - -
echo done
