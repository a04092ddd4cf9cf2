# drain hand-over threshold (RTB_WAVE_COOP) below 8 with the new hierarchy
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="timeout -k 5 300 python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --no-frame-check --warmup 3 --steps 16"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["breakdown"]["trace_ms"], d["roofline"]["walk_counters"]["parked"], d["roofline"]["walk_counters"]["tailRays"], d["roofline"]["walk_counters"]["tailTurns"])'
run() { echo "$1" >> gpurun_out/r02t_coop.txt; shift; "$@" 2>>gpurun_out/r02t_err.txt | python -c "$J" >> gpurun_out/r02t_coop.txt; }
for c in 0 1 2 4 6 8; do
  run "C2 rank0of8 coop=$c" env RTB_WAVE_COOP=$c $B --emulate-rank 0/8
done
for c in 0 2 4; do
  run "C2 coop=$c" env RTB_WAVE_COOP=$c $B
done
for t in 64 128; do
  run "C2 rank0of8 coop=2 turns=$t" env RTB_WAVE_COOP=2 RTB_WAVE_COOP_TURNS=$t $B --emulate-rank 0/8
done
tail -5 gpurun_out/r02t_err.txt
cat gpurun_out/r02t_coop.txt | paste - -
echo done
