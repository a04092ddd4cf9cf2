# frames in flight: what the L2 flush between steps costs the frame running beside it
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="timeout -k 5 300 python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --warmup 3 --steps 24"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["frame_latency_ms"], d["breakdown"]["trace_ms"], d["breakdown"]["bvh_build_ms"], d["e2e"]["ms_per_step"], d["frame_check"]["status"], d["gpu_launches"])'
run() { echo "$1" >> gpurun_out/r02s_fif.txt; shift; "$@" 2>>gpurun_out/r02s_err.txt | python -c "$J" >> gpurun_out/r02s_fif.txt; }
for f in 1 2 3 4; do
  run "C2 rank0of8 fif=$f noflush" $B --frames-in-flight $f --emulate-rank 0/8 --no-l2-flush
done
run "C2 rank0of8 fif=2 noflush main=6 tail=64" env RTB_WAVE_MAIN_CTAS=6 RTB_WAVE_TAIL_THREADS=64 $B --frames-in-flight 2 --emulate-rank 0/8 --no-l2-flush
run "C2 rank0of8 fif=3 noflush main=6 tail=64" env RTB_WAVE_MAIN_CTAS=6 RTB_WAVE_TAIL_THREADS=64 $B --frames-in-flight 3 --emulate-rank 0/8 --no-l2-flush
run "C2 fif=1 noflush" $B --no-l2-flush
run "C2 fif=2 noflush" $B --frames-in-flight 2 --no-l2-flush
tail -5 gpurun_out/r02s_err.txt
cat gpurun_out/r02s_fif.txt | paste - -
echo done
