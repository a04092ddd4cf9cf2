# tail kernel with dealt-out leaf candidates; frames in flight with a reserved slot (RTB_WAVE_MAIN_CTAS=6, tail CTAs of 64)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02r_pytest_gpu.txt
cat gpurun_out/r02r_pytest_gpu.txt
B="timeout -k 5 300 python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --warmup 3 --steps 24"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["frame_latency_ms"], d["breakdown"]["trace_ms"], d["breakdown"]["bvh_build_ms"], d["e2e"]["ms_per_step"], d["frame_check"]["status"], d["gpu_launches"])'
run() { echo "$1" >> gpurun_out/r02r_fif.txt; shift; "$@" 2>>gpurun_out/r02r_err.txt | python -c "$J" >> gpurun_out/r02r_fif.txt; }
run "C2 fif=1" $B
run "C2 rank0of8 fif=1" $B --emulate-rank 0/8
for f in 2 3; do for m in 6 5; do for t in 64 128; do
  run "C2 rank0of8 fif=$f main=$m tail=$t" env RTB_WAVE_MAIN_CTAS=$m RTB_WAVE_TAIL_THREADS=$t $B --frames-in-flight $f --emulate-rank 0/8
done; done; done
run "C2 fif=1 main=6" env RTB_WAVE_MAIN_CTAS=6 $B
run "C2 fif=2 main=6 tail=64" env RTB_WAVE_MAIN_CTAS=6 RTB_WAVE_TAIL_THREADS=64 $B --frames-in-flight 2
run "C2 fif=3 main=6 tail=64" env RTB_WAVE_MAIN_CTAS=6 RTB_WAVE_TAIL_THREADS=64 $B --frames-in-flight 3
run "C2 rank0of8 fif=1 main=6" env RTB_WAVE_MAIN_CTAS=6 $B --emulate-rank 0/8
tail -5 gpurun_out/r02r_err.txt
cat gpurun_out/r02r_fif.txt | paste - -
echo done
