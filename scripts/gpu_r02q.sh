# build overlap + frames in flight (one B200): C2 whole frame and the bands of rank 0 of 8
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02q_pytest_gpu.txt
cat gpurun_out/r02q_pytest_gpu.txt
B="timeout -k 5 300 python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --warmup 3 --steps 24"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["frame_latency_ms"], d["breakdown"]["trace_ms"], d["breakdown"]["bvh_build_ms"], d["e2e"]["ms_per_step"], d["frame_check"]["status"], d["gpu_launches"])'
for f in 1 2 3; do
  echo "C2 fif=$f" >> gpurun_out/r02q_fif.txt
  $B --frames-in-flight $f 2>gpurun_out/r02q_err_$f.txt | python -c "$J" >> gpurun_out/r02q_fif.txt
  echo "C2 rank0of8 fif=$f" >> gpurun_out/r02q_fif.txt
  $B --frames-in-flight $f --emulate-rank 0/8 2>>gpurun_out/r02q_err_$f.txt | python -c "$J" >> gpurun_out/r02q_fif.txt
done
tail -5 gpurun_out/r02q_err_2.txt
cat gpurun_out/r02q_fif.txt | paste - -
echo done
