# two B200s: frames in flight over two library contexts + communicators per rank; the 2-GPU tests
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02u_pytest_multi.txt
cat gpurun_out/r02u_pytest_multi.txt
T="timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["frame_latency_ms"], d["breakdown"]["trace_ms"], d["e2e"]["ms_per_step"], d["frame_check"], d["gpu_launches"])'
for f in 1 2 3; do
  echo "C2 N=2 fif=$f" >> gpurun_out/r02u_fif.txt
  $T --breakdown none --min-seconds 0 --warmup 3 --steps 20 --frames-in-flight $f 2>>gpurun_out/r02u_err.txt | grep '^{' | python -c "$J" >> gpurun_out/r02u_fif.txt
done
tail -8 gpurun_out/r02u_err.txt
cat gpurun_out/r02u_fif.txt
echo done
