# traversal hierarchy (own tree for the walk + hoisted big leaves): tests + per-config timing
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02n_pytest_gpu.txt
cat gpurun_out/r02n_pytest_gpu.txt
B="timeout -k 5 240 python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --warmup 2"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["breakdown"]["trace_ms"], d["breakdown"]["bvh_build_ms"], d["frame_check"]["status"], d["roofline"]["per_ray"], d["roofline"]["lanes_per_traverse_step"])'
for cfg in C1 C2 C3 C4 C5; do
  echo "$cfg" >> gpurun_out/r02n_ab.txt
  $B --config $cfg --steps 5 2>>gpurun_out/r02n_err.txt | python -c "$J" >> gpurun_out/r02n_ab.txt
done
for cfg in C3 C5; do
  echo "$cfg ext" >> gpurun_out/r02n_ab.txt
  $B --config $cfg --ext --steps 3 2>>gpurun_out/r02n_err.txt | python -c "$J" >> gpurun_out/r02n_ab.txt
done
echo "C2 rank0of8" >> gpurun_out/r02n_ab.txt
$B --steps 10 --emulate-rank 0/8 2>>gpurun_out/r02n_err.txt | python -c "$J" >> gpurun_out/r02n_ab.txt
cat gpurun_out/r02n_ab.txt
tail -c 800 gpurun_out/r02n_err.txt
echo done
