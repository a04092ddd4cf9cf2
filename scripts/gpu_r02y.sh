# radix sort with warp-private ranking + shared-memory staging: tests, build times C2 / C4, launch list of a C4 build
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02y_pytest_gpu.txt
cat gpurun_out/r02y_pytest_gpu.txt
B="timeout -k 5 300 python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --no-frame-check --warmup 2"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["breakdown"]["trace_ms"], d["breakdown"]["bvh_build_ms"])'
for c in C2 C4; do
  echo "$c" >> gpurun_out/r02y_build.txt
  $B --config $c --steps 6 2>>gpurun_out/r02y_err.txt | python -c "$J" >> gpurun_out/r02y_build.txt
done
echo "C2 rank0of8" >> gpurun_out/r02y_build.txt
$B --steps 16 --emulate-rank 0/8 2>>gpurun_out/r02y_err.txt | python -c "$J" >> gpurun_out/r02y_build.txt
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02y_launches_c4.csv python bench.py --config C4 --steps 1 --warmup 0 --breakdown none --min-seconds 0 --no-cpu-baseline --no-frame-check > gpurun_out/r02y_l.log 2>&1
tail -3 gpurun_out/r02y_err.txt
paste - - < gpurun_out/r02y_build.txt
echo done
