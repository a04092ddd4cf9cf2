set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02c_pytest_gpu.txt
cat gpurun_out/r02c_pytest_gpu.txt
B="python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --steps 10 --warmup 2"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["breakdown"]["trace_ms"], d["breakdown"]["bvh_build_ms"], d["frame_check"]["status"], d["e2e"]["ms_per_step"])'
for ov in 1 0; do
  echo "overlap=$ov full" >> gpurun_out/r02c_ab.txt
  RTB_WAVE_TAIL_OVERLAP=$ov $B 2>>gpurun_out/r02c_err.txt | python -c "$J" >> gpurun_out/r02c_ab.txt
  echo "overlap=$ov rank0of8" >> gpurun_out/r02c_ab.txt
  RTB_WAVE_TAIL_OVERLAP=$ov $B --emulate-rank 0/8 2>>gpurun_out/r02c_err.txt | python -c "$J" >> gpurun_out/r02c_ab.txt
done
for cfg in C3 C4 C5; do
  echo "$cfg" >> gpurun_out/r02c_ab.txt
  $B --config $cfg --steps 3 --warmup 1 2>>gpurun_out/r02c_err.txt | python -c "$J" >> gpurun_out/r02c_ab.txt
done
cat gpurun_out/r02c_ab.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "walk_counters or tail_handover_forced" > gpurun_out/r02c_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02c_memcheck.txt
tail -5 gpurun_out/r02c_memcheck.txt
echo done
