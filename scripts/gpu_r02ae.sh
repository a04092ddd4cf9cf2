# single-block sort for small arrays: GPU tier, C1 timing, sanitizer on a scene that takes it
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_final.txt
cat gpurun_out/r02_pytest_gpu_final.txt
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool (1 200 triangles: single-block sort)" >> gpurun_out/r02_compute_sanitizer_small.txt
  timeout -k 5 600 compute-sanitizer --tool $tool python scripts/sanitize_case.py 1200 2>&1 | grep -E "sanitize case|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" | head -20 >> gpurun_out/r02_compute_sanitizer_small.txt
done
cat gpurun_out/r02_compute_sanitizer_small.txt
B="timeout -k 5 200 python bench.py --config C1 --breakdown none --min-seconds 0 --no-cpu-baseline --warmup 5 --steps 50"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["breakdown"]["trace_ms"], d["breakdown"]["bvh_build_ms"], d["e2e"]["ms_per_step"], d["frame_check"]["status"], d["gpu_launches"])'
$B 2>>gpurun_out/r02ae_err.txt | python -c "$J" > gpurun_out/r02ae_c1.txt
$B 2>>gpurun_out/r02ae_err.txt | python -c "$J" >> gpurun_out/r02ae_c1.txt
cat gpurun_out/r02ae_c1.txt
echo done
