# compute-sanitizer over one small frame that runs every production kernel (3 206 primitives: fused build, hierarchy, sorts, trace, tail)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool" >> gpurun_out/r02_compute_sanitizer.txt
  timeout -k 5 900 compute-sanitizer --tool $tool python scripts/sanitize_case.py 2>&1 | grep -E "sanitize case|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" | head -20 >> gpurun_out/r02_compute_sanitizer.txt
done
cat gpurun_out/r02_compute_sanitizer.txt
echo done
