# small scenes over the traversal hierarchy: RTB_WIDE_MIN on C1 (complexScene, ~1.1 k primitives, 800x800, 1 spp)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="timeout -k 5 200 python bench.py --config C1 --breakdown none --min-seconds 0 --no-cpu-baseline --warmup 5 --steps 50"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["breakdown"]["trace_ms"], d["breakdown"]["bvh_build_ms"], d["e2e"]["ms_per_step"], d["frame_check"]["status"], d["gpu_launches"], (d["roofline"].get("per_ray") or {}))'
for w in 8192 64; do
  echo "C1 widemin=$w" >> gpurun_out/r02x_c1.txt
  RTB_WIDE_MIN=$w $B 2>>gpurun_out/r02x_err.txt | python -c "$J" >> gpurun_out/r02x_c1.txt
done
RTB_WIDE_MIN=64 timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02x_pytest_widemin64.txt
cat gpurun_out/r02x_pytest_widemin64.txt
tail -5 gpurun_out/r02x_err.txt
cat gpurun_out/r02x_c1.txt | paste - -
echo done
