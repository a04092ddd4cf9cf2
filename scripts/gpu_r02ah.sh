# alpha array split from the colour slots: GPU tier + timing of the configs with many samples
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_final.txt
cat gpurun_out/r02_pytest_gpu_final.txt
B="timeout -k 5 300 python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --warmup 2"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["breakdown"]["trace_ms"], d["frame_check"]["status"], d["roofline"]["frac"])'
for c in C2 C5 C4; do
  echo "$c" >> gpurun_out/r02ah.txt
  $B --config $c --steps 6 2>>gpurun_out/r02ah_err.txt | python -c "$J" >> gpurun_out/r02ah.txt
done
tail -3 gpurun_out/r02ah_err.txt
paste - - < gpurun_out/r02ah.txt
echo done
