# potential study: C3 with the sphere hit-point slack scaled by 0.01 (results not trusted), against the default build
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=raytracergpu_mastersproject_b200
cp $P/librtb200.so /tmp/default.so
B="timeout -k 5 240 python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --warmup 1 --steps 3 --config C3"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["breakdown"]["trace_ms"], d["frame_check"]["status"], d["roofline"]["per_ray"])'
echo "default C3" >> gpurun_out/r02k.txt
$B 2>>gpurun_out/r02k_err.txt | python -c "$J" >> gpurun_out/r02k.txt
cp $P/librtb200_eta.so $P/librtb200.so
echo "eta x0.01 C3" >> gpurun_out/r02k.txt
$B 2>>gpurun_out/r02k_err.txt | python -c "$J" >> gpurun_out/r02k.txt
cp /tmp/default.so $P/librtb200.so
cat gpurun_out/r02k.txt
echo done
