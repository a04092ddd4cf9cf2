# A/B: software prefetch of pushed records / queued candidates (L1 and L2 variants) against the default build
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=raytracergpu_mastersproject_b200
cp $P/librtb200.so /tmp/default.so
B="timeout -k 5 240 python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --warmup 2"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["breakdown"]["trace_ms"], d["frame_check"]["status"])'
for v in default pf1 pf2; do
  if [ $v = default ]; then cp /tmp/default.so $P/librtb200.so; else cp $P/librtb200_$v.so $P/librtb200.so; fi
  for cfg in C2 C3 C4 C5; do
    echo "$v $cfg" >> gpurun_out/r02_prefetch_ab.txt
    $B --config $cfg --steps 5 2>>gpurun_out/r02j_err.txt | python -c "$J" >> gpurun_out/r02_prefetch_ab.txt
  done
done
cp /tmp/default.so $P/librtb200.so
cat gpurun_out/r02_prefetch_ab.txt | paste - -
echo done
