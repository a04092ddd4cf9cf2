# N ranks: e2e leg with the scene all-gather on the upload stream
set -x
N=$1
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["frame_check"]["status"], d["frame_check"].get("e2e_frame"))'
for i in 1 2; do
timeout -k 5 400 $T --master-port 2953$i bench.py --gpus $N --steps 20 --warmup 3 --breakdown none --min-seconds 0 2>> gpurun_out/r02ac_err.txt | grep '^{' | python -c "$J" >> gpurun_out/r02ac_n$N.txt
done
tail -3 gpurun_out/r02ac_err.txt
cat gpurun_out/r02ac_n$N.txt
echo done
