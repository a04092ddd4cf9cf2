# 8 GPUs: bench.py under torchrun (C2 tiles; breakdown: C4 tiles, C5 sample ranges) + the C++ host on 8 GPUs.  Every command under `timeout`.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout -k 5 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02g_bench_c2_n8.json 2> gpurun_out/r02g_err.txt
tail -c 1200 gpurun_out/r02g_err.txt
head -c 500 gpurun_out/r02g_bench_c2_n8.json
mkdir -p /tmp/h8 && cd /tmp/h8 && timeout -k 5 120 $GRAFT_REPO_ROOT/raytracergpu_mastersproject_b200/host/rtb200_main --gpus 8 meshRoom:110:9 1920 1080 > $GRAFT_REPO_ROOT/gpurun_out/r02g_host8.txt 2>&1; echo "host rc=$?" >> $GRAFT_REPO_ROOT/gpurun_out/r02g_host8.txt
tail -4 $GRAFT_REPO_ROOT/gpurun_out/r02g_host8.txt
echo done
