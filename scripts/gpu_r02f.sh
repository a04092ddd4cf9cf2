# 2 GPUs: the library's collectives (pytest) + bench.py under torchrun.  Every command under `timeout`.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout -k 5 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02f_pytest_multi.txt
cat gpurun_out/r02f_pytest_multi.txt
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02f_bench_c2_n2.json 2> gpurun_out/r02f_err.txt
tail -c 1500 gpurun_out/r02f_err.txt
head -c 600 gpurun_out/r02f_bench_c2_n2.json
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02f_ref_n2.json 2>> gpurun_out/r02f_err.txt
cat gpurun_out/r02f_ref_n2.json | head -c 400
echo done
