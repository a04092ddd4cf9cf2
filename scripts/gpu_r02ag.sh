# two B200s: the whole GPU tier on the final tree
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_multi_2gpus.txt
cat gpurun_out/r02_pytest_gpu_multi_2gpus.txt
echo done
