# two B200s: the whole GPU tier (incl. tests/test_gpu_multi.py) with the final library
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02z_pytest_gpu_2gpus.txt
cat gpurun_out/r02z_pytest_gpu_2gpus.txt
cd raytracergpu_mastersproject_b200/host && timeout -k 5 300 ./rtb200_main --gpus 2 --scene meshRoom:110:9 --width 320 --height 180 --spp 4 --out /tmp/o.ppm 2>&1 | tail -3
echo done
