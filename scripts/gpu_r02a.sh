set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_pytest_gpu.txt
cat gpurun_out/r02a_pytest_gpu.txt | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench_c2_n1.json 2> gpurun_out/r02a_bench_err.txt
tail -c 1500 gpurun_out/r02a_bench_err.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_ref.json 2>> gpurun_out/r02a_bench_err.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02a_launches_rank0of8.csv python bench.py --steps 2 --warmup 1 --emulate-rank 0/8 --breakdown none --min-seconds 0 --no-cpu-baseline > gpurun_out/r02a_ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02a_launches_c2.csv python bench.py --steps 2 --warmup 1 --breakdown none --min-seconds 0 --no-cpu-baseline > gpurun_out/r02a_ncu_bench2.log 2>&1
echo done
