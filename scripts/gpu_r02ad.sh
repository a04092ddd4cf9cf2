# last check of the round: the whole GPU tier (incl. bench.py's JSON line) and smoke() with the final tree
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_final.txt
cat gpurun_out/r02_pytest_gpu_final.txt
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo done
