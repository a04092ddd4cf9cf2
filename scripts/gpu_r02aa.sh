# ncu --set full of the heaviest S1 (build) kernels at C4 (10 M primitives): climbs, leaf pass, sort scatter / hist, 4-ary pack
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
BA="--config C4 --steps 1 --warmup 0 --breakdown none --min-seconds 0 --no-cpu-baseline --no-frame-check"
for k in refit_kernel tt_refit_kernel sort_scatter_kernel sort_hist_kernel hlbvh_kernel tt_pack_kernel; do
  timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02g_build_$k python bench.py $BA > gpurun_out/r02aa_$k.log 2>&1
done
ls -la gpurun_out/r02g_build_*.ncu-rep
echo done
