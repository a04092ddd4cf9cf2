# ncu --set full of the main trace launch: default build and the shared-memory-top A/B build, C2 and C4; launch list of the per-rank workload.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
BA="--breakdown none --min-seconds 0 --no-cpu-baseline --no-frame-check"
K='regex:trace_wave_kernel<.bool.0, .bool.0, .bool.0, .int.3>'
for cfg in C2 C4; do
  timeout -k 5 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s 1 -c 1 -f -o gpurun_out/r02_main_${cfg} python bench.py --config $cfg --steps 1 --warmup 1 $BA > gpurun_out/r02h_f_$cfg.log 2>&1
done
cp raytracergpu_mastersproject_b200/librtb200.so /tmp/librtb200_default.so
cp raytracergpu_mastersproject_b200/librtb200_smemtop.so raytracergpu_mastersproject_b200/librtb200.so
for cfg in C2 C4; do
  timeout -k 5 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s 1 -c 1 -f -o gpurun_out/r02_main_${cfg}_smemtop python bench.py --config $cfg --steps 1 --warmup 1 $BA > gpurun_out/r02h_fs_$cfg.log 2>&1
done
cp /tmp/librtb200_default.so raytracergpu_mastersproject_b200/librtb200.so
# C3 (sphere field): the farthest-first-stacking variant
timeout -k 5 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:trace_wave_kernel<.bool.0, .bool.0, .bool.0, .int.4>' -s 1 -c 1 -f -o gpurun_out/r02_main_C3 python bench.py --config C3 --steps 1 --warmup 1 $BA > gpurun_out/r02h_f_C3.log 2>&1
ls -la gpurun_out/*.ncu-rep
grep -h "No kernels" gpurun_out/r02h_f*.log
echo done
