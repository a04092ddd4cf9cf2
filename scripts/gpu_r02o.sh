# final-kernel ncu evidence: launch lists (C2, rank 0 of 8), --set full of the main launch (C2, C3, C4), per-config DRAM traffic
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
BA="--breakdown none --min-seconds 0 --no-cpu-baseline --no-frame-check"
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_c2.csv python bench.py --steps 2 --warmup 1 $BA > gpurun_out/r02o_l.log 2>&1
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_rank0of8.csv python bench.py --steps 2 --warmup 1 --emulate-rank 0/8 $BA > gpurun_out/r02o_l8.log 2>&1
K='regex:trace_wave_kernel<.bool.0, .bool.0, .bool.0, .int.[34]>'
for cfg in C2 C3 C4; do
  timeout -k 5 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s 1 -c 1 -f -o gpurun_out/r02f_main_${cfg} python bench.py --config $cfg --steps 1 --warmup 1 $BA > gpurun_out/r02o_f_$cfg.log 2>&1
done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum
for cfg in C1 C2 C3 C4 C5; do
  timeout -k 5 400 ncu --metrics $M --clock-control none -k regex:trace_ --csv --log-file gpurun_out/r02f_traffic_$cfg.csv python bench.py --config $cfg --steps 1 --warmup 1 $BA > gpurun_out/r02o_t_$cfg.log 2>&1
done
for cfg in C3 C5; do
  timeout -k 5 400 ncu --metrics $M --clock-control none -k regex:trace_ --csv --log-file gpurun_out/r02f_traffic_${cfg}x.csv python bench.py --config $cfg --ext --steps 1 --warmup 1 $BA > gpurun_out/r02o_t_${cfg}x.log 2>&1
done
ls -la gpurun_out/r02f_main*.ncu-rep
echo done
