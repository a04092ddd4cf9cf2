# ncu evidence for round 2: launch list, --set full of the main trace launch (default and the shared-memory-top A/B build, C2 and C4),
# per-config DRAM traffic.  Every command under `timeout`.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
BA="--breakdown none --min-seconds 0 --no-cpu-baseline --no-frame-check"
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 2 --warmup 1 $BA > gpurun_out/r02e_l.log 2>&1
K='regex:trace_wave_kernel<0, 0, 0, 3>'
for cfg in C2 C4; do
  timeout -k 5 600 ncu --set full --clock-control none --import-source on -k "$K" -s 1 -c 1 -f -o gpurun_out/r02_main_${cfg} python bench.py --config $cfg --steps 1 --warmup 1 $BA > gpurun_out/r02e_f_$cfg.log 2>&1
done
# A/B: top 128 records staged in shared memory
cp raytracergpu_mastersproject_b200/librtb200.so /tmp/librtb200_default.so
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["breakdown"]["trace_ms"], d["frame_check"]["status"])'
for cfg in C2 C4; do
  echo "default $cfg" >> gpurun_out/r02_smemtop_ab.txt
  timeout -k 5 200 python bench.py --config $cfg --steps 10 --warmup 2 --breakdown none --min-seconds 0 --no-cpu-baseline 2>/dev/null | python -c "$J" >> gpurun_out/r02_smemtop_ab.txt
done
cp raytracergpu_mastersproject_b200/librtb200_smemtop.so raytracergpu_mastersproject_b200/librtb200.so
for cfg in C2 C4; do
  echo "smemtop128 $cfg" >> gpurun_out/r02_smemtop_ab.txt
  timeout -k 5 200 python bench.py --config $cfg --steps 10 --warmup 2 --breakdown none --min-seconds 0 --no-cpu-baseline 2>/dev/null | python -c "$J" >> gpurun_out/r02_smemtop_ab.txt
  timeout -k 5 600 ncu --set full --clock-control none --import-source on -k "$K" -s 1 -c 1 -f -o gpurun_out/r02_main_${cfg}_smemtop python bench.py --config $cfg --steps 1 --warmup 1 $BA > gpurun_out/r02e_fs_$cfg.log 2>&1
done
cp /tmp/librtb200_default.so raytracergpu_mastersproject_b200/librtb200.so
cat gpurun_out/r02_smemtop_ab.txt
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum
for cfg in C1 C2 C3 C4 C5; do
  timeout -k 5 400 ncu --metrics $M --clock-control none -k regex:trace_ --csv --log-file gpurun_out/r02_traffic_$cfg.csv python bench.py --config $cfg --steps 1 --warmup 1 $BA > gpurun_out/r02e_t_$cfg.log 2>&1
done
for cfg in C3 C5; do
  timeout -k 5 400 ncu --metrics $M --clock-control none -k regex:trace_ --csv --log-file gpurun_out/r02_traffic_${cfg}x.csv python bench.py --config $cfg --ext --steps 1 --warmup 1 $BA > gpurun_out/r02e_t_${cfg}x.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
echo done
# experiment: what the concurrent tail launch waits for (polling bound 20 ms / 2 ms / 0.2 ms)
B="timeout -k 5 120 python bench.py --breakdown none --min-seconds 0 --no-cpu-baseline --no-frame-check --steps 10 --warmup 2 --emulate-rank 0/8"
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["breakdown"]["trace_ms"])'
for us in 20000 2000 200; do
  echo "overlap=1 spin_us=$us rank0of8" >> gpurun_out/r02_tail_spin.txt
  RTB_WAVE_TAIL_OVERLAP=1 RTB_WAVE_TAIL_SPIN_US=$us $B 2>/dev/null | python -c "$J" >> gpurun_out/r02_tail_spin.txt
done
cat gpurun_out/r02_tail_spin.txt
echo done2
