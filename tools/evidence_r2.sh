# Round-2 evidence run (GPU box): ncu launch list with DRAM bytes over a reduced-step bench, ncu sections
# of the hot kernels (exported to CSV on the box: the .ncu-rep is too large to travel), bench lines.
set -x
OU_PIPELINE=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --diffusion-steps 4 --no-gpu-baseline --no-other-configs --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --clock-control none -o /tmp/r2_prof_kernels python tools/ncu_kernels.py > gpurun_out/r2_ncu_kernels.log 2>&1
ncu -i /tmp/r2_prof_kernels.ncu-rep --page raw --csv > gpurun_out/r2_kernels_ncu_raw.csv 2>/dev/null
timeout 600 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2l_bench_reference.json 2> gpurun_out/r2l_bench_reference.err
cut -c1-300 gpurun_out/r2l_bench.json
du -sh gpurun_out; ls -la gpurun_out | tail -8
