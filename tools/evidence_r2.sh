# Round-2 evidence run (GPU box): parity tests, sanitizer, ncu launch list with DRAM bytes over a reduced-step
# bench, ncu sections of the hot kernels (exported to CSV on the box: the .ncu-rep is too large to travel),
# bench lines of both arms, per-layer table.  Outputs: gpurun_out/r2e_*; copied / condensed into profiles/ by
# tools/ncu_traffic.py, tools/summarize_ncu.py, tools/ncu_share.py, tools/ncu_sections_csv.py.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; tail -2 gpurun_out/r2e_pytest.log
SANITIZE_TIMEOUT=600 bash tools/sanitize.sh > /dev/null 2>&1; cat gpurun_out/sanitize_summary.txt
OU_PIPELINE=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 1 --warmup 1 --diffusion-steps 4 --no-gpu-baseline --no-other-configs --no-cpu-baseline > gpurun_out/r2e_ncu_bench.log 2>&1
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv1d_tc_kernel|trunk_kernel|gru_cluster_f16" -o /tmp/r2e_prof_kernels python tools/ncu_kernels.py > gpurun_out/r2e_ncu_kernels.log 2>&1
ncu -i /tmp/r2e_prof_kernels.ncu-rep --page raw --csv > gpurun_out/r2e_kernels_ncu_raw.csv 2>/dev/null
timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2e_bench_reference.json 2> gpurun_out/r2e_bench_reference.err
python tools/profile_layers.py > gpurun_out/r2e_layers.txt 2>&1
python tools/profile_layers.py --config universepp_24k --batch 4 --seconds 10 > gpurun_out/r2e_layers_cfg4.txt 2>&1
cut -c1-300 gpurun_out/r2e_bench.json
tail -5 gpurun_out/r2e_ncu_kernels.log
du -sh gpurun_out; ls -la gpurun_out | tail -8
