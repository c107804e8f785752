#!/bin/bash
# One GPU-box pass: parity tests, bench (ours + reference arm), ncu launch list of a short bench run,
# ncu section capture of one launch of each hot kernel.  (tools/eager_gpu_baseline.py gives the eager
# PyTorch-CUDA context number; `ncu --set full` of tools/ncu_kernels.py takes ~3 min per kernel.)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.json | cut -c1-300; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --diffusion-steps 4 --no-cpu-baseline --no-kernel-events > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python tools/summarize_ncu.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -14 gpurun_out/launches_summary.txt | cut -c1-120
timeout 300 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --clock-control none -k regex:"conv1d_tc_kernel|trunk_kernel|gru_cluster_f16" -c 12 -o gpurun_out/prof_sections python tools/ncu_kernels.py > gpurun_out/ncu_sections.log 2>&1; echo "ncu sections rc=$?"
python tools/profile_layers.py > gpurun_out/layers_final.txt 2>&1; head -3 gpurun_out/layers_final.txt
