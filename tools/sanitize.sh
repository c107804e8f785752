#!/usr/bin/env bash
# compute-sanitizer over the warp-specialised kernels (GPU box only):
#     tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck ...]     (default: memcheck racecheck synccheck)
# Runs a small subset of tests/test_gpu_kernels.py (one case per kernel family / code path) under each
# tool and writes gpurun_out/sanitize_<tool>.log plus a one-line-per-tool summary
# gpurun_out/sanitize_summary.txt (copied to profiles/ by hand).  The subset is chosen with -k so that a
# run stays within a few minutes: the sanitizer slows the tcgen05 / TMA kernels down by 10-100x.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${*:-memcheck racecheck synccheck}
CS=${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
# one conv per role mix (resident / streamed weights, input PReLU transform, FiLM, strided, up), both trunk
# widths, the GRU cluster kernel, signal kernels
SUBSET='(test_conv1d_vs_emulator and (t300 or t1001 or t403 or t500 or t601 or t777 or t1000 or t3200)) or (test_conv_trunk_vs_emulator and (t1000 or t123 or t251 or t868 or t757)) or (test_conv_trunk_with_up_tail_vs_emulator and (t1000 or t123 or t366)) or (test_conv_trunk_with_output_tail_vs_emulator and (t1000 or t251 or t750 or t501)) or (test_gru_vs_explicit and not 801) or test_input_and_output_kernels or test_conv_resident_weights or test_alias_free_snake or test_mel_vs_oracle or test_pad_normalize'
: > gpurun_out/sanitize_summary.txt
for tool in $TOOLS; do
  log=gpurun_out/sanitize_${tool}.log
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check no"
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  OU_GPU_TEST_TIMEOUT=1200 timeout ${SANITIZE_TIMEOUT:-1500} "$CS" --tool "$tool" $extra --target-processes all --print-limit 2000 \
      --error-exitcode 86 \
      python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "$SUBSET" \
      > "$log" 2>&1
  rc=$?
  errs=$(grep -c "^========= .*\(Error\|error\|hazard\|Hazard\)" "$log" || true)
  summ=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
  tests=$(grep -E "passed|failed" "$log" | tail -1)
  echo "$tool: rc=$rc  reports=$errs  [$summ]  pytest: $tests" | tee -a gpurun_out/sanitize_summary.txt
  # where the reports point (kernel + source line), most frequent first
  grep -E "^=========     (at |Write Thread|Read Thread)" "$log" | sed -E 's/\+0x[0-9a-f]+//; s/[Tt]hread \([0-9,]+\)/thread/; s/\(ou::[A-Za-z]+\)//' \
      | sort | uniq -c | sort -rn | head -12 | sed 's/^/    /' | tee -a gpurun_out/sanitize_summary.txt
done
