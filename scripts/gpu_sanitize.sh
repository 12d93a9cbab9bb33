#!/bin/bash
# compute-sanitizer over the single-GPU kernels (memcheck: out-of-bounds / misaligned accesses incl. the TMA and bulk
# copies; racecheck: shared-memory hazards of the staged sweep and belief-update kernels; synccheck: barrier misuse).
# The reference's counterpart is the IPU floating-point / memory trap configuration (ba/ba.cpp:888-896).
# With 2+ GPUs visible the sharded exchange (one process per GPU, CUDA IPC peers) is run under memcheck as well.
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_run.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_$tool.log | tail -1)"
  grep -E "slam final|relin_mode|fast-math" gpurun_out/r2_sanitizer_$tool.log | head -4
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_run.py fast > gpurun_out/r2_sanitizer_memcheck_fast_math.log 2>&1
echo "== memcheck, fast-math build: rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/r2_sanitizer_memcheck_fast_math.log | tail -1)"; tail -4 gpurun_out/r2_sanitizer_memcheck_fast_math.log | cut -c1-200
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  GBP_P2P_TIMEOUT_S=600 timeout 1500 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 \
    python -m pytest tests/test_multigpu.py -q -x -m gpu -k "block_calls and 2" > gpurun_out/r2_sanitizer_memcheck_2gpu.log 2>&1
  echo "== memcheck, 2 ranks: rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/r2_sanitizer_memcheck_2gpu.log | sort | uniq -c | tail -3)"
  tail -3 gpurun_out/r2_sanitizer_memcheck_2gpu.log
fi
