#!/bin/bash
# compute-sanitizer over the GPU tests that exercise every kernel family (race detection / sanitizers).
#   [FILES="tests/a.py tests/b.py"] [TAG=x] tools/sanitize.sh [memcheck|racecheck] [pytest -k expression]      logs -> gpurun_out/sanitize_<tool>[_TAG].log
# Run on a B200 box:  gpurun --timeout 1500 -- 'bash tools/sanitize.sh memcheck; bash tools/sanitize.sh racecheck'
set -u
cd "$(dirname "$0")/.."
TOOL=${1:-memcheck}
SEL=${2:-"block_small_forced or block_vs_sequential_and_oracle_dc or front_op_monte_carlo or front_blocked_path or front_chunks_and_windows or growth_flag or fallback_plan or monte_carlo_transient_trapezoidal or test_op"}
FILES=${FILES:-"tests/test_gpu_chain_block.py tests/test_gpu_front.py tests/test_gpu_accuracy.py tests/test_gpu_parity.py"}
mkdir -p gpurun_out
LOG=gpurun_out/sanitize_${TOOL}${TAG:+_$TAG}.log
# small problem sizes only: the sanitizer slows kernels down 10-100x
compute-sanitizer --tool "$TOOL" --error-exitcode 66 --launch-timeout 600 \
    python -m pytest $FILES -m gpu -q -x -k "$SEL" \
    > "$LOG" 2>&1
RC=$?
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" "$LOG" | tail -5
echo "sanitize $TOOL rc=$RC (66 = the sanitizer reported errors)"
exit $RC
