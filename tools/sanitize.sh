#!/bin/bash
# Runs the small-shape part of the `-m gpu` suite plus the corrupt-payload decoder fuzz under compute-sanitizer.
#   tools/sanitize.sh memcheck | racecheck | synccheck | initcheck        (ONE tool per gpurun call: B200_PROFILING.md)
# Writes gpurun_out/sanitize_<tool>.log and a one-line verdict gpurun_out/sanitize_<tool>.summary; the summaries are committed
# under profiles/.  Large-shape tests (1 MiB+ blocks, whole fixtures) are deselected: the tools slow kernels down 10-100x.
set -u
TOOL=${1:-memcheck}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SMALL='test_gpu_bbwt or test_gpu_mtf or test_gpu_rice or test_gpu_decode or test_gpu_lz77_residual or test_gpu_repair or test_gpu_select or test_gpu_cdc or test_gpu_v2new or test_gpu_fuzz_decoders or test_reference_quirks or test_gpu_pipeline'
SKIP='not 1mib and not 16mib and not 4mib and not large and not fixture and not long_blocks and not big and not one_mib and not 64k'
export KOLM_SANITIZE=1
EXTRA=""
if [ "$TOOL" = "racecheck" ]; then EXTRA="--racecheck-report all"; fi
timeout ${SANITIZE_TIMEOUT:-2400} compute-sanitizer --tool "$TOOL" $EXTRA --log-file "gpurun_out/sanitize_${TOOL}.log" --error-exitcode 99 \
    python -m pytest tests -m gpu -q -x -k "($SMALL) and $SKIP" -p no:cacheprovider > "gpurun_out/sanitize_${TOOL}.pytest.log" 2>&1
RC=$?
ERR=$(grep -c "^========= .*\(Invalid\|Race\|Hazard\|Error\|Uninitialized\|Barrier\)" "gpurun_out/sanitize_${TOOL}.log" 2>/dev/null || true)
TAIL=$(grep "ERROR SUMMARY\|RACECHECK SUMMARY" "gpurun_out/sanitize_${TOOL}.log" | tail -1)
echo "tool=$TOOL exit=$RC reported_lines=$ERR summary='${TAIL}' pytest='$(tail -1 gpurun_out/sanitize_${TOOL}.pytest.log)'" | tee "gpurun_out/sanitize_${TOOL}.summary"
exit 0
