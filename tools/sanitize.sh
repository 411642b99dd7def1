# compute-sanitizer passes over the small-shape GPU tests (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards
# between the phases of a kernel, including the named-barrier hand-over of the warp-specialised live kernel).  Logs -> gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x -m gpu tests/test_gpu_parity.py -k "live_rings_rotate_and_canonicalise or fused_short_call_frontend or features_bit_exact_random_and_edge or model_golden_config0 or staged_remote or reset_by_id" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest -q -x -m gpu tests/test_gpu_parity.py -k "live_rings_rotate_and_canonicalise or fused_short_call_frontend" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitize_racecheck.log
tail -5 gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_racecheck.log
