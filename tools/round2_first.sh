# First GPU call of round 2 (run under gpurun, one GPU): what round 1 built after its GPU budget was spent and could only check on the CPU
# (row-block attention for ViT-B/16, clip_type='double'), then compute-sanitizer over those kernels.
# Everything lands in gpurun_out/; each leg has its own timeout.
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
(timeout 300 python __graft_entry__.py smoke 2>&1) > gpurun_out/r2_smoke.log; leg "smoke: $(tail -n 1 gpurun_out/r2_smoke.log)"
(timeout 600 python -m pytest tests/test_zz_clip_b16_gpu.py -m gpu -q -s --no-header -p no:cacheprovider 2>&1) > gpurun_out/r2_b16.log; leg "ViT-B/16 tests: $(tail -n 1 gpurun_out/r2_b16.log)"
(timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_zz_clip_b16_gpu.py -m gpu -q --no-header -p no:cacheprovider \
   -k "attention" 2>&1) > gpurun_out/r2_memcheck.log; leg "memcheck: $(grep -c 'ERROR SUMMARY: 0 errors' gpurun_out/r2_memcheck.log) clean summaries, $(tail -n 1 gpurun_out/r2_memcheck.log)"
(timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_zz_clip_b16_gpu.py -m gpu -q --no-header -p no:cacheprovider \
   -k "attention_197" 2>&1) > gpurun_out/r2_racecheck.log; leg "racecheck: $(tail -n 1 gpurun_out/r2_racecheck.log)"
(timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1) > gpurun_out/r2_tests.log; leg "whole suite: $(tail -n 1 gpurun_out/r2_tests.log)"
(timeout 400 python bench.py --clip-type double --no-cpu-baseline > gpurun_out/r2_bench_double.json 2> gpurun_out/r2_bench_double.err); leg "bench clip_type=double: $(cut -c1-200 gpurun_out/r2_bench_double.json)"
