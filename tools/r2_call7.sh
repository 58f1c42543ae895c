# Round 2, GPU call 7: NADA losses on the GPU, whole suite, bench; A/B of the vertical-first unprocess
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
PT="timeout 900 python -m pytest -m gpu -q --no-header -p no:cacheprovider"
($PT tests/test_clip_gpu.py -k nada -s 2>&1) > gpurun_out/c7_nada.log; leg "nada: $(tail -n 1 gpurun_out/c7_nada.log)"; grep -E "^nada|Error|assert" gpurun_out/c7_nada.log | head
($PT tests --durations=5 2>&1) > gpurun_out/c7_tests.log; leg "whole suite: $(tail -n 1 gpurun_out/c7_tests.log)"
for v in 0 1 0 1; do
  (STYLEMC_RESAMPLE_VFIRST=$v timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c7_bench_vf$v.json 2> gpurun_out/c7_bench_vf$v.err); leg "bench vfirst=$v: $(cut -c1-170 gpurun_out/c7_bench_vf$v.json)"
done
