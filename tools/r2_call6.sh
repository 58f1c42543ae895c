# Round 2, GPU call 6: backward precision diagnostic (gradient hi-only; weights hi+lo vs hi-only), CLIP backward on the two-term split; tests; A/B bench
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
(timeout 400 python tests/diag/diag_grad_planes.py 2>&1) > gpurun_out/c6_diag.log; leg "diag rc=$?"; grep -v Warning gpurun_out/c6_diag.log
PT="timeout 900 python -m pytest -m gpu -q --no-header -p no:cacheprovider"
($PT tests --durations=5 2>&1) > gpurun_out/c6_tests.log; leg "whole suite: $(tail -n 1 gpurun_out/c6_tests.log)"
for v in x2 x1 x2 x1; do
  (STYLEMC_BWD_PREC=$v timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c6_bench_$v.json 2> gpurun_out/c6_bench_$v.err); leg "bench bwd_prec=$v: $(cut -c1-170 gpurun_out/c6_bench_$v.json)"
done
