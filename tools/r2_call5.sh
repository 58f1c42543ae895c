# Round 2, GPU call 5: two-term split (hi-only gradient planes) in the backward pass: diagnostic, tests, A/B bench, launch list
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
(timeout 300 python tests/diag/diag_grad_planes.py 2>&1) > gpurun_out/c5_diag.log; leg "diag rc=$?"; cat gpurun_out/c5_diag.log | grep -v Warning
PT="timeout 900 python -m pytest -m gpu -q --no-header -p no:cacheprovider"
($PT tests --durations=5 2>&1) > gpurun_out/c5_tests.log; leg "whole suite: $(tail -n 1 gpurun_out/c5_tests.log)"
for v in 1 0 1 0; do
  (STYLEMC_GRAD_LO=$v timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c5_bench_lo$v.json 2> gpurun_out/c5_bench_lo$v.err); leg "bench grad_lo=$v: $(cut -c1-170 gpurun_out/c5_bench_lo$v.json)"
done
CMD="python bench.py --steps 1 --warmup 1 --batch 64 --micro-batch 64 --no-cpu-baseline --profile-step"
(timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1); leg "launch list: $(grep -c hconv_kernel gpurun_out/launches.csv) hconv rows"
