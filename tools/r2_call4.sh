# Round 2, GPU call 4: bring-up of the CTA-pair (cta_group::2) launch.  Every stage under its own timeout.
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
(timeout 120 python tools/pair_probe.py check 2>&1) > gpurun_out/c4_probe_check.log; leg "probe check rc=$?: $(tail -n 3 gpurun_out/c4_probe_check.log | tr '\n' '|')"
(timeout 200 python tools/pair_probe.py time 2>&1) > gpurun_out/c4_probe_time.log; leg "probe time rc=$?: $(tail -n 2 gpurun_out/c4_probe_time.log | tr '\n' '|')"
PT="timeout 900 python -m pytest -m gpu -q --no-header -p no:cacheprovider"
($PT tests --durations=5 2>&1) > gpurun_out/c4_tests.log; leg "whole suite: $(tail -n 1 gpurun_out/c4_tests.log)"
for v in 0 1 0 1; do
  (STYLEMC_HCONV_PAIR=$v timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c4_bench_pair$v.json 2> gpurun_out/c4_bench_pair$v.err); leg "bench pair=$v: $(cut -c1-170 gpurun_out/c4_bench_pair$v.json)"
done
