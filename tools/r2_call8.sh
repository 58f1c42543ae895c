# Round 2, GPU call 8: identity loss (IR-SE50 on repo kernels), whole suite, bench
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
PT="timeout 900 python -m pytest -m gpu -q --no-header -p no:cacheprovider"
($PT tests/test_idloss_gpu.py -s -x 2>&1) > gpurun_out/c8_id.log; leg "idloss: $(tail -n 1 gpurun_out/c8_id.log)"; grep -E "^id loss|^loss|feature|Error|^E  " gpurun_out/c8_id.log | head -20
($PT tests --durations=5 2>&1) > gpurun_out/c8_tests.log; leg "whole suite: $(tail -n 1 gpurun_out/c8_tests.log)"
(timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err); leg "bench: $(cut -c1-170 gpurun_out/c8_bench.json)"
