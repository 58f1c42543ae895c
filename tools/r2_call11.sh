mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
PT="timeout 900 python -m pytest -m gpu -q --no-header -p no:cacheprovider"
($PT tests 2>&1) > gpurun_out/c11_tests.log; leg "whole suite: $(tail -n 1 gpurun_out/c11_tests.log)"
(timeout 500 python tools/ops_vs_cudnn.py > gpurun_out/c11_ops_vs_cudnn.md 2> gpurun_out/c11_ops.err); grep -E "^\| b(256|512|1024)" gpurun_out/c11_ops_vs_cudnn.md | cut -d"|" -f2-6,11-12
for i in 1 2; do (timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c11_bench$i.json 2> gpurun_out/c11_bench$i.err); leg "bench: $(cut -c1-170 gpurun_out/c11_bench$i.json)"; done
