# Round 2, GPU call 9: whole suite + round profile of the current build (launch list, hconv counters, glue counters, top-kernel capture)
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
PT="timeout 900 python -m pytest -m gpu -q --no-header -p no:cacheprovider"
($PT tests --durations=8 2>&1) > gpurun_out/c9_tests.log; leg "whole suite: $(tail -n 1 gpurun_out/c9_tests.log)"
(timeout 400 python bench.py --gpu-library-baseline > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err); leg "bench: $(cut -c1-200 gpurun_out/c9_bench.json)"
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c9_bench_ref.json 2> gpurun_out/c9_bench_ref.err); leg "reference arm: $(cut -c1-160 gpurun_out/c9_bench_ref.json)"
(timeout 200 python bench.py --workload generate_fromS > gpurun_out/c9_bench_gen.json 2> gpurun_out/c9_bench_gen.err); leg "generate_fromS: $(cut -c1-200 gpurun_out/c9_bench_gen.json)"
(timeout 300 python bench.py --clip-type double --no-cpu-baseline > gpurun_out/c9_bench_double.json 2> gpurun_out/c9_bench_double.err); leg "double: $(cut -c1-200 gpurun_out/c9_bench_double.json)"
(timeout 200 python bench.py --resolution 256 --global-seeds 129 --no-cpu-baseline > gpurun_out/c9_bench_256.json 2> gpurun_out/c9_bench_256.err); leg "256px/129 seeds: $(cut -c1-200 gpurun_out/c9_bench_256.json)"
