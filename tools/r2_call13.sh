mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
PT="timeout 900 python -m pytest -m gpu -q --no-header -p no:cacheprovider"

for g in 0 1 0 1; do
  (timeout 300 python bench.py --no-cpu-baseline --cuda-graph $g > gpurun_out/c13_bench_g$g.json 2> gpurun_out/c13_bench_g$g.err); leg "bench 1024 graph=$g: $(cut -c1-170 gpurun_out/c13_bench_g$g.json)"
done
for g in 0 1; do
  (timeout 300 python bench.py --no-cpu-baseline --resolution 256 --batch 17 --micro-batch 17 --cuda-graph $g > gpurun_out/c13_b17_g$g.json 2> gpurun_out/c13_b17_g$g.err); leg "bench 256px batch 17 graph=$g: $(cut -c1-170 gpurun_out/c13_b17_g$g.json)"
done
