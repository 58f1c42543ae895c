# Round 2, GPU call 2: new tests, bench legs (default + gpu library baseline, generate_fromS, 256 px / 129 seeds), cuDNN comparison,
# the original-branch precision diagnostic, launch list and --set full captures of hconv_kernel<128,64,1> (source-level: where the MMA warp waits).
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
PT="python -m pytest -m gpu -q --no-header -p no:cacheprovider"
(timeout 900 $PT tests -x --durations=8 2>&1) > gpurun_out/c2_tests.log; leg "whole suite: $(tail -n 1 gpurun_out/c2_tests.log)"
(timeout 400 python bench.py --gpu-library-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err); leg "bench: $(cut -c1-200 gpurun_out/c2_bench.json)"
(timeout 200 python bench.py --workload generate_fromS > gpurun_out/c2_bench_gen.json 2> gpurun_out/c2_bench_gen.err); leg "generate_fromS: $(cut -c1-200 gpurun_out/c2_bench_gen.json)"
(timeout 200 python bench.py --resolution 256 --global-seeds 129 --no-cpu-baseline > gpurun_out/c2_bench_256.json 2> gpurun_out/c2_bench_256.err); leg "256px/129 seeds: $(cut -c1-200 gpurun_out/c2_bench_256.json)"
(timeout 400 python tools/ops_vs_cudnn.py > gpurun_out/c2_ops_vs_cudnn.md 2> gpurun_out/c2_ops_vs_cudnn.err); leg "ops vs cudnn rows: $(wc -l < gpurun_out/c2_ops_vs_cudnn.md)"
(timeout 300 python tests/diag/diag_original_branch.py > gpurun_out/c2_diag_orig.log 2>&1); leg "original-branch diag: $(wc -l < gpurun_out/c2_diag_orig.log) lines"
CMD="python bench.py --steps 1 --warmup 1 --batch 64 --micro-batch 64 --no-cpu-baseline --profile-step"
(timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1); leg "launch list: $(grep -c hconv_kernel gpurun_out/launches.csv) hconv rows"
for spec in "10 c256" "6 c64" "169 dgrad"; do
  set -- $spec
  (timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base mangled -k 'regex:.*hconv_kernelILi128ELi64ELi1E.*' --launch-skip $1 --launch-count 1 -f -o gpurun_out/c2_full_$2 $CMD > gpurun_out/ncu_full_$2.log 2>&1)
  leg "full capture $2: $(ls -la gpurun_out/c2_full_$2.ncu-rep 2>&1 | cut -c1-80)"
done
