# Round 2, GPU call 3: staged epilogue parameters in hconv.cu -> parity tests, bench, launch list, --set full captures of the small-C kernels
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
PT="python -m pytest -m gpu -q --no-header -p no:cacheprovider"
(timeout 900 $PT tests --durations=5 2>&1) > gpurun_out/c3_tests.log; leg "whole suite: $(tail -n 1 gpurun_out/c3_tests.log)"
(timeout 400 python bench.py --gpu-library-baseline > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err); leg "bench: $(cut -c1-200 gpurun_out/c3_bench.json)"
CMD="python bench.py --steps 1 --warmup 1 --batch 64 --micro-batch 64 --no-cpu-baseline --profile-step"
(timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1); leg "launch list: $(grep -c hconv_kernel gpurun_out/launches.csv) hconv rows"
for spec in "128ELi64ELi1E 10 c256" "32ELi32ELi2E 1 c32" "64ELi64ELi2E 1 c64m" "32ELi64ELi2E 0 c0_1024"; do
  set -- $spec
  (timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:.*hconv_kernelILi$1.*" --launch-skip $2 --launch-count 1 -f -o gpurun_out/c3_full_$3 $CMD > gpurun_out/ncu_full_$3.log 2>&1)
  leg "full capture $3: $(ls -la gpurun_out/c3_full_$3.ncu-rep 2>&1 | cut -c1-80)"
done
