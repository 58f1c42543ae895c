# Round-end GPU check (run under gpurun, one GPU): smoke, the whole `-m gpu` suite, both bench arms (+ the library baseline), the other bench
# workloads, the op microbenches and the launch list of the final build.  Everything lands in gpurun_out/rc_*; each leg has its own timeout.
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
(timeout 300 python __graft_entry__.py smoke 2>&1) > gpurun_out/rc_smoke.log; leg "smoke: $(tail -n 1 gpurun_out/rc_smoke.log)"
(timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --durations=8 2>&1) > gpurun_out/rc_tests.log; leg "tests: $(tail -n 1 gpurun_out/rc_tests.log)"
(timeout 400 python bench.py --gpu-library-baseline > gpurun_out/rc_bench.json 2> gpurun_out/rc_bench.err); leg "bench: $(cut -c1-200 gpurun_out/rc_bench.json)"
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/rc_bench_ref.json 2> gpurun_out/rc_bench_ref.err); leg "reference arm: $(cut -c1-160 gpurun_out/rc_bench_ref.json)"
(timeout 200 python bench.py --workload generate_fromS > gpurun_out/rc_bench_gen.json 2> gpurun_out/rc_bench_gen.err); leg "generate_fromS: $(cut -c1-160 gpurun_out/rc_bench_gen.json)"
(timeout 300 python bench.py --clip-type double --no-cpu-baseline > gpurun_out/rc_bench_double.json 2> gpurun_out/rc_bench_double.err); leg "clip_type double: $(cut -c1-160 gpurun_out/rc_bench_double.json)"
(timeout 200 python bench.py --resolution 256 --global-seeds 129 --no-cpu-baseline > gpurun_out/rc_bench_256.json 2> gpurun_out/rc_bench_256.err); leg "256 px / 129 seeds: $(cut -c1-160 gpurun_out/rc_bench_256.json)"
(timeout 200 python tools/op_bench.py > gpurun_out/rc_ops.md 2> gpurun_out/rc_ops.err); leg "op bench rows: $(wc -l < gpurun_out/rc_ops.md)"
(timeout 400 python tools/ops_vs_cudnn.py > gpurun_out/rc_ops_vs_cudnn.md 2> gpurun_out/rc_ops_vs_cudnn.err); leg "ops vs cudnn rows: $(wc -l < gpurun_out/rc_ops_vs_cudnn.md)"
# launch list of the same build (tools/summarize_profiles.py <tag> turns gpurun_out/launches.csv into profiles/<tag>_launches.md)
CMD="python bench.py --steps 1 --warmup 1 --batch 64 --micro-batch 64 --no-cpu-baseline --profile-step"
(timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1); leg "launch list: $(grep -c hconv_kernel gpurun_out/launches.csv) hconv rows"
