# Round-end GPU check (run under gpurun, one GPU): smoke, the whole `-m gpu` suite, both bench arms, the generate_fromS workload and the
# op microbench.  Everything lands in gpurun_out/; each leg has its own timeout so one hang cannot eat the call.
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
(timeout 300 python __graft_entry__.py smoke 2>&1) > gpurun_out/rc_smoke.log; leg "smoke: $(tail -n 1 gpurun_out/rc_smoke.log)"
(timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider --durations=8 2>&1) > gpurun_out/rc_tests.log; leg "tests: $(tail -n 1 gpurun_out/rc_tests.log)"
(timeout 300 python bench.py > gpurun_out/rc_bench.json 2> gpurun_out/rc_bench.err); leg "bench: $(cut -c1-200 gpurun_out/rc_bench.json)"
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/rc_bench_ref.json 2> gpurun_out/rc_bench_ref.err); leg "reference arm: $(cut -c1-160 gpurun_out/rc_bench_ref.json)"
(timeout 200 python bench.py --workload generate_fromS > gpurun_out/rc_bench_gen.json 2> gpurun_out/rc_bench_gen.err); leg "generate_fromS: $(cut -c1-160 gpurun_out/rc_bench_gen.json)"
(timeout 200 python tools/op_bench.py > gpurun_out/rc_ops.md 2> gpurun_out/rc_ops.err); leg "op bench rows: $(wc -l < gpurun_out/rc_ops.md)"
# launch list of the same build (tools/summarize_profiles.py <tag> turns gpurun_out/launches.csv into profiles/<tag>_launches.md)
CMD="python bench.py --steps 1 --warmup 1 --batch 64 --micro-batch 64 --no-cpu-baseline --profile-step"
(timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1); leg "launch list: $(grep -c hconv_kernel gpurun_out/launches.csv) hconv rows"
