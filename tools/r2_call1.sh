# Round 2, GPU call 1 (one GPU): the new 1024-px golden tests, the whole suite, compute-sanitizer memcheck / racecheck over small
# tests of every .cu file, then the default bench and --clip-type double.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
PT="python -m pytest -m gpu -q --no-header -p no:cacheprovider"
(timeout 300 python __graft_entry__.py smoke 2>&1) > gpurun_out/r2_smoke.log; leg "smoke: $(tail -n 1 gpurun_out/r2_smoke.log)"
(timeout 600 $PT tests/test_step_gpu.py tests/test_generate_gpu.py -s -k "config4 or 1024" 2>&1) > gpurun_out/r2_config4.log; leg "1024-px goldens: $(tail -n 1 gpurun_out/r2_config4.log)"
grep -E "config4|canvas" gpurun_out/r2_config4.log | head -40
(timeout 900 $PT tests --durations=10 2>&1) > gpurun_out/r2_tests.log; leg "whole suite: $(tail -n 1 gpurun_out/r2_tests.log)"
SAN="compute-sanitizer --error-exitcode 1 --print-limit 20"
(timeout 500 $SAN --tool memcheck $PT tests/test_zz_clip_b16_gpu.py tests/test_clip_gpu.py -k "attention or clip_loss or unprocess_golden or encode_golden" 2>&1) > gpurun_out/r2_memcheck_vit.log
leg "memcheck vit.cu: $(grep 'ERROR SUMMARY' gpurun_out/r2_memcheck_vit.log | tail -n 1) / $(grep -E 'passed|failed' gpurun_out/r2_memcheck_vit.log | tail -n 1)"
(timeout 500 $SAN --tool memcheck $PT tests/test_synthesis_gpu.py tests/test_ops_gpu.py -k "golden" 2>&1) > gpurun_out/r2_memcheck_synth.log
leg "memcheck synth.cu/hconv.cu/upfirdn2d.cu/bias_act.cu: $(grep 'ERROR SUMMARY' gpurun_out/r2_memcheck_synth.log | tail -n 1) / $(grep -E 'passed|failed' gpurun_out/r2_memcheck_synth.log | tail -n 1)"
(timeout 400 $SAN --tool racecheck $PT tests/test_zz_clip_b16_gpu.py -k "attention_197 or tiled_attention" 2>&1) > gpurun_out/r2_racecheck_vit.log
leg "racecheck attention: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/r2_racecheck_vit.log | tail -n 1) / $(grep -E 'passed|failed' gpurun_out/r2_racecheck_vit.log | tail -n 1)"
(timeout 400 $SAN --tool racecheck $PT tests/test_synthesis_gpu.py -k "generate_image_golden" 2>&1) > gpurun_out/r2_racecheck_synth.log
leg "racecheck synthesis fwd: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/r2_racecheck_synth.log | tail -n 1) / $(grep -E 'passed|failed' gpurun_out/r2_racecheck_synth.log | tail -n 1)"
(timeout 300 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err); leg "bench: $(cut -c1-300 gpurun_out/r2_bench.json)"
(timeout 400 python bench.py --clip-type double --no-cpu-baseline > gpurun_out/r2_bench_double.json 2> gpurun_out/r2_bench_double.err); leg "bench clip_type=double: $(cut -c1-200 gpurun_out/r2_bench_double.json)"
