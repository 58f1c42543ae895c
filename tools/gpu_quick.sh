# quick GPU check after an engine change: synthesis/step/hconv parity tests + a short bench (run under gpurun)
mkdir -p gpurun_out
for t in ${TESTS:-hconv synthesis step}; do
  (timeout 600 python -m pytest tests/test_${t}_gpu.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1) > gpurun_out/q_$t.log
  echo "== $t: $(tail -n 1 gpurun_out/q_$t.log)"
done
python bench.py --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_q.log 2>gpurun_out/bench_q.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_q.log').read().strip().splitlines()[-1])
    print('bench', d['value'], 'img/s', d['ms_per_step'], 'ms; e2e', d['e2e']['value'], '; clocks', d['clocks'], '; family share', d['roofline']['family']['share_of_step'], 'top', d['roofline']['kernel'], d['roofline']['avg_launch_ms'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_q.err').read()[-2000:])
PY
