# A/B a kernel switch inside ONE gpurun call (same box): AB_ENV=STYLEMC_FIR_ACT3 AB_VALS="0 3 0 3" bash tools/gpu_ab.sh
mkdir -p gpurun_out
for t in ${TESTS-synthesis step}; do
  (timeout 600 python -m pytest tests/test_${t}_gpu.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1) > gpurun_out/q_$t.log
  echo "== $t: $(tail -n 1 gpurun_out/q_$t.log)"
done
i=0
for v in ${AB_VALS:-0 1 0 1}; do
  i=$((i+1))
  env ${AB_ENV}=$v python bench.py --no-cpu-baseline --steps ${STEPS:-6} ${BENCH_ARGS} > gpurun_out/bench_ab$i.log 2>gpurun_out/bench_ab$i.err
  python - $i $v <<'PY'
import json, sys
i, v = sys.argv[1:3]
try:
    d = json.loads(open(f'gpurun_out/bench_ab{i}.log').read().strip().splitlines()[-1])
    print(f'{v}: bench', d['value'], 'img/s', d['ms_per_step'], 'ms; clocks', d['clocks']['sm_mhz'], '; family share', d['roofline']['family']['share_of_step'])
except Exception as e:
    print('bench failed', e); print(open(f'gpurun_out/bench_ab{i}.err').read()[-2000:])
PY
done
