mkdir -p gpurun_out
for t in igemm hconv ops synthesis clip step generate; do
  (timeout 600 python -m pytest tests/test_${t}_gpu.py -m gpu -q -s --no-header -p no:cacheprovider 2>&1) > gpurun_out/r4_$t.log
  echo "== $t: $(tail -n 1 gpurun_out/r4_$t.log)"
done
