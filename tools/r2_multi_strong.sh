# BASELINE configs[2] under torchrun with and without CUDA-graph replay.  N = $1.
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for g in 1 0; do
  (timeout 300 $RUN bench.py --gpus $N --resolution 256 --global-seeds 129 --no-cpu-baseline --cuda-graph $g > gpurun_out/m${N}_strong_g$g.json 2> gpurun_out/m${N}_strong_g$g.err)
  echo "256px/129 N=$N graph=$g: $(grep '^{' gpurun_out/m${N}_strong_g$g.json | cut -c1-200)"
done
