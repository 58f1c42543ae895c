# Round 2, multi-GPU legs (run under `gpurun --gpus N`): weak scaling at the headline configuration (64 seeds / GPU at 1024 px) and
# BASELINE configs[2] (256 px, 129 seeds per step sharded over the ranks: ragged shards, strong scaling).  N = $1.
N=${1:-2}
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
(timeout 400 $RUN bench.py --gpus $N --no-cpu-baseline > gpurun_out/m${N}_weak.json 2> gpurun_out/m${N}_weak.err); leg "weak N=$N: $(cut -c1-200 gpurun_out/m${N}_weak.json)"
(timeout 400 $RUN bench.py --gpus $N --resolution 256 --global-seeds 129 --no-cpu-baseline > gpurun_out/m${N}_strong.json 2> gpurun_out/m${N}_strong.err); leg "256px/129 N=$N: $(cut -c1-200 gpurun_out/m${N}_strong.json)"
