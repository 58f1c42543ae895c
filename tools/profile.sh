# ncu evidence for one find_direction step (run under gpurun): launch list + full capture of the top kernel
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 16 --micro-batch 16 --no-cpu-baseline --profile-step"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm_kernel --launch-skip 38 -c 3 -o gpurun_out/prof_igemm $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
