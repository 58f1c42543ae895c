# ncu launch list of one find_direction step (run under gpurun); summarise with tools/summarize_profiles.py <tag>
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch ${BATCH:-64} --micro-batch ${MB:-64} --no-cpu-baseline --profile-step"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
