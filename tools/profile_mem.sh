# ncu --set full on the largest launches of the HBM-bound kernels of one find_direction step (run under gpurun)
# usage: CAPS="name:regex:skip:count ..." bash tools/profile_mem.sh
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch ${BATCH:-16} --micro-batch ${BATCH:-16} --no-cpu-baseline --profile-step"
for cap in ${CAPS:-fir_act3_grad:fir_act3:7:1 fir_act3_nograd:fir_act3:15:1}; do
  IFS=: read name regex skip count <<< "$cap"
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$regex" --launch-skip $skip -c $count -o gpurun_out/prof_$name -f $CMD > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv > gpurun_out/prof_${name}_source.csv 2>/dev/null
  tail -1 gpurun_out/ncu_$name.log
done
du -sh gpurun_out
