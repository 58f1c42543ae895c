# ncu --set full on the largest launches of the HBM-bound kernels of one find_direction step (run under gpurun)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch ${BATCH:-16} --micro-batch ${BATCH:-16} --no-cpu-baseline --profile-step"
cap() {  # name regex skip count
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$2" --launch-skip $3 -c $4 -o gpurun_out/prof_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv > gpurun_out/prof_$1_source.csv 2>/dev/null
  tail -1 gpurun_out/ncu_$1.log
}
cap fir_act2_grad fir_act2 7 1
cap fir_act2_nograd fir_act2 15 1
cap act_bwd1 act_bwd1 0 2
cap fir_bwd2 fir_bwd2 0 1
ls -la gpurun_out
du -sh gpurun_out
