# Round-end ncu evidence for one find_direction step at the bench configuration (run under gpurun, one GPU):
#   1. launch list (gpu__time_duration per launch)            -> gpurun_out/launches.csv      (tools/summarize_profiles.py <tag>)
#   2. per-launch counters of the conv kernels                -> gpurun_out/counters_hconv.csv (tools/counters_md.py)
#   3. per-launch counters of the HBM-bound glue kernels      -> gpurun_out/counters_mem.csv
#   4. ncu --set full of the heaviest shape (1024 px conv1)   -> gpurun_out/prof_top.ncu-rep + raw csv
mkdir -p gpurun_out
T0=$(date +%s)
leg() { echo "== [$(( $(date +%s) - T0 ))s] $*"; }
CMD="python bench.py --steps 1 --warmup 1 --batch ${BATCH:-64} --micro-batch ${BATCH:-64} --no-cpu-baseline --profile-step"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed
timeout 200 $CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
leg "plain run ok"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
leg "launch list: $(grep -c hconv_kernel gpurun_out/launches.csv) hconv rows"
timeout 400 ncu --profile-from-start off --clock-control none -k regex:hconv_kernel --metrics $M --csv --log-file gpurun_out/counters_hconv.csv $CMD > gpurun_out/ncu2.log 2>&1
leg "hconv counters: $(wc -l < gpurun_out/counters_hconv.csv) lines"
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base mangled -k 'regex:.*hconv_kernelILi32ELi32ELi2E.*' -c 1 -f -o gpurun_out/prof_top $CMD > gpurun_out/ncu4.log 2>&1
ncu -i gpurun_out/prof_top.ncu-rep --page raw --csv > gpurun_out/prof_top_raw.csv 2>/dev/null
leg "top kernel capture: $(ls -la gpurun_out/prof_top.ncu-rep 2>&1 | cut -c1-80)"
timeout 400 ncu --profile-from-start off --clock-control none -k 'regex:fir_|act_bwd|img_finish|torgb|resample|attention|layernorm|quickgelu|upfirdn' --metrics $M --csv --log-file gpurun_out/counters_mem.csv $CMD > gpurun_out/ncu3.log 2>&1
leg "glue counters: $(wc -l < gpurun_out/counters_mem.csv) lines"
