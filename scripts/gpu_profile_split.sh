M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for SP in 0 4; do
ILQG_BP_SPLIT=$SP CHUNKS=1 ncu --metrics $M --clock-control none -k regex:'k_backpass' -c 46 --csv --log-file gpurun_out/bp4096_split${SP}.csv python scripts/gpu_probe2.py 4096 45 > /dev/null 2>&1
done
ILQG_BP_SPLIT=4 CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:'k_backpass' -s 3 -c 1 -f -o gpurun_out/split4096_pass3 python scripts/gpu_probe2.py 4096 6 > /dev/null 2>&1
ILQG_BP_SPLIT=4 CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:'k_backpass' -s 40 -c 1 -f -o gpurun_out/split4096_pass40 python scripts/gpu_probe2.py 4096 42 > /dev/null 2>&1
ls -la gpurun_out | tail -5
