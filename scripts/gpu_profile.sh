# ncu evidence for one round (run under gpurun; results land in gpurun_out/):
#   1. launch list of the whole default-size bench process (shares of the step per kernel)
#   2. per-launch time + DRAM bytes of every hot kernel of four passes at the launch size bench.py times kernels at
#   3. ncu --set full of the hot kernels of one pass (stall reasons, pipes, occupancy) at 32768 problems per launch (latency build of
#      the backward pass, 2 sequential line-search rounds) and at 65536 (throughput build, 3 rounds)
# usage: bash scripts/gpu_profile.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --batch 262144 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
CHUNKS=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none -k regex:'k_(derivs|backpass|ls_round|ls_tail|ls_commit)' -s 40 -c 33 --csv --log-file gpurun_out/traffic65k_${TAG}.csv \
    python scripts/gpu_probe2.py 65536 12 > gpurun_out/probe2_${TAG}.log 2>&1
CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:'k_(derivs|backpass|ls_round|ls_tail|ls_commit)' -s 28 -c 7 -f -o gpurun_out/prof_${TAG} \
    python scripts/gpu_probe2.py 32768 8 > gpurun_out/probe2_full_${TAG}.log 2>&1
CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:'k_(derivs|backpass|ls_round|ls_tail|ls_commit)' -s 32 -c 8 -f -o gpurun_out/prof65k_${TAG} \
    python scripts/gpu_probe2.py 65536 8 > gpurun_out/probe2_full65k_${TAG}.log 2>&1
ls -la gpurun_out
