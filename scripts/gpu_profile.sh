# ncu evidence for one round: launch list (whole short bench) + full capture of the hot kernels of one pass.
# usage: bash scripts/gpu_profile.sh <tag>     (run under gpurun; results land in gpurun_out/)
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --batch 262144 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(derivs|backpass|ls_round)' -s 50 -c 6 -f -o gpurun_out/prof_${TAG} \
    python bench.py --batch 32768 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
