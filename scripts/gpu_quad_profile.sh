# quadrotor (config 5) evidence: kernel-class times at full size + ncu --set full of one k_backpass_warp launch
TAG=${1:-r02}
mkdir -p gpurun_out
PROBLEM=quad DDP=1 ITERS=12 python scripts/gpu_probe.py 16384 2>&1 | tail -1 > gpurun_out/quad_probe_${TAG}.log
cat gpurun_out/quad_probe_${TAG}.log
PROBLEM=quad DDP=1 CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:'k_backpass_warp' -s 2 -c 1 -f -o gpurun_out/prof_quad_${TAG} \
    python scripts/gpu_probe_q.py 16384 4 > gpurun_out/quad_ncu_${TAG}.log 2>&1
