# car through the warp-cooperative backward pass (development aid): -DILQG_FORCE_COOP=1 build in lib_coop, lanes per problem swept
for B in ${BATCHES:-4096 16384 32768}; do
  echo "B=$B lane-per-problem kernel"; CHUNKS=1 ITERS=30 python scripts/gpu_probe.py $B 2>&1 | tail -1 | cut -c1-300
  for LPP in 32 16 8; do
    echo "B=$B coop LPP=$LPP"
    ILQG_LIB_DIR=$PWD/ddp-generator_b200/lib_coop ILQG_CW_LPP=$LPP CHUNKS=1 ITERS=30 python scripts/gpu_probe.py $B 2>&1 | tail -1 | cut -c1-300
  done
done
