# quadrotor backward-pass variants (development aid): lanes per problem
for LPP in 32 16 8; do
  echo "LPP=$LPP"
  ILQG_CW_LPP=$LPP PROBLEM=quad DDP=1 ITERS=${ITERS:-12} python scripts/gpu_probe.py ${B:-16384} 2>&1 | tail -1 | cut -c1-420
done
