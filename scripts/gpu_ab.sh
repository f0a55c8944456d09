# A/B of the sweep variants + ncu captures of the dominant kernels (1 GPU)
mkdir -p gpurun_out
echo "== fused, 6 CTAs/SM";            timeout 120 python scripts/kernel_times.py normal_smc ma2_smc
echo "== fused, 5 CTAs/SM";            KABC_SWEEP_CTAS=5 timeout 120 python scripts/kernel_times.py normal_smc ma2_smc
echo "== propose + work list + simulate"; KABC_UNFUSED=1 timeout 120 python scripts/kernel_times.py normal_smc ma2_smc
for v in fused unfused; do
  if [ $v = unfused ]; then export KABC_UNFUSED=1; K="k_smc_simulate_list"; else unset KABC_UNFUSED; K="k_smc_sweep"; fi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 10 -c 1 -f -o gpurun_out/prof_normal_smc_$v \
      python scripts/ncu_target.py normal_smc 12 > gpurun_out/ncu_$v.log 2>&1
  cp gpurun_out/ncu_units_normal_smc.json gpurun_out/ncu_units_normal_smc_$v.json
done
unset KABC_UNFUSED
ls -la gpurun_out/*.ncu-rep
