# ncu --set full of the fused key-switch kernel, 96-bit accumulators and LATTIGPU_KS_ACC64=1 (A/B)
for v in 0 1; do
  LATTIGPU_KS_ACC64=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:ks_fused -c 1 --launch-skip 1 \
    -f -o gpurun_out/ks_acc64_$v python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-rotate > gpurun_out/ncu_ks_$v.log 2>&1
  ncu -i gpurun_out/ks_acc64_$v.ncu-rep --page raw --csv > gpurun_out/ks_acc64_${v}_raw.csv 2>/dev/null
  python profiles/tools/ncu_table.py gpurun_out/ks_acc64_${v}_raw.csv
done
