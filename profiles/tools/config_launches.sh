set -x
mkdir -p gpurun_out
for C in C2 C3; do
  timeout 110 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/r02_launches_$C.csv python bench.py --config $C --steps 1 --warmup 1 --no-cpu-baseline --no-e2e \
    > gpurun_out/ncu_launch_$C.log 2>&1; echo "$C rc=$?"
  wc -l gpurun_out/r02_launches_$C.csv
done
