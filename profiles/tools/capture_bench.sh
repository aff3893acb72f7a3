mkdir -p gpurun_out
python bench.py > gpurun_out/r01_bench_v8.json 2> gpurun_out/r01_bench_v8.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r01_launches_v8.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-rotate > gpurun_out/ncu_launch_v8.log 2>&1; echo "launch rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference_v8.json 2>/dev/null; echo "ref rc=$?"
tail -c 400 gpurun_out/r01_bench_v8.json
