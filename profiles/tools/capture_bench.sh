# final-code capture: GPU tests, bench line, launch list (see capture_round.sh for the ncu --set full tables)
V=${1:-v9}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r01_gpu_tests_$V.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/r01_gpu_tests_$V.log
python bench.py > gpurun_out/r01_bench_$V.json 2> gpurun_out/r01_bench_$V.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r01_launches_$V.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-rotate > gpurun_out/ncu_launch_$V.log 2>&1; echo "launch rc=$?"
tail -c 300 gpurun_out/r01_bench_$V.json
