set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r01_v6_tests.log 2>&1; echo "tests rc=$?"
python bench.py > gpurun_out/r01_bench_v6.json 2> gpurun_out/r01_bench_v6.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference_v6.json 2> gpurun_out/r01_bench_reference_v6.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r01_launches_v6.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-rotate > gpurun_out/r01_ncu_launch_v6.log 2>&1; echo "ncu1 rc=$?"
tail -c 600 gpurun_out/r01_bench_v6.json
