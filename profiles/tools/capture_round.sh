# End-of-round evidence: GPU tests, bench line, launch list, ncu --set full of one step and of the NTT launch pair,
# the other BASELINE configurations.  Run on the GPU box from the repo root; outputs land in gpurun_out/.
set -x
V=${1:-v9}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r01_gpu_tests_$V.log 2>&1; echo "tests rc=$?"
python bench.py > gpurun_out/r01_bench_$V.json 2> gpurun_out/r01_bench_$V.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r01_launches_$V.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-rotate > gpurun_out/ncu_launch_$V.log 2>&1; echo "launch list rc=$?"
K='regex:ntt_|ks_|modup|tensor_kernel|ew_kernel|range_flags|fanout|permute'
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 31 --launch-count 31 -f -o gpurun_out/step_full_$V \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-rotate > gpurun_out/ncu_full_$V.log 2>&1; echo "full rc=$?"
ncu -i gpurun_out/step_full_$V.ncu-rep --page raw --csv > gpurun_out/step_full_${V}_raw.csv 2>/dev/null
python profiles/tools/ncu_table.py gpurun_out/step_full_${V}_raw.csv > gpurun_out/r01_ncu_full_${V}_summary.txt
timeout 300 ncu --set full --clock-control none -k regex:ntt_fwd_strided\|ntt_contig_pipe --launch-skip 10 --launch-count 2 -f -o gpurun_out/nttfwd_$V \
  python profiles/tools/ntt_time.py 16 34 32 > gpurun_out/ncu_nttfwd_$V.log 2>&1; echo "nttfwd rc=$?"
ncu -i gpurun_out/nttfwd_$V.ncu-rep --page raw --csv > gpurun_out/nttfwd_${V}_raw.csv 2>/dev/null
python profiles/tools/ntt_fwd_traffic.py gpurun_out/nttfwd_${V}_raw.csv 65536 34 32 gpurun_out/r01_ncu_ntt_fwd.json
rm -f gpurun_out/step_full_$V.ncu-rep gpurun_out/nttfwd_$V.ncu-rep
python profiles/tools/bench_configs.py c1 c2 c3 c5 > gpurun_out/r01_configs_$V.jsonl 2> gpurun_out/r01_configs_$V.err; echo "configs rc=$?"
tail -c 700 gpurun_out/r01_bench_$V.json
