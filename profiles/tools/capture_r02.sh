# Round-2 evidence: GPU tests, bench line, launch list, ncu --set full of one step (table) and the SASS-level stall samples
# of the three heaviest kernels.  Run on the GPU box from the repo root; outputs land in gpurun_out/.
#   bash profiles/tools/capture_r02.sh v5 [notests]
set -x
V=${1:-v5}
mkdir -p gpurun_out
if [ "$2" != "notests" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_$V.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02_gpu_tests_$V.log
fi
python bench.py > gpurun_out/r02_bench_$V.json 2> gpurun_out/r02_bench_$V.err; echo "bench rc=$?"
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-rotate"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_$V.csv \
  $B > gpurun_out/ncu_launch_$V.log 2>&1; echo "launch list rc=$?"
python profiles/tools/step_breakdown.py gpurun_out/r02_launches_$V.csv > gpurun_out/r02_step_breakdown_$V.txt; cat gpurun_out/r02_step_breakdown_$V.txt
K='regex:ntt_|ks_|modup|tensor_kernel|ew_kernel|range_flags|fanout|permute'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 32 --launch-count 32 -f -o gpurun_out/step_full_$V \
  $B > gpurun_out/ncu_full_$V.log 2>&1; echo "full rc=$?"
ncu -i gpurun_out/step_full_$V.ncu-rep --page raw --csv > gpurun_out/step_full_${V}_raw.csv 2>/dev/null
python profiles/tools/ncu_table.py gpurun_out/step_full_${V}_raw.csv > gpurun_out/r02_ncu_full_${V}_summary.txt
cat gpurun_out/r02_ncu_full_${V}_summary.txt | cut -c1-250
python profiles/tools/step_traffic.py gpurun_out/r02_ncu_full_${V}_summary.txt 32 65536 gpurun_out/r02_step_traffic.json; cat gpurun_out/r02_step_traffic.json
timeout 300 ncu --set full --clock-control none -k regex:ntt_fwd_strided\|ntt_contig_pipe --launch-skip 10 --launch-count 2 -f -o gpurun_out/nttfwd_$V \
  python profiles/tools/ntt_time.py 16 34 32 > gpurun_out/ncu_nttfwd_$V.log 2>&1; echo "nttfwd rc=$?"
ncu -i gpurun_out/nttfwd_$V.ncu-rep --page raw --csv > gpurun_out/nttfwd_${V}_raw.csv 2>/dev/null
python profiles/tools/ntt_fwd_traffic.py gpurun_out/nttfwd_${V}_raw.csv 65536 34 32 gpurun_out/r02_ncu_ntt_fwd.json; cat gpurun_out/r02_ncu_ntt_fwd.json
rm -f gpurun_out/nttfwd_$V.ncu-rep
for sel in ks_fused_tma_kernel:0 ks_fused_kernel:0 ntt_fwd_strided_tma:0 ntt_contig_pipe:0 ntt_contig_pipe:1 ntt_contig_pipe:2 ntt_contig_pipe:3 modup_fp_kernel:0 ntt_inv_strided:0; do
  kn=${sel%%:*}; sk=${sel##*:}
  ncu -i gpurun_out/step_full_$V.ncu-rep --page source --csv --print-source sass -k regex:$kn --launch-skip $sk --launch-count 1 \
    > gpurun_out/src_${kn}_${sk}_$V.csv 2>/dev/null
  python profiles/tools/sass_stalls.py gpurun_out/src_${kn}_${sk}_$V.csv > gpurun_out/r02_stalls_${kn}_${sk}_$V.txt 2>&1
  cat gpurun_out/r02_stalls_${kn}_${sk}_$V.txt
done
rm -f gpurun_out/step_full_$V.ncu-rep
tail -c 600 gpurun_out/r02_bench_$V.json
