# ncu --set full of ONE kernel of the MulRelin+Rescale step: table line + SASS stall samples
#   bash profiles/tools/ncu_one.sh <kernel-regex> <tag> [launch-skip]
K=$1; TAG=$2; SKIP=${3:-1}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 --launch-skip $SKIP \
  -f -o gpurun_out/one_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-rotate > gpurun_out/ncu_one_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/one_$TAG.ncu-rep --page raw --csv > gpurun_out/one_${TAG}_raw.csv 2>/dev/null
python profiles/tools/ncu_table.py gpurun_out/one_${TAG}_raw.csv | tee gpurun_out/r02_ncu_$TAG.txt | cut -c1-330
ncu -i gpurun_out/one_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/src_$TAG.csv 2>/dev/null
python profiles/tools/sass_stalls.py gpurun_out/src_$TAG.csv | tee gpurun_out/r02_stalls_$TAG.txt
rm -f gpurun_out/one_$TAG.ncu-rep
