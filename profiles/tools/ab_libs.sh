# ABAB of library variants (lattigo-fhe-by-go_b200/lib/variants/<name>.so): bash profiles/tools/ab_libs.sh <out.jsonl> <rounds> <name> <name> ...
OUT=$1; R=$2; shift 2
for r in $(seq 1 $R); do for v in "$@"; do
  LATTIGPU_LIB=$PWD/lattigo-fhe-by-go_b200/lib/variants/$v.so python profiles/tools/ntt_l2_sweep.py time 10 | tee -a $OUT
done; done
