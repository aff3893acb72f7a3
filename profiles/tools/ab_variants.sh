# A/B of library variants under lattigo-fhe-by-go_b200/lib/variants: MulRelin+Rescale ops/s of each
for so in lattigo-fhe-by-go_b200/lib/variants/*.so; do
  LATTIGPU_LIB=$PWD/$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-rotate > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/ab.json')); print('$so', 'ops/s %.1f ms %.3f'%(d['value'], d['ms_per_step']))"
done
