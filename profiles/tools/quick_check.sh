# quick GPU iteration: key-switch parity + a short bench (no CPU baseline / e2e / rotations)
python -m pytest tests/test_gpu_ckks.py tests/test_gpu_bfv.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5
python bench.py --no-cpu-baseline --no-e2e --no-rotate > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/quick_bench.json')); print('ops/s', d['value'], 'ms', d['ms_per_step'], 'fwd', d['ntt']['fwd_limb_ntt_per_s'], 'inv', d['ntt']['inv_limb_ntt_per_s'])"
