# The parity suites under the A/B switches that select the older kernels (every variant must return the same words):
#   bash profiles/tools/switch_matrix.sh
S="tests/test_gpu_ring.py tests/test_gpu_ckks.py tests/test_gpu_fullsize.py tests/test_gpu_graph.py tests/test_gpu_multi.py"
LATTIGPU_NO_STRIDED_TMA=1 LATTIGPU_NO_KS_TMA=1 LATTIGPU_NO_AUX_STREAMS=1 python -m pytest $S -m gpu -x -q 2>&1 | tail -2
LATTIGPU_TILE_FASTEST=0 LATTIGPU_NO_FP_MAC=1 LATTIGPU_MODUP_CPT2=1 python -m pytest $S -m gpu -x -q 2>&1 | tail -2
LATTIGPU_NO_D64_NTT=1 LATTIGPU_NO_FP_MODUP=1 python -m pytest $S -m gpu -x -q 2>&1 | tail -2
