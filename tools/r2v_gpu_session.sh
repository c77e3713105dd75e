# round-2 session V (8 GPUs): the multi-device test that failed in session U, with details
set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "scatter" 2>&1 | tail -40
GAB1_DUO=0 timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "scatter" 2>&1 | tail -3
