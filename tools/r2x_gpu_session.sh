# round-2 session X (1 GPU): pipelined shared-memory halo (GAB1_HALO_PIPE) + 16-byte snapshot stores: GPU tests, then A/B against the shuffle halo
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_gpu_tests.log 2>&1; tail -4 gpurun_out/r2x_gpu_tests.log
AB_TAG=pipe timeout 300 python tools/ab_halo.py 2>&1 | tee gpurun_out/r2x_ab_pipe.txt | tail -12
AB_TAG=shfl GAB1PDE_LIB=tools/_build/libgab1pde_nopipe.so timeout 300 python tools/ab_halo.py 2>&1 | tee gpurun_out/r2x_ab_shfl.txt | tail -12
