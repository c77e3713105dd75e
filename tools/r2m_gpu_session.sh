# round-2 session M (1 GPU): GPU tests at the current tree (certified census included), smoke
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2m_gpu_tests.log 2>&1; tail -5 gpurun_out/r2m_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ls gpurun_out | grep census
