# round-2 session Z6 (1 GPU): ncu --set full of the headline kernel at the final tree on 20 000 prior draws (bench settings)
set -x
timeout 150 ncu --set full --clock-control none --import-source on -k regex:^solve_kernel -c 1 -f -o gpurun_out/r2z_solve_kernel_k2_prior20000 python tools/prof_prior.py 20000 > gpurun_out/r2z_ncu_full.log 2>&1; tail -3 gpurun_out/r2z_ncu_full.log
ls -la gpurun_out/r2z_solve_kernel_k2_prior20000.ncu-rep
