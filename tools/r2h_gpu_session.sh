# round-2 session H (1 GPU): ncu evidence — launch list of bench.py, full capture of the headline kernel on the bench workload,
# full capture of the large-grid gang kernel (Nr = 200)
set -x
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/r2h_bench_noe2e.json 2> gpurun_out/r2h_bench_noe2e.err; cat gpurun_out/r2h_bench_noe2e.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/r2h_ncu_launches.log 2>&1; tail -2 gpurun_out/r2h_ncu_launches.log
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o gpurun_out/r2_solve_kernel_config2 python bench.py --steps 1 --warmup 1 --no-e2e > gpurun_out/r2h_ncu_full.log 2>&1; tail -2 gpurun_out/r2h_ncu_full.log
export GAB1_KERNEL=gang
ncu --set full --clock-control none --import-source on -k regex:gang_kernel -c 1 -f -o gpurun_out/r2_gang_16_13_nr200 python tools/prof_one.py 592 0.05 0.05 > gpurun_out/r2h_ncu_gang200.log 2>&1; tail -1 gpurun_out/r2h_ncu_gang200.log
ncu --set full --clock-control none --import-source on -k regex:gang_kernel -c 1 -f -o gpurun_out/r2_gang_32_13_nr400 python tools/prof_one.py 592 0.01 0.025 > gpurun_out/r2h_ncu_gang400.log 2>&1; tail -1 gpurun_out/r2h_ncu_gang400.log
ls -la gpurun_out/*.ncu-rep
