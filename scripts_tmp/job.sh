timeout 300 python bench.py > gpurun_out/bench_r1b.json 2>gpurun_out/bench_r1b.err; tail -1 gpurun_out/bench_r1b.json | cut -c1-300
timeout 300 python bench.py --workload fno3d_cylinder_64x128_rollout10 --steps 5 --no-cpu-baseline > gpurun_out/bench_r1b_3d.json 2>/dev/null; tail -1 gpurun_out/bench_r1b_3d.json | cut -c1-200
timeout 300 python bench.py --workload fno3d_combustion_128x128x64_rollout10 --steps 3 --no-cpu-baseline > gpurun_out/bench_r1b_c4.json 2>/dev/null; tail -1 gpurun_out/bench_r1b_c4.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --n-auto 2 --e2e-steps 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"tc_|lmul|modes" -s 30 -c 22 -o gpurun_out/prof_r1b_all python bench.py --steps 1 --n-auto 2 --e2e-steps 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
