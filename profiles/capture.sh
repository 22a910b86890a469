#!/usr/bin/env bash
# One gpurun call that regenerates the round's measurements (run from the repo root on the GPU box):
#   gpurun --timeout 1500 -- 'bash profiles/capture.sh r02a'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/ with profiles/summarize.py:
#   python profiles/summarize.py list gpurun_out/<tag>_launches.csv profiles/<tag>_launches.csv
#   python profiles/summarize.py full gpurun_out/<tag>_full.ncu-rep profiles/<tag>_ncu_full.csv
# Numbers printed by the runs under ncu are never bench values.
set -u
tag=${1:-rXX}
o=gpurun_out
mkdir -p $o
run() { echo "== $*" >&2; timeout "$@"; }

run 600 python -m pytest tests -m gpu -x -q > $o/${tag}_gpu_tests.log 2>&1
# headline (C2), 3-D cylinder, C4, surrogate chunk, reference CPU arm
run 400 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err
run 300 python bench.py --workload fno3d_cylinder_64x128_rollout10 --no-cpu-baseline > $o/${tag}_bench_3d.json 2>/dev/null
run 300 python bench.py --workload fno3d_combustion_128x128x64_rollout10 --no-cpu-baseline > $o/${tag}_bench_c4.json 2>/dev/null
run 300 python bench.py --workload fno3d_surrogate_128x128_c17_forward --no-cpu-baseline > $o/${tag}_bench_surrogate.json 2>/dev/null
run 300 python profiles/surrogate_timing.py > $o/${tag}_surrogate_timing.json 2>/dev/null
run 300 python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_bench_ref.json 2>/dev/null
run 300 python bench_train.py > $o/${tag}_train_bench.json 2>/dev/null
run 300 python bench_train.py --dtype bf16 > $o/${tag}_train_bench_bf16.json 2>/dev/null
run 300 python bench.py --dtype bf16 --no-cpu-baseline > $o/${tag}_bench_bf16.json 2>/dev/null
# multi-GPU boxes (gpurun --gpus N): torchrun --nproc-per-node N bench_train.py --gpus N [--dtype bf16]; bench.py --gpus N --workload ...
for k in 12 16 24 32 48 64; do
  run 200 python bench.py --workload fno2d_modes${k}_256x256 --steps 20 --no-cpu-baseline 2>/dev/null | tail -1
done > $o/${tag}_c5_mode_sweep.jsonl
# launch list of one short run, then a full capture of one step's kernels (after the packing / warm-up launches)
run 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $o/${tag}_launches.csv \
  python bench.py --steps 1 --n-auto 2 --e2e-steps 1 --no-cpu-baseline > /dev/null 2>&1
run 900 ncu --set full --clock-control none --import-source on -k regex:'tc_|lmul|modes_kernel' -s 120 -c 24 \
  -o $o/${tag}_full -f python bench.py --steps 1 --n-auto 2 --e2e-steps 1 --no-cpu-baseline > $o/${tag}_ncu.log 2>&1
# the weight-streaming mode-mixing kernel on the 3-D cylinder model (100.7 MB of weights per layer)
run 400 ncu --set full --clock-control none --import-source on -k regex:tc_modes_kernel -s 8 -c 1 -o $o/${tag}_modes3d -f \
  python bench.py --workload fno3d_cylinder_64x128_rollout10 --steps 1 --n-auto 2 --e2e-steps 1 --no-cpu-baseline > $o/${tag}_ncu_modes3d.log 2>&1
ls -la $o | tail -20
