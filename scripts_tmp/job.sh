timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_epi2.json 2>gpurun_out/bench_epi2.err; cat gpurun_out/bench_epi2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stages_ms_per_rollout'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
