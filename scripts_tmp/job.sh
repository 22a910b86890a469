timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_torch_path_timing.py 2>&1 | tail -3
timeout 200 python bench.py --workload fno2d_modes12_256x256 --steps 20 --e2e-steps 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), d['stages_ms_per_rollout'])"
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('2D', d['value'], d['ms_per_step'], d['stages_ms_per_rollout'])"
