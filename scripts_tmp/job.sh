timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_torch_path_timing.py 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('2D', d['value'], d['ms_per_step'], d['stages_ms_per_rollout'])"
timeout 300 python bench.py --workload fno3d_cylinder_64x128_rollout10 --steps 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('3D', d['value'], d['ms_per_step'], d['stages_ms_per_rollout'])"
timeout 300 python bench_train.py --steps 10 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train', d['value'], d['ms_per_step'], d['phases_ms'])"
