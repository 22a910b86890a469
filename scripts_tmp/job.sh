timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_torch_path_timing.py 2>&1 | tail -3
timeout 300 python bench_train.py --steps 10 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases_ms'])"
timeout 300 python bench_train.py --workload fno3d_cylinder_64x128_train --steps 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases_ms'], d['config']['batch_per_gpu'])"
