timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_torch_path_timing.py 2>&1 | tail -3
rm -f gpurun_out/c5_sweep3.jsonl
for k in 12 16 24 32 48 64; do timeout 200 python bench.py --workload fno2d_modes${k}_256x256 --steps 20 --e2e-steps 2 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/c5_sweep3.jsonl; done
python - <<'PY'
import json
for l in open('gpurun_out/c5_sweep3.jsonl'):
    d=json.loads(l); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['whole_step']['frac_of_hbm_peak'],3), d['stages_ms_per_rollout'])
PY
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('2D', d['value'], d['ms_per_step'], d['stages_ms_per_rollout']['fwdW'])"
