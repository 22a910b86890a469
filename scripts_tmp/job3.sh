for na in 1 2 20; do timeout 300 python bench.py --no-cpu-baseline --n-auto $na --steps 10 --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stages_ms_per_rollout']; n=d['config']['n_autoregressive']; print(n, 'proj per call', s['proj']/n, 'lift', s['lift']/n, 'layer', s['layer']/n/4)"; done
