timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_torch_path_timing.py 2>&1 | tail -3
timeout 300 python bench.py 2>gpurun_out/bench_pdl.err | tail -1 > gpurun_out/bench_pdl.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_pdl.json').read()); print('PDL   ', d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'], d['e2e']['value'])"
B200FNO_NO_PDL=1 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no PDL', d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
timeout 300 python bench.py --workload fno3d_cylinder_64x128_rollout10 --steps 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_pdl_3d.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_pdl_3d.json').read()); print('3D', d['value'], d['ms_per_step'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
