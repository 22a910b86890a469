timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
cat gpurun_out/torch_gpu_path.json; echo
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_final.json 2>gpurun_out/bench_final.err; tail -1 gpurun_out/bench_final.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('2D', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['whole_step']['frac_of_hbm_peak'], d['clocks'], d['gpu_launches'])"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
