timeout 500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_train.py -x -q 2>&1 | tail -6
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench_train.py --gpus 2 --steps 10 > gpurun_out/train_n2.json 2> gpurun_out/train_n2.err; tail -1 gpurun_out/train_n2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases_ms'], d['allreduce_bytes_per_step'], d['allreduce_overlapped'])"; tail -3 gpurun_out/train_n2.err | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench_train.py --gpus 2 --steps 10 --no-overlap 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases_ms'], d['allreduce_bytes_per_step'], d['allreduce_overlapped'])"
