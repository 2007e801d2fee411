set -x
timeout 300 python -m pytest tests/test_gpu_algorithms.py tests/test_gpu_requant.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r1c_bench_2gpu.json 2> gpurun_out/r1c_bench_2gpu.err
tail -c 1500 gpurun_out/r1c_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1c_bench_2gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['n_gpus'], d['config'])
PY
AEQB_BENCH_NCCL_GATHER=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nccl', d['value'], d['ms_per_step'])"
