timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r1c_bench_8gpu.json 2> gpurun_out/r1c_bench_8gpu.err
tail -c 800 gpurun_out/r1c_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1c_bench_8gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['n_gpus'], d['config']['collective'][:60], d['config']['scale_exchange_matches_nccl_all_gather'], d['config']['peer_mapping_error'])
PY
