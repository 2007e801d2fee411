set -x
timeout 300 python -m pytest tests/test_gpu_algorithms.py -x -q -m gpu -k "hadamard or octav" 2>&1 | tail -5
timeout 120 python tools/ktime.py --only hadamard 2>&1 | tail -3 > gpurun_out/r1b_ktime_had.txt
AEQB_HADAMARD_NO_TILES=1 timeout 120 python tools/ktime.py --only hadamard 2>&1 | tail -3 >> gpurun_out/r1b_ktime_had.txt
cat gpurun_out/r1b_ktime_had.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:hadamard_tiles -c 1 -o gpurun_out/r1b_hadamard_tiles -f python tools/ktime.py --only "hadamard_rows n=4096" --iters 2 > gpurun_out/ncu_had.log 2>&1
tail -3 gpurun_out/ncu_had.log
