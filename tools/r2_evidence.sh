#!/bin/bash
# Round-2 evidence run on one B200 (gpurun): GPU tests, the bench lines, per-kernel and per-shape
# timings and the ncu launch list of the bench command.  Everything lands in gpurun_out/.
#   gpurun --timeout 1700 -- 'bash tools/r2_evidence.sh'
set -u
O=gpurun_out
mkdir -p $O
export AEQB_BENCH_TMP=${AEQB_BENCH_TMP:-/dev/shm}
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $O/r2_gputest.log 2>&1
tail -3 $O/r2_gputest.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > $O/r2_bench.json 2> $O/r2_bench.err
tail -c 600 $O/r2_bench.err
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
timeout 300 python tools/ktime.py > $O/r2_ktime.txt 2>&1
timeout 400 python tools/shape_bench.py > $O/r2_shape_bench.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv \
  --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --modes headline --no-cpu-baseline \
  > $O/r2_launches_bench.log 2>&1
python tools/ncu_summary.py launches $O/r2_launches.csv > $O/r2_launches.txt 2>&1
ls -la $O
