#!/bin/bash
# Round-2 evidence run on one B200 (gpurun): smoke, GPU tests, the bench lines, per-kernel and per-shape
# timings, the ncu launch list of the bench command and full captures of the kernels the round worked
# on.  Everything lands in gpurun_out/; the summaries are copied to profiles/ by hand.
#   gpurun --timeout 1900 -- 'bash tools/r2_evidence.sh'
set -u
O=gpurun_out
mkdir -p $O
export AEQB_BENCH_TMP=${AEQB_BENCH_TMP:-/dev/shm}
NCU="ncu --clock-control none"
( time timeout 300 python __graft_entry__.py smoke ) > $O/r2_smoke.log 2>&1
tail -2 $O/r2_smoke.log
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/r2_gputest.log 2>&1
tail -3 $O/r2_gputest.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > $O/r2_bench.json 2> $O/r2_bench.err
tail -c 300 $O/r2_bench.err
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
timeout 600 python bench.py --steps 10 --warmup 3 --modes all --no-cpu-baseline > $O/r2_bench_all.json 2> $O/r2_bench_all.err
timeout 300 python tools/ktime.py > $O/r2_ktime.txt 2>&1
timeout 200 python tools/hinv_time.py > $O/r2_hinv_time.txt 2>&1
timeout 400 python tools/shape_bench.py > $O/r2_shape_bench.txt 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 9000 --csv \
  --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --modes headline --no-cpu-baseline \
  > $O/r2_launches_bench.log 2>&1
python tools/ncu_summary.py launches $O/r2_launches.csv > $O/r2_launches.txt 2>&1
HINV_VARIANTS="dmma+lookahead" timeout 300 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file $O/r2_hinv_launches.csv \
  python tools/hinv_time.py > $O/r2_hinv_under_ncu.log 2>&1
python tools/ncu_summary.py launches $O/r2_hinv_launches.csv > $O/r2_hinv_launches.txt 2>&1
# full captures: the OBS loop's two kernels (middle of the layer), the Cholesky kernels at K = 11008
timeout 300 $NCU --set full --import-source on -k regex:gptq_ -s 60 -c 6 -f -o $O/r2_gptq_obs \
  python tools/ktime.py --only "gptq_quantize" --iters 1 > $O/r2_ncu_gptq.log 2>&1
HINV_VARIANTS="dmma serial" timeout 300 $NCU --set full --import-source on -k regex:"chol_" -s 12 -c 4 -f -o $O/r2_chol_dmma \
  python tools/hinv_time.py 11008 > $O/r2_ncu_chol.log 2>&1
for f in r2_gptq_obs r2_chol_dmma; do
  python tools/ncu_summary.py full $O/$f.ncu-rep > $O/$f.txt 2>&1
done
ls -la $O | tail -30
