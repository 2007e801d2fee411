#!/bin/bash
# End-of-round-2 evidence on one B200 after the last kernel changes (96 KiB row tiles, folded row maxima,
# per-output pass 2, 64 KiB block stages): smoke, GPU tests, bench lines, per-kernel / per-shape timings,
# the ncu launch list of the bench command (our kernels) and full captures of the two headline kernels.
#   gpurun --timeout 900 -- 'bash tools/r2_final.sh'
set -u
O=gpurun_out
mkdir -p $O
export AEQB_BENCH_TMP=${AEQB_BENCH_TMP:-/dev/shm}
NCU="ncu --clock-control none"
( time timeout 120 python __graft_entry__.py smoke ) > $O/r2_smoke.log 2>&1
tail -2 $O/r2_smoke.log
( time timeout 200 python -m pytest tests -m gpu -q ) > $O/r2_gputest.log 2>&1
tail -3 $O/r2_gputest.log
( time timeout 400 python bench.py ) > $O/r2_bench.json 2> $O/r2_bench.err
tail -c 300 $O/r2_bench.err
( time timeout 200 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
timeout 120 python tools/ktime.py > $O/r2_ktime.txt 2>&1
timeout 100 python tools/shape_bench.py > $O/r2_shape_bench.txt 2>&1
timeout 100 python tools/sustain_ab.py > $O/r2_sustain_rows.txt 2>&1
SUSTAIN_KIND=blocks timeout 100 python tools/sustain_ab.py > $O/r2_sustain_blocks.txt 2>&1
timeout 240 $NCU --metrics gpu__time_duration.sum -k regex:"requant_|minmax_|mirror_|ema_|pack_kernel|scale_|quantize_kernel|row_stats|octav_|hadamard_" -c 4000 --csv \
  --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --modes headline --no-cpu-baseline \
  > $O/r2_launches_bench.log 2>&1
python tools/ncu_summary.py launches $O/r2_launches.csv > $O/r2_launches.txt 2>&1
timeout 200 $NCU --set full --import-source on -k regex:requant_rows_stream -s 2 -c 1 -f -o $O/r2_requant_rows_stream_477 \
  python bench.py --steps 1 --warmup 3 --modes headline --no-cpu-baseline > $O/r2_ncu_rows.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:requant_blocks_stream -s 2 -c 1 -f -o $O/r2_requant_blocks_stream_477 \
  python bench.py --steps 1 --warmup 3 --modes headline --no-cpu-baseline > $O/r2_ncu_blocks.log 2>&1
for f in r2_requant_rows_stream_477 r2_requant_blocks_stream_477; do
  python tools/ncu_summary.py full $O/$f.ncu-rep > $O/$f.txt 2>&1
done
ls -la $O | tail -20
