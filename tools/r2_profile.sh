#!/bin/bash
# Round-2 profiles on one B200 (gpurun): per-kernel launch lists of the secondary kernels and
# `ncu --set full` captures of the kernels the round worked on.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/r2_profile.sh'
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
# (1) launch lists: every kernel of one call of each C-ABI entry point / of the Hessian inverse
timeout 300 $NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file $O/r2_ktime_launches.csv \
  python tools/ktime.py --iters 2 > $O/r2_ktime_under_ncu.log 2>&1
python tools/ncu_summary.py launches $O/r2_ktime_launches.csv > $O/r2_ktime_launches.txt 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file $O/r2_hinv_launches.csv \
  python tools/hinv_time.py > $O/r2_hinv_under_ncu.log 2>&1
python tools/ncu_summary.py launches $O/r2_hinv_launches.csv > $O/r2_hinv_launches.txt 2>&1
# (2) full captures
timeout 400 $NCU --set full --import-source on -k regex:requant_rows_stream -s 2 -c 1 -f -o $O/r2_requant_rows_stream_477 \
  python bench.py --steps 1 --warmup 3 --modes headline --no-cpu-baseline > $O/r2_ncu_rows.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:gptq_ -s 60 -c 8 -f -o $O/r2_gptq_obs \
  python tools/ktime.py --only "gptq_quantize" --iters 1 > $O/r2_ncu_gptq.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:octav_rows_warp -s 2 -c 1 -f -o $O/r2_octav_rows_warp \
  python tools/ktime.py --only "octav_clip_rows b4" --iters 2 > $O/r2_ncu_octav.log 2>&1
for f in r2_requant_rows_stream_477 r2_gptq_obs r2_octav_rows_warp; do
  python tools/ncu_summary.py full $O/$f.ncu-rep > $O/$f.txt 2>&1
done
python tools/hinv_time.py > $O/r2_hinv_time.txt 2>&1
ls -la $O
