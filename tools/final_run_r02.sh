#!/bin/bash
# Round-2 evidence run on the GPU box (under gpurun, 1 GPU, from the repo root):
#   full GPU suite, smoke, the bench line + reference arm, launch lists, ncu captures of the hot kernels.
set -u
mkdir -p gpurun_out
T=r02
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${T}_gputests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err; tail -2 gpurun_out/bench_${T}.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${T}_ref.json 2> gpurun_out/bench_${T}_ref.err
timeout 600 python bench.py --workload vet_sl12_2048 --no-extras --no-cpu --no-parity --steps 3 --warmup 3 > gpurun_out/bench_${T}_vet.json 2> gpurun_out/bench_${T}_vet.err
timeout 300 python tools/sl_timing.py 2>&1 | tail -1 | tee gpurun_out/${T}_sl_timing.json
timeout 300 python tools/lk_timing.py 2>&1 | tail -5 | tee gpurun_out/${T}_lk_timing.log
timeout 300 python tools/sl_f32_timing.py 2>/dev/null | tail -1 | tee gpurun_out/${T}_sl_f32_timing.json
timeout 300 python tools/vet_time.py 2>&1 | tail -2 | tee gpurun_out/${T}_vet_time.log
REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_lk.csv python tools/lk_once.py > /dev/null 2>&1
REPS=3 FIELD=lk timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_sl.csv python tools/sl_once.py > /dev/null 2>&1
REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_sl_f32.csv python tools/sl_f32_once.py > /dev/null 2>&1
REPS=2 FIELD=lk timeout 600 ncu --set full --clock-control none --import-source on -k regex:sl_multistep -s 1 -c 1 -f -o gpurun_out/${T}_sl python tools/sl_once.py > /dev/null 2>&1
REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sl_f32_kernel -s 1 -c 1 -f -o gpurun_out/${T}_sl_f32 python tools/sl_f32_once.py > /dev/null 2>&1
REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"idw32_kernel|front_kernel|lk_track_kernel|box_chain|outliers_warp|idw_fix_warp|kd_build|select_smem" -s 20 -c 12 -f -o gpurun_out/${T}_lk python tools/lk_once.py > /dev/null 2>&1
MAXITER=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vet_eval -s 40 -c 2 -f -o gpurun_out/${T}_vet python tools/vet_once.py > /dev/null 2>&1
ls -la gpurun_out | grep ${T} | awk '{print $5, $9}'
