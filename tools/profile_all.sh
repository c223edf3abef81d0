#!/bin/bash
# Evidence for one build, run on the GPU box:   gpurun --timeout 1500 -- 'bash tools/profile_all.sh r2a'
# Writes into gpurun_out/ (copied under profiles/ once read):
#   <tag>_bench.json          the bench line of this build (NOT under a profiler) + per-layer event times (<tag>_bench_detail.txt)
#   <tag>_launches.csv        every launch of ONE steady-state frame with its device time (cold-cache, serialised: compare shares)
#   <tag>_frame_raw.csv       ncu --set full of every kernel of one steady-state frame, raw page (tensor pipe %, DRAM bytes, issue %, registers)
#   <tag>_prog_raw.csv / <tag>_prog_source.csv   ncu --set full + source of one conv_prog_kernel launch (per-instruction stall reasons)
# The .ncu-rep files stay in /tmp on the box: gpurun_out/ is capped at 64 MiB.
tag=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
BENCH_DETAIL=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${tag}_bench_detail.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > gpurun_out/${tag}_launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -f -o /tmp/${tag}_frame \
    python tools/profile_step.py > gpurun_out/${tag}_frame.log 2>&1
ncu -i /tmp/${tag}_frame.ncu-rep --page raw --csv > gpurun_out/${tag}_frame_raw.csv 2>/dev/null
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_prog_kernel -s 3 -c 1 -f \
    -o /tmp/${tag}_prog python tools/profile_step.py > gpurun_out/${tag}_prog.log 2>&1
ncu -i /tmp/${tag}_prog.ncu-rep --page raw --csv > gpurun_out/${tag}_prog_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_prog.ncu-rep --page source --csv > gpurun_out/${tag}_prog_source.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out/
