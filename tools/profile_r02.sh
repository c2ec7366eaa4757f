#!/bin/bash
# ncu evidence of the round (run on the GPU box through gpurun, one GPU; outputs under gpurun_out/, summaries copied to profiles/ by hand):
#   1. the profiled command exits 0 without ncu
#   2. launch list of one C3 sequence, kernel-by-kernel launches: gpu__time_duration.sum + smsp__inst_executed.sum per launch
#      (sequential line searches: what the multi-sequence runs execute; then the wide ones a lone sequence gets)
#   3. one --set full capture of the solver / KLT / SSD / filter kernels of a few steady-state frames
# usage: bash tools/profile_r02.sh [tag]
tag=${1:-r02b}
out=gpurun_out
set -x
SVO_SOLVER_WIDTH=0 python tools/profile_frames.py 60 C3 > $out/${tag}_plain.log 2>&1 || exit 1
SVO_SOLVER_WIDTH=0 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file $out/${tag}_launches_seq.csv \
    python tools/profile_frames.py 60 C3 > $out/${tag}_ncu_seq.log 2>&1
SVO_SOLVER_WIDTH=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file $out/${tag}_launches_wide.csv \
    python tools/profile_frames.py 40 C3 > $out/${tag}_ncu_wide.log 2>&1
SVO_SOLVER_WIDTH=0 ncu --set full --clock-control none --import-source on -k regex:'sparse_align|klt31w|reproj_refine|stereo_ssd_mma|depth_filter' \
    --launch-skip 100 --launch-count 15 -f -o $out/${tag}_prof python tools/profile_frames.py 40 C3 > $out/${tag}_ncu_full.log 2>&1
ls -la $out/${tag}_*
