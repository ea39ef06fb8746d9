#!/bin/bash
# One gpurun call: plain run of the bench command, then the ncu launch list and the full captures of it.
# usage (on the GPU box): bash scripts/profile_round.sh <tag>
tag=${1:-r01b}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-graph"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:md2_march -s 4 -c 1 -f -o gpurun_out/prof_march_$tag $CMD > gpurun_out/ncu_march_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:md2_final -s 4 -c 1 -f -o gpurun_out/prof_final_$tag $CMD > gpurun_out/ncu_final_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:md2_identity -s 4 -c 1 -f -o gpurun_out/prof_identity_$tag $CMD > gpurun_out/ncu_identity_$tag.log 2>&1
ls -la gpurun_out/*$tag*
