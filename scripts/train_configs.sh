#!/bin/bash
# BASELINE.json configs 3-5 as full training steps (bench.py --mode train); usage: train_configs.sh <ngpus> [impl]
N=${1:-1}; IMPL=${2:-own}
run() {  # workload, layers, tag
  if [ "$N" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200))"; fi
  $L bench.py --gpus $N --mode train --impl $IMPL --workload $1 --num-layers $2 --steps 20 --warmup 5 > gpurun_out/train_${IMPL}_$3_n$N.json 2> gpurun_out/train_${IMPL}_$3_n$N.err
  python -c "
import json; d=json.load(open('gpurun_out/train_${IMPL}_$3_n$N.json')); print('$3', 'n=$N', '$IMPL', round(d['ms_per_step'],2), 'ms/step', round(d['value'],2), 'steps/s', round(d['frames_per_s']), 'frames/s')" || tail -3 gpurun_out/train_${IMPL}_$3_n$N.err
}
run mono+stereo_640x192_b12 18 c3_monostereo_r18
run mono_1024x320_b12 18 c4_hires_r18
run mono_640x192_b12_avg_reprojection 50 c5_avg_r50
run mono_640x192_b12_disable_automasking 50 c5_noauto_r50
