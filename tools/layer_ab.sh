#!/bin/bash
# A/B of conv-kernel options on ONE box (boxes differ by 5-10 % in the clocks they hold): per-forward ms of RetinaFace-R50,
# 64 x 1080p, for each option set given as "NAME:ENV=VAL,ENV=VAL".
for spec in "$@"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  ( IFS=','; for kv in $envs; do export "$kv"; done; python tools/layer_times.py retina_r50 64 10 > "gpurun_out/r02_layer_times_r50_${name}.txt" 2>&1; echo "$name $(head -1 gpurun_out/r02_layer_times_r50_${name}.txt)" )
done
