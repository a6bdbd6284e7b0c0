#!/bin/bash
# round 2: what bounds the 4x4 / 8x8 convolutions?  single layers back to back in a CUDA graph, B = 256, tile-size /
# K-rotation / pair / ring-depth knobs (tools/bench_conv_graph.py)
cd "$(dirname "$0")/.."
for sel in '@4' '@8'; do
  for env in "" "B200_MAX_NP=256" "B200_MAX_NP=128" "B200_MAX_NP=64" "B200_K_ROTATE=0" "B200_PAIR=0" "B200_STAGES=3" "B200_VTAP=0"; do
    echo "--- ONLY=$sel $env"
    env ONLY=$sel $env python tools/bench_conv_graph.py 2>&1 | grep -v "^env"
  done
done
