#!/bin/bash
# diagnostic matrix of the conv kernel: epilogue knob x pipeline depth
cd "$(dirname "$0")/.."
for cfg in "B200_EPI_DBG=0" "B200_EPI_DBG=1" "B200_EPI_DBG=2" "B200_EPI_DBG=3" "B200_EPI_DBG=4" "B200_STAGES=3" "B200_STAGES=2" "B200_STAGES=3 B200_EPI_DBG=4" "B200_EPI_HALVES=1" "B200_PAIR=0" "B200_PAIR=0 B200_EPI_DBG=4"; do
  echo "=== $cfg"; env $cfg python tools/bench_conv_graph.py 2>&1 | grep -E "@16 \+res|@32|proj|@8 \+res"
done
