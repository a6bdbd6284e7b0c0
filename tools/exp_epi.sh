#!/bin/bash
# experiment: epilogue cost (B200_EPI_DBG bits: 1 no residual loads, 2 no stores, 4 nothing) x CTA pairs on/off
for pair in 1 0; do for dbg in ${DBGS:-0 4}; do echo "PAIR=$pair DBG=$dbg"; B200_PAIR=$pair B200_EPI_DBG=$dbg python tools/bench_conv_graph.py 2>&1 | grep -E "@"; done; done
