#!/bin/bash
# is the conv kernel power-bound?  SM clock / power sampled every 100 ms during a few seconds of back-to-back launches
cd "$(dirname "$0")/.."
for cfg in "B200_EPI_DBG=0" "B200_EPI_DBG=4" "B200_EPI_DBG=3"; do
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu --format=csv,noheader -lms 100 > /tmp/smi.log &
  SMI=$!
  echo "=== $cfg"; env $cfg REPLAYS=3000 ONLY="128->128 @32" python tools/bench_conv_graph.py 2>&1 | grep "@"
  kill $SMI; wait $SMI 2>/dev/null
  sort /tmp/smi.log | uniq -c | sort -k1,1nr | head -6
done
