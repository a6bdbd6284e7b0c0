for st in 2 3 4; do echo "=== STAGES=$st"; B200_STAGES=$st python tools/bench_conv.py 2>&1 | sed -n '1p;7p;9p'; done
echo "=== MAX_NP=128 (6 stages)"; B200_MAX_NP=128 python tools/bench_conv.py 2>&1 | sed -n '1p;7p;9p'
echo "=== MAX_NP=128 stages 3"; B200_MAX_NP=128 B200_STAGES=3 python tools/bench_conv.py 2>&1 | sed -n '1p;7p;9p'
