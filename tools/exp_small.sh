for f in 0 1 2 3; do echo "=== MIN_FILL=$f/4 SMs"; B200_MIN_FILL=$f python tools/bench_conv.py 2>&1 | sed -n '10,13p'; done
