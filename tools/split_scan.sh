for s in 0.64 0.68 0.71 0.73 0.75 0.78 0.82 0.88; do UM_TPRF_SPLIT_SPEED=$s python tools/sweep_time.py 0 262144 8192 20 2>&1 | tail -1 | sed "s/^/split $s: /"; done
