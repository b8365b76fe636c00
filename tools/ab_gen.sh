#!/bin/bash
# A/B of library builds on one box: tools/ab_gen.sh lib1.so lib2.so ...   (prints value / ms per step of the default workload)
for lib in "$@"; do
  for rep in 1 2; do
    BT_LIB_PATH=$PWD/$lib python bench.py --steps 200 --warmup 10 --no-cpu-baseline --preprocess-source 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['ms_per_step'],4), round(d['roofline']['avg_launch_ms'],4))"
  done
done
