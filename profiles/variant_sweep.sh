#!/bin/bash
# Tuning aid: time K1 for every library variant under build/variants/ (built with different -DMSP_CULL_* macros).
for so in build/variants/*.so; do
  MSP_LIB=$PWD/$so python bench.py --scenes 12 --steps 5 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$so', 'pairs/s=%.0f' % d['value'], 'k1_ms=%.4f' % d['roofline']['avg_launch_ms'])"
done
