#!/bin/bash
# resident CTAs (nulls) per SM vs throughput of k_trade: pads the dynamic shared memory (experiments only)
for pad in 0 13000; do      # 5 and 4 resident nulls per SM (larger pads also change the warp count: not comparable)
  v=$(SDX_TRADE_PAD_SMEM=$pad timeout 200 python bench.py --steps 3 --warmup 3 --no-extra 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%.0f nulls/s  trade_launch_ms=%.2f' % (d['value'], d['roofline']['avg_launch_ms']))
    elif 'rror' in l: print(l.strip()[:200])
")
  echo "pad=$pad : $v"
done
