#!/bin/bash
: ${SDX_TUNE_CFGS:="4,4,1 4,8,5 4,6,4 8,8,2 8,8,5 8,6,4"}
# sweep of the k_trade launch configuration on the C1 workload (benchmarks only)
for cfg in ${SDX_TUNE_CFGS}; do
  set -- ${cfg//,/ }
  v=$(SDX_TRADE_G=$1 SDX_TRADE_WARPS=$2 SDX_TRADE_LB=$3 python bench.py --steps 3 --warmup 2 --no-extra --nulls 100000 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%.0f nulls/s  trade_launch_ms=%.2f plan_share=%.3f' % (d['value'], d['roofline']['avg_launch_ms'], d['roofline']['plan_share_of_step']))
    elif 'rror' in l: print(l.strip()[:200])
")
  echo "G=$1 WARPS=$2 LB=$3 : $v"
done
