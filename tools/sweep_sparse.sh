for cfg in "0 0" "1 0" "1 32" "1 40" "1 48" "1 64"; do
  set -- $cfg
  v=$(SDX_TRADE_SPARSE=$1 SDX_PLAN_CAP=$2 timeout 120 python bench.py --steps 3 --warmup 3 --no-extra 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%.0f nulls/s  trade_launch_ms=%.2f plan_share=%.3f exceed=%d' % (d['value'], d['roofline']['avg_launch_ms'], d['roofline']['plan_share_of_step'], d['exceedances']))
    elif 'rror' in l: print(l.strip()[:200])
")
  echo "SPARSE=$1 CAP=$2 : $v"
done
