for cfg in "8 0" "3 16" "4 16" "5 16" "6 16" "8 16" "3 16"; do set -- $cfg;
  if [ "$2" = "0" ]; then unset MXQ_STAT_UNROLL; else export MXQ_STAT_UNROLL=$2; fi
  MXQ_STAT_CTAS=$1 timeout 200 python bench.py --no-e2e --no-components --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas $1 unroll $2:', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],3), round(d['roofline']['share_of_step'],3))"
done
