# statistics kernel variants under the two-stream PTQ pass: register kernel (8 CTAs / SM) vs the
# shared-memory ring kernel (MXQ_STAT_CTAS=0), pipelined and serial
for cfg in "8" "0" "0 --serial" "8 --serial"; do set -- $cfg;
  MXQ_STAT_CTAS=$1 timeout 200 python bench.py --no-e2e --no-components --steps 5 --warmup 3 $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas $1 $2:', round(d['value']), 'GB/s', round(d['ms_per_step'],2), 'ms  stats frac', round(d['roofline']['frac'],3), 'share', round(d['roofline']['share_of_step'],3))"
done
