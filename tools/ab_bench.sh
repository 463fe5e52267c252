# Developer A/B of variant libraries on the bench's device-resident step: VARIANTS="base sp" bash tools/ab_bench.sh
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-base}; do
  if [ "$v" = base ]; then lib=jacks_b200/libjacks_b200.so; else lib=jacks_b200/libjacks_b200_$v.so; fi
  JACKS_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --no-aux --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'ms_per_step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'job', round(d['job']['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2))"
done
