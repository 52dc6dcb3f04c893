for s in 0 48 32 24; do python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gates --slots $s 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('slots', d['config']['slots'], 'value %.1f e2e %.1f pageable %.1f frac %.3f clocks %s' % (d['value'], d['e2e']['value'], d['e2e']['pageable_value'], d['roofline']['frac'], d['clocks']['sm_mhz']))
"; done
