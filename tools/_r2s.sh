for s in 0 96 128 192; do python bench.py --mode strong --batch 8 --batch-frames 32 --distinct 8 --steps 3 --warmup 1 --slots $s 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('slots $s', 'value %.1f ms %.1f' % (d['value'], d['ms_per_step']))
"; done
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "batch" 2>&1 | tail -2
