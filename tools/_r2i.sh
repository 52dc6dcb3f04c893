timeout 900 python -m pytest tests/test_finalize_gpu.py tests/test_baseline_configs_gpu.py tests/test_saliency.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2i_test.log; cat gpurun_out/r2i_test.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gates > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "rc=$?"; tail -2 gpurun_out/r2i_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['analysis_config3'], d['clocks'])
P
python bench.py --mode strong --batch 4 --batch-frames 8 --distinct 2 --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-300
