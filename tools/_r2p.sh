python -m pytest tests -m gpu -x -q > gpurun_out/r2p_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_gputest.log; tail -3 gpurun_out/r2p_gputest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2p_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'], d['roofline'], d['clocks'], d['gates']['mean_epe'], d['gates']['indices_equal'], d['cpu_baseline']['value'])
P
