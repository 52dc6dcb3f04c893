python -m pytest tests -m gpu -x -q > gpurun_out/r2y_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_gputest.log; tail -3 gpurun_out/r2y_gputest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2y_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e']['pageable_value'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'], d['gates']['mean_epe'], d['gates']['bit_identical_pairs'], d['gates']['indices_equal'], d['cpu_baseline']['value'], d['analysis_config3']['ms_per_clip'])
print({k: round(v['ms'],4) for k,v in d['phase_roofline'].items() if isinstance(v, dict)})
P
ncu --set full --clock-control none --import-source on -k regex:tvl1_flow_kernel -s 1 -c 1 -o gpurun_out/r2y_flow_full -f python tools/profile_clip.py 64 2 > gpurun_out/r2y_ncu_full.log 2>&1; echo "full rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2y_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gates > gpurun_out/r2y_ncu_bench.log 2>&1; echo "launch list rc=$?"
bash tools/variant_ab.sh " " > gpurun_out/r2y_ab.log 2>&1; TEEFLOW_LIB=$PWD/tee_optical_flow_b200/libteeflow_tma.so python tools/ab_clip.py 2>&1 | tail -1 >> gpurun_out/r2y_ab.log; cat gpurun_out/r2y_ab.log
