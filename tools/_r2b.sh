set -x
python tools/profile_clip.py 64 3 > gpurun_out/r2b_clip.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gates > gpurun_out/r2b_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tvl1_flow_kernel -s 1 -c 1 -o gpurun_out/r2b_flow_full -f python tools/profile_clip.py 64 2 > gpurun_out/r2b_ncu_full.log 2>&1
bash tools/variant_probe.sh " @TEEFLOW_SPEC=1.5" " @TEEFLOW_SPEC=1.1" " @TEEFLOW_SPEC=3" > gpurun_out/r2b_spec.log 2>&1
TEEFLOW_NVCC_EXTRA="-DTEEFLOW_FLOW_STATS=1" python -m tee_optical_flow_b200.build --force > /dev/null 2>&1
python tools/profile_clip.py 64 3 > gpurun_out/r2b_flowstats.log 2>&1
TEEFLOW_SPEC=1.5 python tools/profile_clip.py 64 3 > gpurun_out/r2b_flowstats_spec.log 2>&1
tail -3 gpurun_out/r2b_clip.log; cat gpurun_out/r2b_spec.log; tail -12 gpurun_out/r2b_flowstats.log; tail -12 gpurun_out/r2b_flowstats_spec.log
