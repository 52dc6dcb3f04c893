ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gates > gpurun_out/r2q_ncu_bench.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tvl1_flow_kernel -s 1 -c 1 -o gpurun_out/r2q_flow_full -f python tools/profile_clip.py 64 2 > gpurun_out/r2q_ncu_full.log 2>&1; echo "full rc=$?"
TEEFLOW_LIB=$PWD/tee_optical_flow_b200/libteeflow_tma.so ncu --set full --clock-control none --import-source on -k regex:tvl1_flow_kernel -s 1 -c 1 -o gpurun_out/r2q_flow_tma_full -f python tools/profile_clip.py 64 2 > gpurun_out/r2q_ncu_tma_full.log 2>&1; echo "tma full rc=$?"
ls -la gpurun_out/*.ncu-rep
