timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2e_test.log
cat gpurun_out/r2e_test.log
bash tools/variant_ab.sh " " "-DTEEFLOW_RHO_PACK=0" "-DTEEFLOW_EARLY_PROBE=1" > gpurun_out/r2e_ab.log 2>&1
cat gpurun_out/r2e_ab.log
python tools/phase_times.py 2>&1 | grep -A1 '"inner"\|launch_ms' | head -8
