timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2f_test.log
cat gpurun_out/r2f_test.log
bash tools/variant_ab.sh " " "-DTEEFLOW_INNER_PAIR=0" "-DTEEFLOW_MIN_CTAS=7" "-DTEEFLOW_MIN_CTAS=8" > gpurun_out/r2f_ab.log 2>&1
cat gpurun_out/r2f_ab.log
python tools/phase_times.py 2>&1 | grep -A1 '"inner"' | head -3
