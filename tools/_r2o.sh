bash tools/variant_ab.sh " " "-DTEEFLOW_TMA_ROWS=4 -DTEEFLOW_MIN_CTAS=5" "-DTEEFLOW_TMA_ROWS=4 -DTEEFLOW_MIN_CTAS=4" "-DTEEFLOW_TMA_INNER=0" > gpurun_out/r2o_ab.log 2>&1
cat gpurun_out/r2o_ab.log
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -2
