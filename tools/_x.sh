bash tools/variant_ab.sh " " "-DTEEFLOW_POINT_ROWS=32" "-DTEEFLOW_WARP_PF=5" "-DTEEFLOW_POINT_ROWS=24" " " > gpurun_out/r2x_ab.log 2>&1; cat gpurun_out/r2x_ab.log
python tools/phase_times.py 2>&1 | grep '"ms"' | head -4
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -2
