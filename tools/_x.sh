bash tools/variant_ab.sh " " "-DTEEFLOW_MEDIAN_AHEAD=1" " " "-DTEEFLOW_MEDIAN_AHEAD=1" > gpurun_out/r2z_ab.log 2>&1; cat gpurun_out/r2z_ab.log
TEEFLOW_NVCC_EXTRA="-DTEEFLOW_MEDIAN_AHEAD=1" python -m tee_optical_flow_b200.build --force > /dev/null 2>&1
python tools/phase_times.py 2>&1 | grep '"ms"' | head -4
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -2
python -m tee_optical_flow_b200.build --force > /dev/null 2>&1
