bash tools/variant_ab.sh " " "-DTEEFLOW_WARP_PF1=1" "-DTEEFLOW_WARP_PF1=2" "-DTEEFLOW_MEDIAN_PF1=1" "-DTEEFLOW_WARP_PF1=1 -DTEEFLOW_MEDIAN_PF1=1" "-DTEEFLOW_WARP_PF=2" " " > gpurun_out/r2k_ab.log 2>&1
cat gpurun_out/r2k_ab.log
TEEFLOW_NVCC_EXTRA="-DTEEFLOW_WARP_PF1=1 -DTEEFLOW_MEDIAN_PF1=1" python -m tee_optical_flow_b200.build --force > /dev/null 2>&1
python tools/phase_times.py 2>&1 | grep -B1 -A1 '"ms"' | grep '"ms"\|{' | head -12
