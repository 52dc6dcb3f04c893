bash tools/variant_ab.sh " " "-DTEEFLOW_TMA_STAGES=3 -DTEEFLOW_MIN_CTAS=5" "-DTEEFLOW_TMA_ROWS=4 -DTEEFLOW_MIN_CTAS=4" "-DTEEFLOW_TMA_STAGES=3" "-DTEEFLOW_TMA_INNER=0" > gpurun_out/r2n_ab.log 2>&1
cat gpurun_out/r2n_ab.log
python tools/phase_times.py 2>&1 | grep -A1 '"inner"' | head -3
