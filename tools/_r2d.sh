bash tools/variant_ab.sh " " "-DTEEFLOW_EARLY_TICKET=1" "-DTEEFLOW_EARLY_PROBE=1" "-DTEEFLOW_EARLY_TICKET=1 -DTEEFLOW_EARLY_PROBE=1" "-DTEEFLOW_MIN_CTAS=4" "-DTEEFLOW_MIN_CTAS=5" " " > gpurun_out/r2d_ab.log 2>&1
cat gpurun_out/r2d_ab.log
