bash tools/variant_probe.sh "-DTEEFLOW_MIN_CTAS=3" "-DTEEFLOW_MIN_CTAS=4" "-DTEEFLOW_PF_ROWS=0" "-DTEEFLOW_PF_ROWS=8" "-DTEEFLOW_MIN_CTAS=3 -DTEEFLOW_PF_ROWS=8" > gpurun_out/r2c_occ.log 2>&1
cat gpurun_out/r2c_occ.log
