timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2l_test.log; cat gpurun_out/r2l_test.log
timeout 300 python tools/ab_clip.py 2>&1 | tail -1
