timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2g_pytest.log; cat gpurun_out/r2g_pytest.log
timeout 600 python tools/run_config.py c3 > gpurun_out/r2g_c3.txt 2>&1; cat gpurun_out/r2g_c3.txt
timeout 600 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 1500 gpurun_out/r2g_bench.json; tail -3 gpurun_out/r2g_bench.err
