mkdir -p gpurun_out
python tools/cublas_shapes.py gpurun_out/r2_cublas_shapes.json > gpurun_out/r2_cublas_shapes.txt 2>&1
B200LDLT_LEVELTIMES=1 python tools/profile_run.py 64 3 > gpurun_out/r2a_levels.txt 2>&1
python tools/run_config.py c4 > gpurun_out/r2a_c4.txt 2>&1
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -30 gpurun_out/r2a_levels.txt; cat gpurun_out/r2_cublas_shapes.txt; cat gpurun_out/r2a_c4.txt
