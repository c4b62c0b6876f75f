B200LDLT_LEVELTIMES=1 timeout 600 python tools/run_config.py c3 > gpurun_out/r2h_c3_levels.txt 2>&1; grep -E "level|factor flag" gpurun_out/r2h_c3_levels.txt | tail -45
