python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s3_pytest.log; python bench.py > gpurun_out/s3_bench2.json 2> gpurun_out/s3_bench.err; tail -3 gpurun_out/s3_pytest.log; python -c "
import json; d=json.load(open('gpurun_out/s3_bench2.json')); print(d['value'], d['factor_ms'], d['solve_ms'], d['roofline']['achieved'], d['roofline']['frac'], d['backward_error'])"
