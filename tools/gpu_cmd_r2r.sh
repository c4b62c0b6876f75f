timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2r_pytest.log; cat gpurun_out/r2r_pytest.log
for cfg in "X=1" "B200LDLT_DEFER_MINM=1024" "B200LDLT_DEFER_MINM=1024 B200LDLT_DEFER=2" "B200LDLT_DEFER=2" "B200LDLT_DEFER=1"; do echo "== $cfg"; env $cfg timeout 300 python tools/profile_run.py 64 3 2>&1 | grep -E "factor ms" | tail -2; done
B200LDLT_DEFER_MINM=1024 B200LDLT_LEVELTIMES=1 timeout 300 python tools/profile_run.py 64 2 2>&1 | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2r_bench2.json 2> gpurun_out/r2r_bench2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2r_bench2.json').read().strip().splitlines()[-1]); print(d['value'], d['factor_ms'], d['solve_ms'], d['stage_ms_rank0'])"; tail -2 gpurun_out/r2r_bench2.err
