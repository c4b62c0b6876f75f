# usage: bash tools/gpu_cmd_n.sh N   -- bench.py on N GPUs under torchrun
N=$1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/s3_bench_n$N.json 2> gpurun_out/s3_bench_n$N.err
tail -3 gpurun_out/s3_bench_n$N.err | cut -c1-400
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/s3_bench_n$N.json") if l.startswith("{")][-1])
print({k:d[k] for k in ('value','n_gpus','factor_ms','solve_ms','stage_ms_rank0','backward_error','num_neg')}, d['config']['parallelism'])
PY
