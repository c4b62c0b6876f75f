for N in 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2p_bench$N.json 2> gpurun_out/r2p_bench$N.err; tail -c 1300 gpurun_out/r2p_bench$N.json; echo
done
