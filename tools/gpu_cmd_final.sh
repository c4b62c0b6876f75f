# final artefacts of the round: bench line, reference arm, ncu launch list of the bench command, full ncu of the dominant kernel
python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-kkt > gpurun_out/final_ncu_list.log 2>&1
tail -c 300 gpurun_out/final_ncu_list.log
