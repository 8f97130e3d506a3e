B3M_EXPERIMENT_MATCH=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | grep "^{" > gpurun_out/bench_exp_match.json
