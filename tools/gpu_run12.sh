B3M_EXPERIMENT_SORTED_PASS=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | grep "^{" > gpurun_out/bench_exp_sorted.json
