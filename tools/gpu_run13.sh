set -x
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py 2> gpurun_out/bench_cfg3_final_r01.err | grep "^{" > gpurun_out/bench_cfg3_final_r01.json
tail -c 300 gpurun_out/bench_cfg3_final_r01.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep "^{" > gpurun_out/bench_ref_final_r01.json
