set -x
timeout 300 python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 2 --numblocks 4 > gpurun_out/plain_profile_nb4_r01s.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cfg3q_nb4_r01s.csv python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 1 --numblocks 4 > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_gap|k_walk|merge_run' -c 8 -o gpurun_out/prof_cfg3q_nb4_r01s -f python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 1 --numblocks 4 > gpurun_out/ncu2.log 2>&1
cat gpurun_out/plain_profile_nb4_r01s.log; tail -n 2 gpurun_out/ncu1.log gpurun_out/ncu2.log; ls -la gpurun_out
timeout 300 python bench.py --workload cfg3 --scale 0.25 --numblocks 4 --steps 3 --warmup 3 --no-cpu 2>/dev/null | grep "^{" > gpurun_out/bench_cfg3q_nb4_r01s.json
