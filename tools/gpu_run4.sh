set -x
timeout 300 python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 2 > gpurun_out/plain_profile_r01m.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_resolve' -s 1 -c 1 -o gpurun_out/prof_cfg3q_resolve_r01m -f python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 2 > gpurun_out/ncu2.log 2>&1
cat gpurun_out/plain_profile_r01m.log; tail -n 3 gpurun_out/ncu2.log
