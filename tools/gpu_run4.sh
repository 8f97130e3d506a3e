set -x
timeout 300 python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 2 > gpurun_out/plain_profile_r01i.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_radix_onesweep' -s 4 -c 2 -o gpurun_out/prof_cfg3q_r01i -f python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 2 > gpurun_out/ncu2.log 2>&1
cat gpurun_out/plain_profile_r01i.log; tail -n 3 gpurun_out/ncu2.log
