set -x
timeout 300 python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 2 > gpurun_out/plain_profile_r01g.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_cfg3q_r01g.csv python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 2 > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_radix_onesweep|k_resolve|k_hist_4mers' -s 6 -c 6 -o gpurun_out/prof_cfg3q_r01g -f python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 2 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/plain_profile_r01g.log gpurun_out/ncu1.log gpurun_out/ncu2.log
