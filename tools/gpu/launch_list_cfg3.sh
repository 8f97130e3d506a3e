set -x
# launch list of one full-size cfg3 step of the final tree (plain run first, then the same command under ncu)
timeout 60 python tools/profile_step.py --workload cfg3 --scale 1.0 --steps 2 > gpurun_out/plain_profile_r03b.log 2>&1 && \
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg3_r03b.csv python tools/profile_step.py --workload cfg3 --scale 1.0 --steps 2 > gpurun_out/ncu1_r03b.log 2>&1
cat gpurun_out/plain_profile_r03b.log; tail -n 2 gpurun_out/ncu1_r03b.log
