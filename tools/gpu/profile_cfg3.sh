set -x
timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_cfg3_r01q.err | grep "^{" > gpurun_out/bench_cfg3_r01q.json
tail -c 300 gpurun_out/bench_cfg3_r01q.err
timeout 300 python tools/profile_step.py --workload cfg3 --scale 1.0 --steps 2 > gpurun_out/plain_profile_r01q.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg3_r01q.csv python tools/profile_step.py --workload cfg3 --scale 1.0 --steps 2 > gpurun_out/ncu1.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_radix_onesweep|k_resolve|k_hist_4mers|k_unpack_pac|k_pack2' -s 8 -c 8 -o gpurun_out/prof_cfg3_r01q -f python tools/profile_step.py --workload cfg3 --scale 1.0 --steps 2 > gpurun_out/ncu2.log 2>&1
cat gpurun_out/plain_profile_r01q.log; tail -n 2 gpurun_out/ncu1.log gpurun_out/ncu2.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/bench_ref_r01q.err | grep "^{" > gpurun_out/bench_ref_r01q.json
