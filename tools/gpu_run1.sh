set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg2_r01b.json 2> gpurun_out/bench_cfg2_r01b.err
tail -c 600 gpurun_out/bench_cfg2_r01b.err
python bench.py --steps 5 --warmup 3 --numblocks 4 --no-cpu > gpurun_out/bench_cfg2_nb4_r01b.json 2>&1
python tools/profile_step.py --steps 2 > gpurun_out/plain_profile.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_cfg2_r01b.csv python tools/profile_step.py --steps 2 > gpurun_out/ncu1.log 2>&1
python tools/profile_step.py --steps 2 > gpurun_out/plain_profile.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_radix_onesweep|k_scan_tile|k_make_keys|k_walk|k_extract_bwt' -s 20 -c 14 -o gpurun_out/prof_cfg2_r01b -f python tools/profile_step.py --steps 2 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
python tools/profile_step.py --steps 2 --numblocks 4
python tools/profile_step.py --workload cfg3 --scale 0.33 --steps 2
python tools/profile_step.py --workload cfg3 --scale 0.33 --steps 2 --numblocks 8
