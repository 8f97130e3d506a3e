set -x
timeout 300 python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 2 --numblocks 4 > gpurun_out/plain_profile_nb4_r02h.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:'k_scan_tile|k_copy_big|k_dict2_pack' -c 16 -o gpurun_out/prof_cfg3q_nb4_merge_r02h -f python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 1 --numblocks 4 > gpurun_out/ncu2.log 2>&1
cat gpurun_out/plain_profile_nb4_r02h.log; tail -n 2 gpurun_out/ncu2.log; ls -la gpurun_out/prof_cfg3q_nb4_merge_r02h.ncu-rep
