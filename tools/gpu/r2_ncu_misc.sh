# --set full captures of the round-2 kernels outside the cfg3 headline path: K5 (gap chains + list count), doubling rounds
TAG=${1:-r2z}
set -x
timeout 400 ncu --set full --clock-control none -k regex:'^k_gap' -s 2 -c 4 -o gpurun_out/${TAG}_prof_gap -f python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 1 --numblocks 4 > gpurun_out/${TAG}_ncu_gap.log 2>&1
tail -n 2 gpurun_out/${TAG}_ncu_gap.log
timeout 400 ncu --set full --clock-control none -k regex:'k_dbl_tile|k_dbl_compact|k_rank_write' -s 3 -c 5 -o gpurun_out/${TAG}_prof_dbl -f python tools/profile_step.py --workload cfg4 --scale 0.1 --steps 1 > gpurun_out/${TAG}_ncu_dbl.log 2>&1
tail -n 2 gpurun_out/${TAG}_ncu_dbl.log
