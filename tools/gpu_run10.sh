for g in 32 64 128; do
B3M_L2_FETCH=$g timeout 120 python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 3 --numblocks 4 2>&1 | tail -2
B3M_L2_FETCH=$g timeout 120 python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 3 --numblocks 1 2>&1 | tail -1
done
timeout 120 python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 3 --numblocks 4 2>&1 | tail -1
