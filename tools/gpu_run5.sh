B3M_EXPERIMENT_NOLOOKBACK=1 B3M_TRACE=1 timeout 40 python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 1 2>&1 | head -8
