# final validation of the round-2 tree on one GPU: the whole GPU suite, smoke, the bench line (short CPU leg), the ncu
# launch list of one full-size cfg3 step, and -- if asked -- --set full captures of the block-path and doubling kernels
TAG=${1:-r2z}; FULL=${2:-0}
set -x
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x --durations=6 > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -14 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-400
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-sample 64000000 --cpu-curve "" 2> gpurun_out/${TAG}_bench_cfg3_n1.err | grep "^{" > gpurun_out/${TAG}_bench_cfg3_n1.json
tail -c 300 gpurun_out/${TAG}_bench_cfg3_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg3_n1.json"))
print("ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "roof", d["roofline"])
print(d["kernels_ms_per_step"]); print(d["clocks"]); print(d["cpu_baseline"]); print(d.get("file_level"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_cfg3_launches.csv python tools/profile_step.py --workload cfg3 --scale 1.0 --steps 2 > gpurun_out/${TAG}_ncu1.log 2>&1
tail -n 2 gpurun_out/${TAG}_ncu1.log
if [ "$FULL" = "1" ]; then
timeout 600 ncu --set full --clock-control none -k regex:'k_gap$|k_gap_hist|k_rank_write|k_dbl_tile|k_dbl_compact' -c 6 -o gpurun_out/${TAG}_prof_misc -f python tools/profile_step.py --workload cfg4 --scale 0.1 --steps 1 > gpurun_out/${TAG}_ncu2.log 2>&1
tail -n 2 gpurun_out/${TAG}_ncu2.log
fi
