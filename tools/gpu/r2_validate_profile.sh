# round 2, call A: the whole GPU suite (with the full-size oracle checkbwt tests), smoke, the default bench line,
# the ncu launch list and --set full captures of the MSD kernels, bench lines of the other configs
TAG=${1:-r2a}
set -x
timeout 1500 python -m pytest tests -m gpu -q --tb=short --durations=12 > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -25 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-400
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-sample 64000000 --cpu-curve "" 2> gpurun_out/${TAG}_bench_cfg3_n1.err | grep "^{" > gpurun_out/${TAG}_bench_cfg3_n1.json
tail -c 400 gpurun_out/${TAG}_bench_cfg3_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg3_n1.json"))
print("ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "roof", d["roofline"])
print(d["phases_ms"]); print(d["kernels_ms_per_step"]); print(d["clocks"])
PY
timeout 300 python tools/profile_step.py --workload cfg3 --scale 1.0 --steps 2 > gpurun_out/${TAG}_plain_profile.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_cfg3_launches.csv python tools/profile_step.py --workload cfg3 --scale 1.0 --steps 2 > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_msd_scatter|k_msd_local|k_msd_finish|k_msd_count|k_msd_col' -s 7 -c 7 -o gpurun_out/${TAG}_prof_cfg3 -f python tools/profile_step.py --workload cfg3 --scale 1.0 --steps 2 > gpurun_out/${TAG}_ncu2.log 2>&1
cat gpurun_out/${TAG}_plain_profile.log; tail -n 2 gpurun_out/${TAG}_ncu1.log gpurun_out/${TAG}_ncu2.log
for W in cfg2 cfg5 cfg4 cfg1; do
  timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu 2> gpurun_out/${TAG}_bench_${W}_n1.err | grep "^{" > gpurun_out/${TAG}_bench_${W}_n1.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_${W}_n1.json"))
    print("$W ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], d["phases_ms"], d["kernels_ms_per_step"], d["counters"])
except Exception as e:
    print("$W failed", e)
PY
done
