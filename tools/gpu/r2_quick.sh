# round 2, quick loop: the sorter's tests, a short cfg3 bench line (no CPU leg), launch list + --set full of the named kernels
TAG=${1:-r2b}; KERN=${2:-k_msd_finish}; SKIP=${3:-0}; CNT=${4:-1}
set -x
timeout 900 python -m pytest tests/test_gpu_msd.py tests/test_gpu_single_block.py tests/test_gpu_shard.py tests/test_gpu_memory.py tests/test_gpu_blocks.py tests/test_golden.py -m gpu -q -x --tb=short > gpurun_out/${TAG}_pytest.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log | cut -c1-600
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/${TAG}_bench_cfg3_n1.err | grep "^{" > gpurun_out/${TAG}_bench_cfg3_n1.json
tail -c 600 gpurun_out/${TAG}_bench_cfg3_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg3_n1.json"))
print("ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"])
print(d["phases_ms"]); print(d["kernels_ms_per_step"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KERN" -s $SKIP -c $CNT -o gpurun_out/${TAG}_prof_cfg3 -f python tools/profile_step.py --workload cfg3 --scale 1.0 --steps 1 > gpurun_out/${TAG}_ncu2.log 2>&1
tail -n 3 gpurun_out/${TAG}_ncu2.log
# the block path (the reference's algorithm) on the repetitive config: leaves of one copy each hold no long repeats
if [ "${5:-0}" = "1" ]; then
for NB in 64 16; do
  timeout 600 python bench.py --workload cfg4 --numblocks $NB --steps 2 --warmup 3 --no-cpu --e2e-steps 1 2> gpurun_out/${TAG}_bench_cfg4_nb$NB.err | grep "^{" > gpurun_out/${TAG}_bench_cfg4_nb$NB.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_cfg4_nb$NB.json"))
    print("cfg4 nb=$NB ms/step", d["ms_per_step"], d["phases_ms"], d["kernels_ms_per_step"], d["counters"])
except Exception as e:
    print("cfg4 nb=$NB failed", e); print(open("gpurun_out/${TAG}_bench_cfg4_nb$NB.err").read()[-1500:])
PY
done
fi
