# K5 gap counting: tests of both modes, then the block path (4 blocks, 775 Mbp) with each, and an ncu capture of the gap kernels
TAG=${1:-r2g}
set -x
timeout 900 python -m pytest tests/test_gpu_blocks.py -m gpu -q -x --tb=short 2>&1 | tail -6 | cut -c1-800
for GM in atomic list; do
timeout 600 python bench.py --workload cfg3 --scale 0.25 --numblocks 4 --gapmode $GM --steps 3 --warmup 3 --no-cpu --no-file-level --e2e-steps 1 2> gpurun_out/${TAG}_bench_cfg3q_nb4_$GM.err | grep "^{" > gpurun_out/${TAG}_bench_cfg3q_nb4_$GM.json
tail -c 400 gpurun_out/${TAG}_bench_cfg3q_nb4_$GM.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg3q_nb4_$GM.json"))
print("$GM ms/step", d["ms_per_step"], d["phases_ms"]); print(d["kernels_ms_per_step"]); print(d["counters"])
PY
done
timeout 600 ncu --set full --clock-control none -k regex:'k_gap|k_gap_hist|k_radix_onesweep' -c 6 -o gpurun_out/${TAG}_prof_gap -f python tools/profile_step.py --workload cfg3 --scale 0.25 --steps 1 --numblocks 4 > gpurun_out/${TAG}_ncu.log 2>&1
tail -n 3 gpurun_out/${TAG}_ncu.log
