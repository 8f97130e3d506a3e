# A/B of the finish kernel's CTA shape: tests + short cfg3 bench line for 512 and 1024 threads, ncu of the default
TAG=${1:-r2c}
set -x
for T in 512 1024; do
  export B3M_FIN_THREADS=$T
  timeout 900 python -m pytest tests/test_gpu_msd.py tests/test_gpu_single_block.py tests/test_gpu_shard.py tests/test_golden.py -m gpu -q -x --tb=short > gpurun_out/${TAG}_pytest_$T.log 2>&1
  tail -4 gpurun_out/${TAG}_pytest_$T.log | cut -c1-600
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-file-level 2> gpurun_out/${TAG}_bench_cfg3_n1_$T.err | grep "^{" > gpurun_out/${TAG}_bench_cfg3_n1_$T.json
  tail -c 400 gpurun_out/${TAG}_bench_cfg3_n1_$T.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_cfg3_n1_$T.json"))
print("T=$T ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"])
print(d["kernels_ms_per_step"])
PY
done
for T in 512 1024; do
export B3M_FIN_THREADS=$T
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_msd_finish" -s 0 -c 1 -o gpurun_out/${TAG}_prof_cfg3_$T -f python tools/profile_step.py --workload cfg3 --scale 1.0 --steps 1 > gpurun_out/${TAG}_ncu_$T.log 2>&1
tail -n 2 gpurun_out/${TAG}_ncu_$T.log
done
