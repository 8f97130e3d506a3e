set -x
# the two prepared experiments (experiments/README.md), one GPU, about 4 minutes
make -s experiments 2>&1 | tail -2
timeout 120 bin/exp_rec64_pass 28 2>&1 | tee gpurun_out/exp_rec64_pass.log
# deferred second keys in k_resolve: parity first (single-block, file-level and full-size tests), then the kernel times
B3M_EXPERIMENT_RESOLVE_DEFER=1 timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu_defer.log 2>&1; tail -4 gpurun_out/pytest_gpu_defer.log
timeout 300 python bench.py --no-cpu 2> gpurun_out/bench_base.err | grep "^{" > gpurun_out/bench_base.json
B3M_EXPERIMENT_RESOLVE_DEFER=1 timeout 300 python bench.py --no-cpu 2> gpurun_out/bench_defer.err | grep "^{" > gpurun_out/bench_defer.json
python - <<'PY'
import json
for n in ("base", "defer"):
    d = json.load(open("gpurun_out/bench_%s.json" % n))
    print(n, d["ms_per_step"], d["kernels_ms_per_step"], d["e2e"]["ms_per_step"])
PY
