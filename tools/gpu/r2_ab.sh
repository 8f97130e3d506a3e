# A/B of a variant library against the tree's: the cfg3 line twice each, alternating (variant = bwtb3m_b200/libb3m_v1.so.bin)
TAG=${1:-r2i}
set -x
cp bwtb3m_b200/libb3m.so /tmp/base.so
for R in 1 2; do
for V in base v1; do
if [ $V = base ]; then cp /tmp/base.so bwtb3m_b200/libb3m.so; else cp bwtb3m_b200/libb3m_v1.so.bin bwtb3m_b200/libb3m.so; fi
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-file-level --e2e-steps 2 2> /dev/null | grep "^{" > gpurun_out/${TAG}_bench_${V}_$R.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_${V}_$R.json"))
print("$V $R ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), d["kernels_ms_per_step"])
PY
done
done
cp /tmp/base.so bwtb3m_b200/libb3m.so
