# 2 GPUs: parity tests of the position-sharded build, then the bench line
N=${1:-2}; TAG=${2:-r2k}
set -x
timeout 900 python -m pytest tests/test_gpu_multi_abi.py tests/test_gpu_xshard.py -x -q --tb=short 2>&1 | tail -6 | cut -c1-1200
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q --tb=short -k "args4 or args5 or args6 or args7" 2>&1 | tail -6 | cut -c1-1200
bash tools/gpu/r2_n8b.sh $N $TAG
