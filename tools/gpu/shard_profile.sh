set -x
timeout 600 python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -5
timeout 300 python tools/shard_profile.py --nparts 8 --part 3
timeout 300 python tools/shard_profile.py --nparts 2 --part 1
