"""GPU, needs >= 2 devices: the NCCL multi-GPU build equals the single-GPU build (skipped on a
one-GPU box; the schedule itself is covered on CPU by tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("args", [["--nsym", "200003", "--itype", "pacterm"], ["--nsym", "150001", "--itype", "bytestream"],
                                  ["--nsym", "300000", "--itype", "pac", "--local-blocks", "2"],
                                  ["--nsym", "200003", "--itype", "pacterm", "--strategy", "merge"],
                                  ["--nsym", "250001", "--itype", "pac", "--strategy", "shard"],
                                  ["--nsym", "20000003", "--itype", "pacterm", "--strategy", "shard"],
                                  ["--nsym", "20000003", "--itype", "pacterm", "--strategy", "shard", "--io"],
                                  ["--nsym", "100003", "--itype", "pacterm", "--strategy", "shard", "--io"],
                                  ["--workload", "cfg3", "--strategy", "shard"]])
def test_nccl_build_equals_single(args):
    if ngpus() < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(ngpus(), 8)  # every visible GPU: the result at N = 4 and 8 is checked, not only timed
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py")] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
