"""Row e on real GPUs (needs >= 2 visible devices; skipped on a single-GPU box): NCCL path and fused peer-store path of
the column-sharded layer against the unsharded layer, bit for bit.  See tests/multi_worker.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_layer_nccl_and_fused_gather_equal_the_unsharded_layer():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (covered on CPU with gloo in test_sharding_cpu.py and on one GPU by "
                    "test_column_shards_reproduce_the_unsharded_layer_on_the_gpu)")
    world = 2
    env = dict(os.environ)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_worker.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "MULTI_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
