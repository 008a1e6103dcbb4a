"""world_size-2 gloo run of the time-sharding exchange logic on CPU (the per-shard arithmetic is the
oracle's; the product CUDA kernels are covered by the -m gpu tests)."""
import os
import subprocess
import sys

from util import ROOT


def test_time_sharding_protocol_gloo_world2():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "sharded_worker.py"), "gloo"]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SHARDED_OK" in out.stdout
