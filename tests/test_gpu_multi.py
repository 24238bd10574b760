"""Multi-GPU path (needs >= 2 GPUs; skipped otherwise): sharded population + NCCL reduction/gather through torch.distributed
and through rmt_comm_*, sharded dynamic ensemble."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_sharded_population_matches_unsharded():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    env = dict(os.environ, B="8192")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tools", "dist_check.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    # population through torch.distributed, the same through the C ABI's own NCCL transport (rmt_comm_*), and a sharded
    # dynamic ensemble with the packed gather of the final profiles
    assert res["ok"] is True and res["ok_population"] and res["ok_rmt_comm_transport"] and res["ok_n2_sharded"], res


def test_single_process_sharded_api_equals_batch_api():
    import numpy as np
    import cases
    from rmt_app_b200 import ensemble, rmtExeBatch
    base = cases.methanol_readme_input("N1")
    sw = cases.config3_sweep(1000, seed=8)
    a = ensemble.rmtExeBatchSharded(base, sw)
    b = rmtExeBatch(base, sw)
    np.testing.assert_array_equal(a["dataYs"], b["dataYs"])
    assert a["failed"] == 0 and a["range"] == (0, 1000)
