"""N>1 path: one process per rank, host orchestrator replicated, search sharded (queries for large windows, windows for
the recursion batches; parsnp_b200/csrc/host/sharded.cpp) with the exchange over torch.distributed.
CPU: gloo + the csgmum search (oracle/ref_backend.cpp). GPU: NCCL + the CUDA engine (needs >= 2 GPUs)."""
import json
import os
import socket
import subprocess
import sys

import pytest

from tests.conftest import ROOT

have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpb200_hosttest.so"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, mode, case, extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py"), mode, case] + list(extra)
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = "1"
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900, env=env, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-3000:]
    return json.loads(lines[-1])


@pytest.mark.skipif(not have_ref, reason="oracle/_ref not built")
@pytest.mark.parametrize("world,case", [(2, "rearr_60k"), (3, "pop_30k_x12"), (2, "windows_50k")])
def test_sharded_search_gloo(world, case):
    out = _run(world, "cpu", case)
    assert out["ok"], out
    assert out["counters"][0] >= 1            # at least one query-sharded window (the anchors)
    assert out["counters"][1] >= 1 or case == "pop_30k_x12"
    assert out["calls"]["allreduce"] >= 2 and out["calls"]["allgather"] >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("case,extra", [("rearr_60k", ()), ("pop_30k_x12", ("forcebig",)), ("c1c", ())])
def test_sharded_search_nccl(case, extra):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run(2, "gpu", case, extra)
    assert out["ok"], out
    assert out["calls"]["bcast"] >= 3         # SA, lrp, seed table broadcast from rank 0
