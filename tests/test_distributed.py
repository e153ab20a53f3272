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


@pytest.mark.skipif(not have_ref, reason="oracle/_ref not built")
@pytest.mark.parametrize("world,case", [(2, "rearr_60k"), (4, "rearr_60k"), (3, "windows_50k")])
def test_sharded_search_thread_ranks(world, case):
    """the same N>1 host path with the ranks as threads of this process (oracle/hosttest.py ThreadRanks): every rank == golden.
    rearr_60k has inversions, i.e. regions that overlap in a query genome: the case in which a speculative accept racing on the
    scratch layout made the ranks ask for different windows (found by FUZZ_WORLD=2 tools/fuzz_host.py)"""
    from oracle import hosttest
    from parsnp_b200 import api
    from tests.conftest import golden_case
    from tests.refcmp import result_to_dump, diff_dumps
    g, kw, gold = golden_case(case)
    for _ in range(3):
        outs, counters = hosttest.ThreadRanks(world).align(g, api.make_params(**kw))
        for o in outs:
            assert diff_dumps(result_to_dump(o), gold) == []
    assert counters[0] >= 1 and counters[1] >= 1


@pytest.mark.skipif(not have_ref, reason="oracle/_ref not built")
def test_sharded_search_refuses_ranks_out_of_step():
    """ranks that arrive at a search with different window lists (here: different ini values) get an error, not each other's
    candidates (parsnp_b200/csrc/host/sharded.cpp, the check at the top of ShardedBackend::search)"""
    from oracle import hosttest
    from parsnp_b200 import api
    from tests.conftest import golden_case
    g, kw, _ = golden_case("windows_50k")
    other = dict(kw)
    other["q"] = kw.get("q", 30) + 40
    with pytest.raises(RuntimeError, match="not searching the same windows"):
        hosttest.ThreadRanks(2).align(g, api.make_params(**kw), params_of_rank={1: api.make_params(**other)})


@pytest.mark.skipif(not have_ref, reason="oracle/_ref not built")
def test_sharded_host_fuzz_thread_ranks():
    """tools/fuzz_host.py with FUZZ_WORLD=2: 12 random cases, every rank of the sharded path == the single-rank result == the
    reference binary (seed 51031 is the case of test_sharded_search_thread_ranks' docstring)"""
    env = dict(os.environ, MALLOC_PERTURB_="255", GLIBC_TUNABLES="glibc.malloc.tcache_count=0", PB200_HOST_THREADS="2", FUZZ_WORLD="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_host.py"), "51024", "12"], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "done 12 cases, 0 mismatches" in r.stdout, r.stdout[-2000:]
