"""torchrun worker for the N>1 tests.
  CPU (gloo):  python -m torch.distributed.run --nproc-per-node 2 tests/dist_worker.py cpu <case>
  GPU (nccl):  python -m torch.distributed.run --nproc-per-node 2 tests/dist_worker.py gpu <case>
Every rank runs the replicated host orchestrator with the search sharded over ranks and compares its result with the
committed golden dump (all ranks must hold the identical result)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                      # noqa: E402
import torch                            # noqa: E402
import torch.distributed as dist        # noqa: E402
from parsnp_b200 import api             # noqa: E402
from tests.conftest import golden_case  # noqa: E402
from tests.refcmp import result_to_dump, diff_dumps   # noqa: E402


def main():
    mode, case = sys.argv[1], sys.argv[2]
    rank = int(os.environ["RANK"]); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if mode == "gpu":
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group("gloo")
    comm = api.TorchComm()
    g, kw, gold = golden_case(case)
    prm = api.make_params(**kw)
    if mode == "gpu":
        if len(sys.argv) > 3 and sys.argv[3] == "forcebig":
            os.environ["PB200_FORCE_PATH"] = "big"
        G = api.Genomes(g, device=local_rank)
        G.set_comm(comm, bcast_index=True)
        res = G.align(prm)
        G.clear_comm()
        G.close()
        counters = None
    else:
        from oracle import hosttest
        lib = hosttest.load()
        lib.pbtest_align_sharded.argtypes = [C.c_int, C.c_int, api.AG_CB, api.AR_CB, api.BC_CB, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p]
        keep, ptrs, lens = api._seq_arrays(g)
        out = C.c_void_p()
        counters = np.zeros(2, np.int64)
        rc = lib.pbtest_align_sharded(comm.rank, comm.world, comm.ag, comm.ar, comm.bc, len(keep), ptrs, api._ptr(lens), C.byref(prm),
                                      C.byref(out), api._ptr(counters))
        assert rc in (0, -5), lib.pb200_last_error()
        res = api.unpack_result(lib, out)
    d = diff_dumps(result_to_dump(res), gold)
    ok = torch.tensor([0 if d else 1], dtype=torch.int32, device="cuda" if mode == "gpu" else "cpu")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"ok": bool(ok.item()), "diff": d[:3], "calls": comm.calls,
                          "counters": counters.tolist() if counters is not None else None, "world": comm.world}))
    dist.destroy_process_group()
    sys.exit(0 if ok.item() else 1)


if __name__ == "__main__":
    main()
