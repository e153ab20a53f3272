#!/usr/bin/env python3
"""bench.py - genome-bases/s through the MUM+LCB path (BASELINE.json metric).

One "step" = one complete pass of the hot path (anchor search over the reference window(s), recursive inter-anchor
search, LCB chaining; reference src/parsnp.cpp:3187-3273) over one synthetic genome set.

  N = 1 : BASELINE.json configs[1] = G_indep(5 Mbp reference + 8 queries, 1 % independent divergence, seed 1)
  N > 1 : weak scaling - every rank runs the same-shaped workload on its own genome set (seed = 1 + rank), i.e. N
          independent partitions as in the reference's partition mode (parsnp:1553-1615); no data-path collective.

`value`  : sum of genome bases / time, genomes already resident in HBM (pb200_align_resident)
`e2e`    : same metric through the user-facing call with HOST buffers (pb200_align: H2D of the genomes + result D2H)
`--impl reference` : the reference's own CPU implementation (oracle/_ref/parsnp_core_ref = unmodified marbl/parsnp
                     built by oracle/build_ref.py) on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if "PB200_NCCL_DEBUG" in os.environ:
    os.environ["NCCL_DEBUG"] = os.environ["PB200_NCCL_DEBUG"]
# stdout carries exactly ONE line, the JSON record: everything else any library prints to fd 1 (NCCL's version banner, ...) goes
# to stderr; the record is written to the saved descriptor at the end
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

L_FULL, NQ, DIV, SEED = 5_000_000, 8, 0.01, 1
L_CPU_SAMPLE = 1_000_000          # cpu_baseline leg: ~20 s of single-thread CPU work
L_REF_STEP = 500_000              # --impl reference: ~7 s per step


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, gpu_index=0):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(L, workdir, nq=NQ, seed=SEED):
    """the reference binary on G_indep(L, nq): returns (bases, mum+lcb seconds, #mums)"""
    from parsnp_b200 import synth
    from oracle import runner
    g = synth.g_indep(L, nq, DIV, seed)
    ref, qs = synth.write_dataset(os.path.join(workdir, "data"), g)
    r = runner.run_ref(ref, qs, os.path.join(workdir, "run"), cores=os.cpu_count() or 1, zero_heap=False)    # timed: stock allocator
    bases = sum(len(x) for x in g)
    return bases, r["mumlcb_seconds"], len(r["dump"]["mums"]) if r["dump"] else 0


REF_ARM_RECORD = os.path.join(tempfile.gettempdir(), "pb200_reference_arm.json")     # left by --impl reference for the b200 arm's cpu_baseline


def _reference_child(L, seed, workdir, out):
    b, sec, nm = cpu_reference_run(L, workdir, seed=seed)
    json.dump({"bases": b, "seconds": sec, "mums": nm}, open(out, "w"))


def run_reference(args, rank, world):
    """The reference arm: the UNMODIFIED reference binary (oracle/_ref/parsnp_core_ref) on the SAME config as the b200 arm - the
    full configs[1] genome set per GPU (seed 1 + r for r < N, as N concurrent processes: what the reference's partition mode
    runs, parsnp:1553-1615).  Its MUM+LCB path is single-threaded and one pass takes 9-15 minutes, so exactly ONE step is
    measured whatever --steps/--warmup say (a deterministic CPU program: no warm-up effect worth 10 minutes); the 500 kbp
    sample of round 1 is kept as a secondary field."""
    if rank != 0:
        return
    L = args.length
    n_proc = max(1, args.gpus)
    with tempfile.TemporaryDirectory() as td:
        t0 = time.time()
        procs = []
        for r in range(n_proc):
            out = os.path.join(td, "r%d.json" % r)
            code = "import bench; bench._reference_child(%d, %d, %r, %r)" % (L, SEED + r, os.path.join(td, "w%d" % r), out)
            procs.append((out, subprocess.Popen([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.DEVNULL)))
        sample = None
        try:
            sb, ssec, _ = cpu_reference_run(L_REF_STEP, os.path.join(td, "sample"))
            sample = {"workload": "G_indep(%d,8,0.01,1)" % L_REF_STEP, "seconds": ssec, "value": sb / ssec, "unit": "bases/s"}
        except Exception as ex:
            sample = {"unavailable": str(ex)}
        recs = []
        for out, p in procs:
            p.wait()
            recs.append(json.load(open(out)))
        wall = time.time() - t0
    sec = max(r["seconds"] for r in recs)                      # the job is done when the slowest partition is
    bases = sum(r["bases"] for r in recs)
    ms = 1000.0 * sec
    val = bases / sec
    line = {"impl": "reference", "metric": "genome_bases_per_sec_mum_lcb", "value": val, "unit": "bases/s", "n_gpus": args.gpus,
            "steps": 1, "warmup": 0, "steps_measured": 1, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: G_indep(%d bp reference + %d queries, 1%% independent divergence, seed 1+rank), "
                                   "ini = template defaults (c=21 d=300 q=30 p=15000000 diagdiff=0.12)" % (L, NQ),
                       "bases_per_step_per_gpu": recs[0]["bases"],
                       "parallelism": "%d concurrent parsnp_core processes (one genome set each), 1 thread each on the MUM+LCB path" % n_proc},
            "cpu_baseline": {"value": val, "unit": "bases/s", "cores": n_proc, "kind": "reference",
                             "sample": "the full workload, one step: %d x G_indep(%d,8,0.01,seed 1+r), MUM+LCB seconds per process %s "
                                       "(wall %.0f s); the MUM+LCB path of parsnp_core is single-threaded (ini cores=%d only affects MUSCLE)"
                                       % (n_proc, L, [round(r["seconds"], 1) for r in recs], wall, os.cpu_count() or 1)},
            "result": {"mums": [r["mums"] for r in recs]},
            "bounded_sample": sample,
            "e2e": {"value": val, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    try:
        json.dump({"length": L, "n_proc": n_proc, "value": val, "seconds": [r["seconds"] for r in recs], "bases": bases,
                   "host_cores": os.cpu_count() or 1, "when": time.time()}, open(REF_ARM_RECORD, "w"))
    except Exception:
        pass
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--length", type=int, default=L_FULL, help="reference length (default = configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="configs1", choices=["configs1", "pop", "c4"],
                    help="configs1 = G_indep(L, 8 q) [default, BASELINE configs[1]]; pop = G_pop(L, --nq) [configs[2] shape]; "
                         "c4 = configs[3] shape: G_pop(--length [50 Mbp], --nq [64]) as 10 contigs, queries N-padded at contig breaks")
    ap.add_argument("--nq", type=int, default=NQ)
    ap.add_argument("--seed", type=int, default=SEED, help="genome set of rank 0 (rank r: seed + r); default = the configs[1] set of the goldens")
    ap.add_argument("--no-pin", action="store_true", help="N>1: leave the ranks' host threads unpinned")
    ap.add_argument("--sharded", action="store_true", help="N>1: one alignment sharded over the ranks instead of one partition per rank")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    # one process per GPU: every rank keeps its host threads (the library's pool inherits the mask) on its own share of the cores,
    # a contiguous block in local-rank order - GPUs and cores of one socket stay together on the usual two-socket boxes
    pinned_cores = None
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
    if local_world > 1 and not args.no_pin and hasattr(os, "sched_setaffinity"):
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // local_world)
            mine = cores[local_rank * per:(local_rank + 1) * per] or cores
            os.sched_setaffinity(0, mine)
            os.environ.setdefault("PB200_HOST_THREADS", str(len(mine)))
            pinned_cores = [mine[0], mine[-1]]
        except OSError:
            pass
    import torch
    from parsnp_b200 import api, synth
    if not torch.cuda.is_available() or not api.cuda_available():
        raise SystemExit("bench.py: no CUDA device - parsnp_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    L = args.length
    sharded = args.sharded and world > 1
    seed = args.seed if sharded else args.seed + rank
    if args.workload == "c4":
        genomes = synth.g_pop(L, args.nq, DIV, seed)
        genomes = [genomes[0]] + [synth.with_contig_padding(g, 10) for g in genomes[1:]]
    elif args.workload == "pop":
        genomes = synth.g_pop(L, args.nq, DIV, seed)
    else:
        genomes = synth.g_indep(L, NQ, DIV, seed)
    bases = sum(len(g) for g in genomes)
    # pinned host copies for the end-to-end leg
    pinned = []
    for g in genomes:
        t = torch.empty(len(g), dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = g
        pinned.append(t)
    host_g = [t.numpy() for t in pinned]
    prm = api.make_params()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    G = api.Genomes(host_g, device=local_rank)
    comm = None
    if sharded:
        comm = api.TorchComm()
        G.set_comm(comm, bcast_index=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = None
    for _ in range(args.warmup):
        res = G.align(prm)
    sampler = ClockSampler(local_rank)
    sampler.start()
    G.reset_timers()
    step_ms = []
    for _ in range(args.steps):
        flush.zero_()                      # evict L2 between timed iterations
        barrier()
        e0.record()
        res = G.align(prm)
        e1.record()
        barrier()
        step_ms.append(e0.elapsed_time(e1))
    timers = G.engine_timers()
    clocks = sampler.stop()
    # end-to-end: host buffers in, results out, through the user-facing call (one untimed call first: the first engine of
    # a process that is created while another one holds the pool's memory has to obtain fresh device memory from the driver)
    e2e_ms = []
    for it in range(1 + max(1, min(args.steps, 3))):
        flush.zero_()
        barrier()
        e0.record()
        if sharded:
            G2 = api.Genomes(host_g, device=local_rank)
            G2.set_comm(comm, bcast_index=True)
            res_e = G2.align(prm)
            G2.close()
        else:
            res_e = api.align(host_g, prm, device=local_rank)
        e1.record()
        barrier()
        if it > 0:
            e2e_ms.append(e0.elapsed_time(e1))
    t_sum = torch.tensor([sum(step_ms), sum(e2e_ms) / len(e2e_ms) * args.steps], dtype=torch.float64, device="cuda")
    per_rank = None
    if dist is not None:
        mine = {"rank": rank, "ms_per_step": sum(step_ms) / args.steps, "t_total_ms": res["stats"]["t_total"] * 1e3,
                "t_replay_ms": res["stats"]["t_replay"] * 1e3, "host_threads": int(res["stats"]["host_threads"]), "cores": pinned_cores}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
        dist.all_reduce(t_sum, op=dist.ReduceOp.MAX)
    tot_ms, tot_e2e_ms = t_sum.tolist()
    ms_per_step = tot_ms / args.steps
    units = 1 if sharded else world          # sharded: the ranks share ONE alignment (strong scaling)
    value = bases * units / (ms_per_step / 1000.0)
    e2e_value = bases * units / (tot_e2e_ms / args.steps / 1000.0)

    peak, peak_src = read_peaks()
    nsteps = args.steps
    # dominant kernel by device time: choose the largest kernel group; algorithmic bytes per SURVEY.md 8(d)
    groups = {k[:-3]: v for k, v in timers.items() if k.endswith("_ms")}
    counts = {k[2:]: v for k, v in timers.items() if k.startswith("n_")}
    gpu_ms_total = sum(groups.values())
    nq = len(genomes) - 1
    # SA build = radix-sort passes of the 21-mer keys: A_sa(r) = 13 + 208 r per window base -> the sort itself moves 8 passes x 24 B
    # MUM scan (seed_extend): A_scan = 20 B per query base (both strands)
    st = res["stats"]
    def per_launch(name):
        c = max(1.0, counts.get(name, 1.0))
        return groups.get(name, 0.0) / c, c
    sort_ms, sort_n = per_launch("index_sort")
    seed_ms, seed_n = per_launch("scan_seed")
    # the recursion = search kernels of the three window classes + the accept kernel of every level
    groups["small_regions"] = groups.get("small_regions", 0.0) + groups.pop("small_b", 0.0) + groups.pop("small_c", 0.0) + groups.pop("small_accept", 0.0)
    counts["small_regions"] = counts.get("small_regions", 0.0) + counts.pop("small_b", 0.0) + counts.pop("small_c", 0.0) + counts.pop("small_accept", 0.0)
    small_ms, small_n = per_launch("small_regions")
    rounds = 1.0 + timers.get("index_rounds", 0.0) / max(1.0, timers.get("big_windows", 1.0))
    roof_kernels = {
        # one launch group = all LSD passes of one window's packed seed keys; accounted at SURVEY 8(d)'s sort share of A_sa: 8 passes x (12 B read + 12 B write) per window base
        # (the 2-bit/32-bit key variant moves fewer bytes than this; the figure is the reference-algorithm byte count, not bytes moved)
        "sa_build_radix_sort(tile_hist+digit_scan+scatter_kernel passes)": {"bytes": 192.0 * timers.get("big_ref_bases", 0.0) / sort_n, "ms": sort_ms},
        # A_scan = 20 B per query base (both strands)
        "mum_scan(seed_extend_kernel)": {"bytes": 20.0 * timers.get("big_query_bases", 0.0) / seed_n, "ms": seed_ms},
        # SURVEY 8(d) "recursion: same formulas applied to the sum of region lengths": A_sa(r=1) = 221 B per window base + 20 B per query base
        "recursion(recursion_level_kernel+recursion_accept_kernel)": {"bytes": (221.0 * timers.get("small_ref_bases", 0.0) + 20.0 * timers.get("small_query_bases", 0.0)) / small_n,
                                           "ms": small_ms},
    }
    dom_total = {"sa_build_radix_sort(tile_hist+digit_scan+scatter_kernel passes)": groups.get("index_sort", 0.0), "mum_scan(seed_extend_kernel)": groups.get("scan_seed", 0.0),
                 "recursion(recursion_level_kernel+recursion_accept_kernel)": groups.get("small_regions", 0.0)}
    dom = max(dom_total, key=lambda k: dom_total[k])
    rk = roof_kernels[dom]
    ach = rk["bytes"] / (rk["ms"] / 1000.0) / 1e9 if rk["ms"] > 0 else 0.0
    # DRAM bytes per launch of the dominant kernel from the committed ncu capture of this workload (tools/ncu_traffic.py)
    traffic, traffic_src, issue = None, None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        knames = {"recursion(recursion_level_kernel+recursion_accept_kernel)": ["recursion_level_kernel", "recursion_accept_kernel"],
                  "mum_scan(seed_extend_kernel)": ["seed_extend_kernel"],
                  "sa_build_radix_sort(tile_hist+digit_scan+scatter_kernel passes)": ["tile_hist_kernel", "digit_scan_kernel", "scatter_kernel"]}.get(dom, [])
        have = [tj["kernels"][k] for k in knames if k in tj["kernels"]]
        if args.workload == "configs1" and L == L_FULL and have and len(have) == len(knames):
            # per launch like `achieved`: the group's DRAM bytes over one step / its launches in that step
            traffic = sum(k["dram_bytes"] for k in have) / max(1, sum(k["launches"] for k in have))
            traffic_src = "profiles/ncu_traffic.json (" + tj["source"] + ")"
            if "issue_active_pct" in tj and dom in tj["issue_active_pct"]:
                issue = tj["issue_active_pct"][dom]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                # the recursion kernels are instruction / shared-memory bound by construction (DRAM < 1 %): beside the byte ruler of
                # SURVEY 8(d), the issue-slot utilisation of their main launches from the committed `ncu --set full` capture
                "issue_slot_utilisation": issue,
                "algorithmic_bytes_per_launch": rk["bytes"], "ms_per_launch": rk["ms"],
                "share_of_gpu_time": dom_total[dom] / gpu_ms_total if gpu_ms_total else 0.0,
                "other": {k: {"GBps": (v["bytes"] / (v["ms"] / 1000.0) / 1e9 if v["ms"] > 0 else 0.0), "ms": v["ms"]}
                          for k, v in roof_kernels.items()},
                "gpu_ms_per_step_by_group": {k: v / nsteps for k, v in groups.items()},
                "gpu_ms_per_step": gpu_ms_total / nsteps}
    line = {"metric": "genome_bases_per_sec_mum_lcb", "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": ("configs[1]: G_indep(%d bp reference + %d queries, 1%% independent divergence, seed %d+rank), "
                                    "ini = template defaults (c=21 d=300 q=30 p=15000000 diagdiff=0.12)" % (L, NQ, args.seed)) if args.workload == "configs1"
                       else (("configs[2] shape: G_pop(%d bp reference + %d queries, 1%% divergence, seed 1), ini = template defaults" % (L, args.nq))
                             if args.workload == "pop" else
                             ("configs[3] shape: G_pop(%d bp reference in 10 contigs + %d queries of 10 contigs with 310-N padding, 1%% divergence), "
                              "p=15000000 (%d reference windows), ini = template defaults" % (L, args.nq, (L + 14999999) // 15000000))),
                       "bases_per_step_per_gpu": bases, "l2": "flushed between timed steps (256 MiB write)",
                       "parallelism": ("one alignment sharded over %d GPUs (queries / windows), NCCL exchange" % world) if sharded
                       else ("1 partition per GPU" if world > 1 else "single GPU")},
            "e2e": {"value": e2e_value, "unit": "bases/s", "h2d_bytes_per_step": int(bases),
                    "d2h_bytes_per_step": int(st["candidates"] * (8 + 5 * nq))},
            "gpu_launches": int(timers.get("kernel_launches", 0)),
            "clocks": clocks, "roofline": roofline,
            "result": {"mums": int(len(res["mum_length"])), "lcbs": int((res["cluster_type"] == 1).sum()),
                       "anchors": int(st["anchors"]), "regions_searched": int(st["regions_searched"]),
                       "replay_misses": int(st["replay_misses"]), "spec_levels": int(st["spec_levels"]),
                       "replay": {k: int(st[k]) for k in ("replay_tasks", "replay_workers", "replay_foreign_reads", "replay_foreign_writes",
                                                          "replay_restarts", "replay_fallback", "spec_deferred", "slow_queue_iters", "replay_gaps", "replay_final_gaps",
                                                          "replay_final_mums") if k in st}},
            "host_seconds": {k: st[k] for k in st if k.startswith("t_")},
            "engine": {k: timers[k] for k in timers if not k.endswith("_ms") and not k.startswith("n_")},
            "small_class_ms": {"a": timers.get("small_regions_ms", 0.0) / nsteps, "b": timers.get("small_b_ms", 0.0) / nsteps,
                               "c": timers.get("small_c_ms", 0.0) / nsteps,
                               "accept": timers.get("small_accept_ms", 0.0) / nsteps}}
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        full = None
        try:
            full = json.load(open(REF_ARM_RECORD))
            if full.get("length") != L or full.get("n_proc") != 1 or args.workload != "configs1" or time.time() - full.get("when", 0) > 6 * 3600:
                full = None
        except Exception:
            full = None
        if full:
            # `bench.py --impl reference` ran on this box just before (the driver's order): quote that run - the SAME config
            line["cpu_baseline"] = {"value": full["value"], "unit": "bases/s", "cores": 1, "kind": "reference",
                                    "sample": "the full workload (configs[1], G_indep(%d,8,0.01,1)): %.1f s for %d bases, measured by "
                                              "`bench.py --impl reference` on this box (%d host cores; the reference's MUM+LCB path is "
                                              "single-threaded)" % (L, full["seconds"][0], full["bases"], full["host_cores"])}
        else:
            try:
                with tempfile.TemporaryDirectory() as td:
                    b, sec, nm = cpu_reference_run(L_CPU_SAMPLE, td)
                gold = {}
                try:
                    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "full_size.json"))).get("c2_indep_5m_8q_stock_heap", {})
                except Exception:
                    pass
                line["cpu_baseline"] = {"value": b / sec, "unit": "bases/s", "cores": 1, "kind": "reference",
                                        "sample": "oracle/_ref/parsnp_core_ref on G_indep(%d,8,0.01,1): %.1f s for %d bases; the reference's MUM+LCB "
                                                  "path is single-threaded (host has %d cores). The full workload is timed by `bench.py --impl "
                                                  "reference` (one 9-15 min step); on the build container it took %s s (tests/golden/full_size.json)"
                                                  % (L_CPU_SAMPLE, sec, b, os.cpu_count() or 1, gold.get("reference_mumlcb_seconds"))}
            except Exception as ex:  # the oracle binary is test infrastructure; absence must not break the bench
                line["cpu_baseline"] = {"value": None, "unit": "bases/s", "cores": 1, "kind": "reference", "sample": "unavailable: %s" % ex}
    if per_rank is not None:
        line["per_rank"] = per_rank
    if rank == 0:
        emit(line)
    G.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
