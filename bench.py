#!/usr/bin/env python
"""bench.py — branch logP evaluations / second of the B200 hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c4shard|c2|c3] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path (clock-rate map + spike / rate / stub prior logP + tree prior +
ProportionOfSaltation sums) over one resident synthetic batch: gsm_lean_kernel + gsm_post_kernel per rank; at N>1
the NCCL all-gather of a step's per-tree result rows (the only collective on the path) runs on a second stream and
overlaps the next step's kernels.
Default workload: BASELINE config 4 as written, ONE batch of 1M trees x 2,000 taxa sharded across the N GPUs of the run
(strong scaling: 128 GB of inputs resident on the one GPU at N=1, 125,000 trees = 16 GB per GPU at N=8; far larger than
L2 at every N).  `--workload c4shard` is the weak-scaling twin (125,000 trees per GPU at every N).

Prints ONE JSON line (rank 0).  `--impl reference` instead times the CPU restatement of the reference
(oracle/, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# keep stdout clean for the JSON line: libraries that print to fd 1 (NCCL "NCCL version ...") go to stderr instead
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)

METRIC = "branch_logp_evals_per_sec"
UNIT = "branch evals/s"
BYTES_PER_BRANCH = 32      # height 8 + rate 8 + spike 8 + stub count 4 + parent 4
BYTES_PER_TREE = 72 + 1 + 64  # params row + indicator in, result row out

WORKLOADS = {
    # name: (kind, trees per GPU, taxa, description)
    "c4shard": ("C4", 125_000, 2000, "BASELINE config 4 shard: 125,000 trees x 2,000 taxa per GPU (1M trees at 8 GPUs), "
                                     "BDWS stub counts, rho=1 / rho<1 alternating, spikes on"),
    "c3": ("C3", 100_000, 1000, "BASELINE config 3: 100k trees x 1,000 taxa, FBDWS, sampled ancestors, rho<1, indicator per state"),
    "c2": ("C2", 10_000, 200, "BASELINE config 2: 10k trees x 200 taxa, BDWS, spikes on"),
}
# BASELINE config 4 as written: ONE batch of 1,000,000 trees x 2,000 taxa sharded across the N GPUs of the run (strong scaling:
# 128 GB of inputs resident on one GPU at N = 1, 16 GB per GPU at N = 8).  "c4shard" is its weak-scaling twin (125,000 per GPU).
C4_TOTAL_TREES = 1_000_000
WORKLOADS["c4"] = ("C4", C4_TOTAL_TREES, 2000, "BASELINE config 4: 1,000,000 trees x 2,000 taxa sharded across the GPUs of the run "
                                               "(all of it resident on one GPU at N = 1), BDWS stub counts, rho=1 / rho<1 alternating, spikes on")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s; MEASURED_PEAKS.json absent)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        while not self.stop_flag:
            try:
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        vals = sorted(m for m, _ in self.samples)
        return {"sm_mhz": vals[len(vals) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(vals)}


def csrc_sha():
    """hash of the kernel sources (the counters file under profiles/ is stamped with the hash it was captured from)"""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "gammaspikemodel_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".inc")):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:16]


def kernel_counters(workload):
    """ncu counters of the dominant kernel captured on this workload (profiles/kernel_counters.json, written by
    tools/ncu_counters.py); `stale` says whether the kernel sources have changed since the capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_counters.json")) as f:
            kc = json.load(f)
    except Exception:
        return None
    if kc.get("workload") != workload:
        return None
    kc["stale"] = kc.get("csrc_sha") != csrc_sha()
    return kc


def host_threads():
    """host cores this process may use (the CPU arm runs OpenMP over trees on all of them)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else written to fd 1 (NCCL's version banner ...) was
    redirected to stderr at start-up"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def cpu_time(a, flags, mode, threads, target_s):
    """Time the CPU restatement of the reference (oracle/) on host arrays `a`, repeating the pass until about
    `target_s` seconds of work have been measured.  Returns (branch evals / s, seconds, passes)."""
    import oracle
    T, N = a["height"].shape

    def one():
        oracle.eval_batch(flags, mode, a["height"], a["parent"], a["rate"], a["spike"], a["nstubs"], a["params"],
                          a["use_spike"], n_threads=threads)
    t0 = time.perf_counter()
    one()                                   # warm-up pass, also the speed estimate
    est = time.perf_counter() - t0
    reps = max(1, int(target_s / max(est, 1e-3)))
    t0 = time.perf_counter()
    for _ in range(reps):
        one()
    dt = time.perf_counter() - t0
    return reps * T * N / dt, dt, reps


def host_sample(batch, trees):
    """first `trees` trees of a synthetic.Batch as dense host arrays (the layout the oracle and the C ABI take)"""
    N = batch.n_nodes
    g = lambda x: x[:trees, :N].contiguous().cpu().numpy()
    return dict(height=g(batch.height), parent=g(batch.parent), rate=g(batch.rate), spike=g(batch.spike),
                nstubs=g(batch.nstubs), params=batch.params[:trees].cpu().numpy(), use_spike=batch.use_spike[:trees].cpu().numpy())


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Java/BEAST 2 and
    cannot run here (no JVM), so this is its C restatement (oracle/, kind "port") on all host threads.  One step =
    a bounded sample of the workload (8,192 trees, or fewer for small workloads) evaluated for about 2 s."""
    if rank != 0:
        return
    import oracle
    from gammaspikemodel_b200 import synthetic
    kind, trees_per_gpu, taxa, desc = WORKLOADS[args.workload]
    strong = args.workload == "c4"
    threads = host_threads()   # not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    trees = min(trees_per_gpu, 8192)
    batch = synthetic.make_batch_chunked(kind, trees, taxa, seed=synthetic.SEEDS[kind], chunk=2048)
    a = host_sample(batch, trees)
    n_nodes = batch.n_nodes
    done, total_dt, total_br = 0, 0.0, 0
    for s in range(args.warmup + args.steps):
        v, dt, reps = cpu_time(a, batch.flags, batch.stub_mode, threads, 2.0)
        if s >= args.warmup:
            done += 1
            total_dt += dt
            total_br += reps * trees * n_nodes
    value = total_br / total_dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_dt / max(1, done), "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "trees_in_sample": trees, "taxa": taxa, "nodes_per_tree": n_nodes},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d trees x %d nodes, passes repeated for ~2 s per step (bounded sample of the workload), "
                                       "OpenMP over trees, C restatement of the Java reference (no JVM in this image)" % (trees, n_nodes)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def run_extras(dev, device_index, peak_gbs):
    """The other BASELINE configs on this GPU, each timed with CUDA events on its launch stream after 3 warm-up launches
    (inputs resident in HBM), every one guarded so that a failure here cannot take the headline line down."""
    import torch
    from gammaspikemodel_b200 import GsmEngine, abi, synthetic
    res = []

    def resident(name, kind, trees, taxa, mode=None, flags=None, steps=10, note="", root_at=None):
        try:
            b = synthetic.make_batch_chunked(kind, trees, taxa, seed=synthetic.SEEDS[kind], device=dev, chunk=31250)
            if root_at is not None:
                synthetic.move_root(b, b.n_nodes + root_at if root_at < 0 else root_at)
            e = GsmEngine(device=device_index)
            e.set_mode(b.stub_mode if mode is None else mode, b.flags if flags is None else flags)
            e.bind_batch(b)
            out = torch.empty((b.n_trees, abi.NOUT), dtype=torch.float64, device=dev)
            st = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(st):
                for _ in range(3):
                    e.eval_device(out, None, None, st.cuda_stream)
                st.synchronize()
                l0 = e.launch_count
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(steps):
                    e.eval_device(out, None, None, st.cuda_stream)
                e1.record(st)
                st.synchronize()
            ms = e0.elapsed_time(e1) / steps
            br = b.n_trees * b.n_nodes
            nbytes = br * BYTES_PER_BRANCH + b.n_trees * BYTES_PER_TREE
            res.append({"workload": name, "trees": b.n_trees, "taxa": taxa, "nodes_per_tree": b.n_nodes, "ms_per_step": ms,
                        "value": br / ms * 1e3, "unit": UNIT, "hbm_gbs": nbytes / ms / 1e6, "hbm_frac": nbytes / ms / 1e6 / peak_gbs,
                        "launches_per_step": (e.launch_count - l0) // steps, "finite_frac": float(torch.isfinite(out[:, 0]).double().mean()),
                        "l2": "inputs (%.2f GB) %s L2" % (nbytes / 1e9, "larger than" if nbytes > 256e6 else "fit"), "note": note})
            e.close()
            del b, out
            torch.cuda.empty_cache()
        except Exception as ex:
            res.append({"workload": name, "error": repr(ex)})

    resident("BASELINE config 3: 100k trees x 1,000 taxa, FBDWS, sampled ancestors, rho<1, indicator per state", "C3", 100_000, 1000)
    resident("BASELINE config 2: 10k trees x 200 taxa, BDWS, spikes on", "C2", 10_000, 200, steps=20)
    resident("config 4 trees with the root at node n-2 instead of n-1 (as MCMC operators leave them): 60k trees x 2,000 taxa", "C4", 60_000, 2000,
             root_at=-2, note="the pair of nodes that holds the root goes to the tail lanes, every other pair keeps the ordinary-tree bodies")
    resident("mixture-mode spike prior (Stubs nostub mode, BSP:102-172; the wiring of both BEAUti templates): 20k trees x 1,000 taxa",
             "C3", 20_000, 1000, mode=abi.STUBS_NOSTUB, steps=5, note="fp64-bound: Poisson x Gamma mixture over the stub count of every branch")
    try:
        import tools.bench_proposals as bp
        r = bp.run(64, 4096, 500, reps=5, device=device_index)
        r["unit"] = "proposals/s"
        r["value"] = r["proposals_per_s"]
        r["note"] = ("BASELINE config 5 through gsm_eval_dirty: host CSR arrays (pinned) in, result rows out, wall clock of the "
                     "whole call incl. H2D / D2H; incremental O(dirty) path + full path for scalers / hyper-parameter moves")
        res.append(r)
    except Exception as ex:
        res.append({"workload": "C5", "error": repr(ex)})
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--trees", type=int, default=0, help="override trees per GPU (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how the per-tree result rows reach rank 0: copy engines into rank 0's memory over NVLink (CUDA IPC, "
                         "no SM used) or an NCCL all-gather; p2p falls back to nccl when peer memory cannot be opened")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra workloads (C2, C3, mixture mode, C5) timed after the headline region")
    ap.add_argument("--want-rates", action="store_true", help="also write per-branch effective rates (44 B/branch)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from gammaspikemodel_b200 import GsmEngine, abi, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: --gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # The gather (64 B per tree) runs under the next step's kernels, which fill every SM: keep NCCL's footprint small
        # (a few CTAs move 64 MB in well under one step) instead of letting it take tens of SMs for a short burst.
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "4")
        dist.init_process_group("nccl", device_id=dev)

    kind, trees_per_gpu, taxa, desc = WORKLOADS[args.workload]
    strong = args.workload == "c4"
    if strong:
        trees_per_gpu = C4_TOTAL_TREES // world     # contiguous shards of the one batch (sharding.shard_range)
    if args.trees:
        trees_per_gpu = args.trees
    seed = synthetic.SEEDS[kind] + rank
    t_gen = time.perf_counter()
    batch = synthetic.make_batch_chunked(kind, trees_per_gpu, taxa, seed=seed, device=dev, chunk=31250)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    T, N, S = batch.n_trees, batch.n_nodes, batch.stride
    branches = T * N

    # One step = one evaluation of the whole shard (gsm_lean_kernel + gsm_post_kernel).  At N > 1 the NCCL all-gather of
    # a step's result rows (the only collective on the path, 64 B per tree) is issued on a second stream and overlaps the
    # next step's kernels: three result buffers rotate, and a buffer is rewritten only after its gather has finished.
    # (NCCL's CTAs find room on the SMs mostly between two kernels; with three buffers the kernel of step s+2 does not
    # wait for the gather of step s.)
    nbuf = 3 if distributed else 1
    out_b = [torch.empty((T, abi.NOUT), dtype=torch.float64, device=dev) for _ in range(nbuf)]
    out_rates = torch.empty((T, S), dtype=torch.float64, device=dev) if args.want_rates else None
    peer = None
    if distributed and args.gather == "p2p":
        from gammaspikemodel_b200.sharding import PeerGather
        ok = torch.ones(1, device=dev)
        try:
            peer = PeerGather(rank, world, local_rank, T, abi.NOUT * 8, nbuf)
        except Exception as ex:   # peer memory not available in this environment: every rank falls back together
            sys.stderr.write("bench.py: p2p gather unavailable on rank %d (%r); using nccl\n" % (rank, ex))
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            if peer is not None:
                peer.close()
            peer = None
    gathered = [torch.empty((world * T, abi.NOUT), dtype=torch.float64, device=dev) for _ in range(nbuf)] if (distributed and peer is None) else None
    eng = GsmEngine(device=local_rank)
    eng.set_mode(batch.stub_mode, batch.flags)
    eng.bind_batch(batch)
    engs = [eng]
    # real (non-default) streams: the kernels and the timing events go on `tstream`, the gathers on `cstream`
    tstream = torch.cuda.Stream(device=dev)
    cstream = torch.cuda.Stream(device=dev, priority=-1) if distributed else None
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    done_ev = [torch.cuda.Event() for _ in range(nbuf)]       # step evaluated (kernel stream)
    free_ev = [torch.cuda.Event() for _ in range(nbuf)]       # its rows gathered: the buffer may be rewritten
    launch_total = lambda: sum(e.launch_count for e in engs)
    step_no = [0]

    def step(k_events=None):
        i = step_no[0] % nbuf
        step_no[0] += 1
        if distributed:
            tstream.wait_event(free_ev[i])
        if k_events is not None:
            k_events[0][0].record()
        eng.eval_device(out_b[i], out_rates, None, stream)
        if k_events is not None:
            k_events[0][1].record()
        if distributed:
            done_ev[i].record(tstream)
            with torch.cuda.stream(cstream):
                cstream.wait_event(done_ev[i])
                if peer is not None:
                    peer.push(i, out_b[i].data_ptr(), cstream.cuda_stream)
                else:
                    dist.all_gather_into_tensor(gathered[i], out_b[i])
                free_ev[i].record(cstream)
        return out_b[i]

    gather_note = ("every step's [T,8] result rows are pushed into rank 0's memory by the copy engines over NVLink (CUDA IPC peer "
                   "memory, cudaMemcpyAsync on a second stream: no SM is taken from the kernels; three result buffers); the timed "
                   "region ends after the last push has landed" if peer is not None else
                   "nccl all_gather of the [T,8] result rows of every step, on a second stream: it overlaps the next steps' "
                   "kernels (three result buffers); the timed region ends after the last gather")
    chunks = 1
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    k_ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(chunks)]
            for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = launch_total()
    e0.record()
    out = None
    for s in range(args.steps):
        out = step(k_ev[s])
    if distributed:
        tstream.wait_stream(cstream)      # the timed region ends when the last gather has landed
    e1.record()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches = launch_total() - launches0
    total_ms = e0.elapsed_time(e1)
    kern_ms = sum(a.elapsed_time(b) for evs in k_ev for a, b in evs) / args.steps
    tmax = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms, kern_ms_max = float(tmax[0]), float(tmax[1])
    value = world * branches * args.steps / (total_ms * 1e-3)

    # ---- the gathered rows of the last step on rank 0 are the rows every rank computed ----
    gather_check = None
    if distributed:
        last = (step_no[0] - 1) % nbuf
        mine = out_b[last].view(torch.int64).sum(dim=0)            # per-column checksum of the bit patterns (exact, order-free)
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        if rank == 0:
            if peer is not None:
                g = torch.from_numpy(peer.read_root(last)).to(dev)
            else:
                g = gathered[last]
            got = g.contiguous().view(torch.int64).view(world, T, abi.NOUT).sum(dim=1)
            want = torch.stack(sums)
            same = bool(torch.equal(got, want))
            first_rows = bool(torch.equal(torch.nan_to_num(g[:T]), torch.nan_to_num(out_b[last])))
            gather_check = {"column_checksums_match": same, "rank0_rows_identical": first_rows}
            if not (same and first_rows):
                raise SystemExit("bench.py: gathered result rows differ from the rows the ranks computed")

    # ---- parity spot check of the timed outputs against the oracle (rank 0, 48 trees) ----
    parity = None
    if rank == 0:
        import oracle
        idx = torch.linspace(0, T - 1, 48, device=dev).long()
        sub = lambda x: x[idx].cpu().numpy()
        ref = oracle.eval_batch(batch.flags, batch.stub_mode, sub(batch.height)[:, :N], sub(batch.parent)[:, :N],
                                sub(batch.rate)[:, :N], sub(batch.spike)[:, :N], sub(batch.nstubs)[:, :N],
                                sub(batch.params), sub(batch.use_spike), n_threads=0)["out"]
        got = out[idx].cpu().numpy()
        with np.errstate(all="ignore"):
            err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
        err = np.where((got == ref) | (np.isnan(got) & np.isnan(ref)), 0.0, err)
        parity = {"trees": 48, "max_rel_err_vs_oracle": float(err.max())}
        if not err.max() <= 1e-10:
            raise SystemExit("bench.py: device results differ from the oracle (max rel err %g)" % err.max())

    # ---- end-to-end through the host-buffer C-ABI call (pinned host inputs, copies inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        Te = min(T, 16384)
        pin = lambda x: x[:Te, :N].contiguous().cpu().pin_memory()
        hb = dict(height=pin(batch.height), parent=pin(batch.parent), rate=pin(batch.rate), spike=pin(batch.spike),
                  nstubs=pin(batch.nstubs), params=batch.params[:Te].cpu().pin_memory(),
                  use_spike=batch.use_spike[:Te].cpu().pin_memory())
        hout = torch.empty((Te, abi.NOUT), dtype=torch.float64).pin_memory()
        eng2 = GsmEngine(device=local_rank)
        eng2.set_mode(batch.stub_mode, batch.flags)

        def e2e_step():
            eng2.eval_host_ptrs(Te, N, hb["height"].data_ptr(), hb["parent"].data_ptr(), None, hb["rate"].data_ptr(),
                                hb["spike"].data_ptr(), hb["nstubs"].data_ptr(), hb["params"].data_ptr(),
                                hb["use_spike"].data_ptr(), hout.data_ptr())
        for _ in range(2):
            e2e_step()
        if distributed:
            dist.barrier()
        e2e_steps = max(3, min(args.steps, 10))
        l0 = eng2.launch_count
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        dt = time.perf_counter() - t0
        e2e_launches = eng2.launch_count - l0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        if rank == 0:
            a_, b_ = hout.numpy(), out[:Te].cpu().numpy()
            if not np.array_equal(a_, b_, equal_nan=True):
                bad = np.argwhere(~((a_ == b_) | (np.isnan(a_) & np.isnan(b_))))
                raise SystemExit("bench.py: host-buffer path and device-resident path disagree at %d entries, first %s: %r vs %r"
                                 % (len(bad), bad[0], a_[tuple(bad[0])], b_[tuple(bad[0])]))
        # host -> device ceiling of this box at this N: the same pinned arrays as dense 1-D copies, every rank at once
        dtmp = {k: torch.empty_like(v, device=dev) for k, v in hb.items()}
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            for k, v in hb.items():
                dtmp[k].copy_(v, non_blocking=True)
        torch.cuda.synchronize()
        dtc = time.perf_counter() - t0
        tc = torch.tensor([dtc], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        h2d_bytes = sum(v.numel() * v.element_size() for v in hb.values())
        ceiling = world * 3 * h2d_bytes / float(tc[0]) / 1e9
        del dtmp
        e2e_gbs = world * e2e_steps * (Te * N * BYTES_PER_BRANCH + Te * 73) / dt / 1e9
        e2e = {"value": world * Te * N * e2e_steps / dt, "unit": UNIT,
               "h2d_gbs": e2e_gbs, "h2d_ceiling_gbs": ceiling, "frac_of_h2d_ceiling": e2e_gbs / ceiling,
               "h2d_ceiling_source": "the same pinned host arrays copied as dense 1-D cudaMemcpyAsync by all %d rank(s) at once, "
                                     "3 passes, wall clock, max over ranks (PCIe / host-memory roofline of this box)" % world,
               "fraction_of_a_full_step": Te / T,
               "h2d_bytes_per_step": Te * N * BYTES_PER_BRANCH + Te * 73, "d2h_bytes_per_step": Te * 64,
               "trees_per_step": Te, "steps": e2e_steps, "launches": e2e_launches,
               "note": "gsm_eval_host on pinned host arrays, chunked + double-buffered; bounded batch of the same workload"}
        eng2.close()

    # ---- CPU baseline (rank 0, N=1 only): oracle port on a bounded sample of the same batch ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        threads = host_threads()
        t_all, t_one = min(T, 8192), min(T, 512)
        v_all, dt_all, reps_all = cpu_time(host_sample(batch, t_all), batch.flags, batch.stub_mode, threads, 10.0)
        v_one, dt_one, reps_one = cpu_time(host_sample(batch, t_one), batch.flags, batch.stub_mode, 1, 5.0)
        cpu = {"value": v_all, "unit": UNIT, "cores": threads, "kind": "port", "single_thread_value": v_one,
               "sample": "first %d trees x %d nodes of the batch, %d passes (%.1f s) on %d threads; first %d trees, %d passes "
                         "(%.1f s) on 1 thread; C restatement of the Java reference (no JVM in this image)"
                         % (t_all, N, reps_all, dt_all, threads, t_one, reps_one, dt_one)}

    # ---- fp64 pipe: measured DFMA peak of this GPU (micro-benchmark, outside the timed region) ----
    fp64 = None
    if rank == 0:
        try:
            fma_per_s = eng.measure_fp64_peak()
            kc = kernel_counters(args.workload)
            fp64 = {"peak": 2 * fma_per_s / 1e12, "unit": "TFLOP/s",
                    "peak_source": "gsm_measure_fp64_peak: DFMA micro-benchmark on this GPU in this run (8 independent chains per thread, "
                                   "2,048 threads per SM, best of 5); every fp64-pipe instruction counted as one FMA slot = 2 flop"}
            if kc:
                ipb = kc["fp64_thread_instr_per_branch"]
                ach = 2 * ipb * branches / (kern_ms_max * 1e-3) / 1e12
                fp64.update({"achieved": ach, "frac": ach / fp64["peak"], "fp64_instr_per_branch": ipb,
                             "instr_per_branch": kc.get("thread_instr_per_branch"),
                             "counters_source": kc.get("source"), "counters_csrc_sha": kc.get("csrc_sha"), "counters_stale": kc["stale"]})
        except Exception as ex:   # the headline line must survive a failing side measurement
            fp64 = {"error": repr(ex)}

    # ---- extra workloads (rank 0, N=1): the other BASELINE configs, timed after and outside the headline region ----
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        del out_b, out_rates
        eng.close()
        eng._keep = None
        engs.clear()
        del batch
        torch.cuda.empty_cache()
        extras = run_extras(dev, local_rank, measured_peak()[0])

    if rank == 0:
        peak, peak_src = measured_peak()
        per_launch_bytes = branches * (BYTES_PER_BRANCH + (8 if args.want_rates else 0)) + T * BYTES_PER_TREE
        achieved = per_launch_bytes / (kern_ms_max * 1e-3) / 1e9
        kc = kernel_counters(args.workload)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "trees_per_gpu": T, "trees_total": world * T, "taxa": taxa, "nodes_per_tree": N,
                       "branches_per_step": world * branches, "outputs": "logP+rates" if args.want_rates else "logP",
                       "l2": "inputs per launch (%.1f GB) larger than L2, no flush" % (per_launch_bytes / 1e9)
                       if per_launch_bytes > 256e6 else "inputs fit L2; not flushed",
                       "gather": gather_note if distributed else "none (1 GPU)", "gather_check": gather_check,
                       "launches_per_step": launches // max(1, args.steps),
                       "generate_s": round(t_gen, 1), "parity_sample": parity},
            "clocks": sampler.summary(),
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (kc["dram_bytes_per_launch"] * branches / kc["branches_per_launch"]) if kc else None,
                         "traffic_source": ("%s, %d branches per launch, scaled to this launch's branch count (csrc %s%s)"
                                            % (kc.get("source"), kc.get("branches_per_launch"), kc.get("csrc_sha"),
                                               ", STALE: kernel sources changed since" if kc["stale"] else ", current")) if kc else None,
                         "kernel": "gsm_lean_kernel (kernel_ms also contains gsm_post_kernel, < 2 % of the step)", "kernel_ms": kern_ms_max,
                         "algorithmic_bytes_per_launch": per_launch_bytes, "peak_source": peak_src, "fp64": fp64},
            "cpu_baseline": cpu,
            "extra_workloads": extras,
        }
        emit(line)
    for e in engs:
        e.close()
    if peer is not None:
        dist.barrier()
        peer.close()
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
