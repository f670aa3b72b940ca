#!/usr/bin/env python
"""bench.py -- depth-evals/s of the statdepth hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload mbd|bd]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload at every N (strong scaling: the total work is BASELINE's configuration, fixed):
  mbd (default) = BASELINE config 2: modified band depth (relax=True, J=2) of 100 000 curves x 1024
                  time points, float64 random walks.  One step = one depth evaluation of ALL curves.
                  With N ranks every rank ranks a contiguous block of T/N time rows (MBD is additive
                  over rows) and the int64 numerators are all-reduced (NCCL).
  bd            = BASELINE config 3: strict band depth of 8192 curves x 512 points (reported when asked).

One JSON line on rank 0.  `value` = depth-evals/s with the input resident in HBM (device-pointer C ABI);
`e2e` = the same through the host-buffer C ABI call (pinned host input, H2D + kernels + D2H + float depth).
`--impl reference` times the CPU oracle port (oracle/sd_oracle.c; the reference itself is pure Python and
cannot travel to the GPU box) on all host threads, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from math import comb

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_BYTES = 126 << 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mbd", choices=["mbd", "bd"])
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--T", type=int, default=None)
    ap.add_argument("--cpu-sample-rows", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bd-impl", default="auto", choices=["auto", "bits", "gemm", "match"])
    ap.add_argument("--nq", type=int, default=None, help="bd workload: number of query curves (default all)")
    ap.add_argument("--ties", action="store_true", help="tie-stress variant: round the random walks to integers")
    return ap.parse_args()


def workload_shape(args):
    if args.workload == "mbd":
        n, T, name = 100_000, 1024, "cfg2: modified band depth (relax=True, J=2), 100k curves x 1024 points"
    else:
        n, T, name = 8192, 512, "cfg3: strict band depth (relax=False, J=2), 8192 curves x 512 points"
    if args.n or args.T:
        n, T = args.n or n, args.T or T
        name += " [overridden to n=%d T=%d]" % (n, T)
    return n, T, name


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


def cpu_baseline(n, T, relax, rows=None, budget_s=20.0):
    """Oracle port on all host threads over a bounded sample (a block of time rows / a set of queries)."""
    from oracle import cpu_oracle
    cpu_oracle.build()
    cores = cpu_oracle.num_threads()
    rng = np.random.default_rng(1)
    if relax:
        rows = min(T, rows or T)  # default: the whole workload (~2 s wall on 16+ cores, ~40 core-seconds)
        X = rng.standard_normal((rows, n)).cumsum(0)
        t0 = time.perf_counter()
        cpu_oracle.mbd_counts_all(X)
        dt = time.perf_counter() - t0
        full = dt * T / rows  # MBD cost is linear in the number of rows
        return dict(value=n / full, unit="depth-evals/s", cores=cores, kind="port",
                    sample="oracle/sd_oracle.c sdo_mbd_counts_all (qsort ranks per row, %d pthreads) on %d of %d time "
                           "rows x %d curves in %.2f s%s" % (cores, rows, T, n, dt,
                                                             "" if rows == T else ", scaled linearly to %d rows" % T),
                    seconds_sample=dt)
    X = rng.standard_normal((T, n)).cumsum(0)
    nq = max(cores, 16)
    q = rng.choice(n, nq, replace=False)
    t0 = time.perf_counter()
    cpu_oracle.bd_counts(X, q)
    dt = time.perf_counter() - t0
    return dict(value=nq / dt, unit="depth-evals/s", cores=cores, kind="port",
                sample="oracle/sd_oracle.c sdo_bd_counts (bit masks, early exit) on %d of %d query curves in %.2f s"
                       % (nq, n, dt), seconds_sample=dt)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, T, name = workload_shape(args)
    relax = args.workload == "mbd"
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(n, T, relax, rows=args.cpu_sample_rows)
        if i >= args.warmup:
            vals.append(cb["value"])
            secs.append(cb["seconds_sample"])
    v = float(np.mean(vals))
    cb.update(value=v)
    cb.pop("seconds_sample", None)
    emit({
        "impl": "reference", "metric": "BD/MBD depth-evals/sec", "value": v, "unit": "depth-evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "curves": n, "time_points": T,
                   "note": "CPU oracle port of the reference path (the reference is pure Python, ~0.05 "
                           "depth-evals/s/core at 200x100, and cannot run at this size); each step is a bounded sample"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "depth-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def emit(line: dict):
    """Print THE one JSON line on the real stdout (see _quiet_stdout)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _quiet_stdout():
    """Libraries (NCCL prints 'NCCL version ...' at init) may write to fd 1: keep fd 1 for the JSON line
    only by pointing it at stderr for the rest of the run."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    args = parse()
    _quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from statdepth_b200 import _dist as sdist
    from statdepth_b200 import _engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, T, name = workload_shape(args)
    relax = args.workload == "mbd"
    eng = E.Engine(local)
    eng.set_option(E.OPT_PROFILE, 1)
    eng.set_option(E.OPT_BD_IMPL, {"auto": E.BD_AUTO, "bits": E.BD_BITS, "gemm": E.BD_GEMM, "match": E.BD_MATCH}[args.bd_impl])
    dev = torch.device("cuda", local)

    # ---- synthetic input: float64 random walks (they cross), identical on every rank ---------------
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    X = torch.empty((T, n), dtype=torch.float64, device=dev)
    for r0 in range(0, T, 128):  # generate in row blocks to bound scratch memory
        X[r0:r0 + 128] = torch.randn((min(128, T - r0), n), dtype=torch.float64, device=dev, generator=g)
    X = X.cumsum(0)
    if args.ties:
        X = X.round()
        name += " [tie stress: rounded to integers]"

    if relax:   # rows sharded, counts all-reduced
        lo, hi = sdist.block(T, rank, world)
        Xl = X[lo:hi].contiguous()
        nq_local = n
        q_dev = None
    else:       # queries sharded, counts all-gathered
        lo, hi = sdist.block(args.nq or n, rank, world)
        Xl = X
        nq_local = hi - lo
        q_dev = torch.arange(lo, hi, dtype=torch.int64, device=dev)
    del X
    Tl = Xl.shape[0]
    out_dev = torch.zeros(max(nq_local, 1), dtype=torch.int64, device=dev)
    host_in = torch.empty(Xl.shape, dtype=torch.float64).pin_memory()
    host_in.copy_(Xl)
    q_host = None if q_dev is None else q_dev.cpu().numpy()
    local_bytes = Xl.numel() * 8
    flush = None
    if local_bytes < 2 * L2_BYTES:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.ExternalStream(eng.stream(), device=dev)

    def collective(t):
        if world == 1:
            return t
        if relax:
            dist.all_reduce(t)
            return t
        parts = [torch.empty_like(t) for _ in range(world)] if n % world == 0 else None
        if parts is None:
            raise SystemExit("bd workload needs n divisible by the number of GPUs")
        dist.all_gather(parts, t)
        return torch.cat(parts)

    # multi-GPU relaxed depth: the device entry point only enqueues (SD_OPT_ASYNC_DEVICE) and the all-reduce is
    # queued behind the kernels on the engine's stream, so a step has one host synchronisation instead of three
    queued = world > 1 and relax
    if queued:
        from statdepth_b200._engine import OPT_ASYNC_DEVICE
        eng.set_option(OPT_ASYNC_DEVICE, 1)

    def step_resident():
        eng.band_depth_counts_dev(Xl.data_ptr(), Tl, n, n, out_dev.data_ptr(),
                                  None if q_dev is None else q_dev.data_ptr(), nq_local, 2, relax)
        if queued:
            with torch.cuda.stream(stream):
                res = collective(out_dev)
            eng.sync()
            return res, eng.timings()
        tm = eng.timings()
        return collective(out_dev), tm

    def step_e2e():
        cnt = eng.band_depth_counts_ptr(host_in.data_ptr(), Tl, n, n, q_host, 2, relax)
        tm = eng.timings()
        if world > 1:
            cnt = collective(torch.from_numpy(cnt).to(dev)).cpu().numpy()
        s = cnt.astype(np.float64)
        depth = (s / float(T) if relax else s) / float(comb(n, 2))  # the float64 depth the API returns
        return depth, tm

    def timed(fn, steps, warmup):
        """W warm-up steps, then K steps timed with CUDA events on the engine's stream (max over ranks)."""
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms, launches, kern_ns, phases, last = 0.0, 0, 0, {}, None
        for _ in range(steps):
            if flush is not None:
                flush.fill_(1)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()  # ranks leave the (untimed) L2 flush together: a step must not wait for a late flush
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            last, tm = fn()
            torch.cuda.synchronize()  # collectives run on torch's stream: drain it before the end event
            e1.record(stream)
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
            launches += tm["launches"] + (1 if world > 1 else 0)
            kern_ns += tm["kernel_ns"]
            for k, v in eng.phase_ns().items():
                phases[k] = phases.get(k, 0) + v
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, kern_ns, phases, last

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, kern_ns, phases, last = timed(step_resident, args.steps, args.warmup)
    ms_e2e, _, _, _, depth = timed(step_e2e, args.steps, max(1, args.warmup))
    clocks = sampler.stop() if rank == 0 else None

    # ---- sanity inside the bench: tie-free checksum of the relaxed numerators ----------------------
    if relax:
        total = int(last.sum().item())
        expect = T * (n * comb(n - 1, 2) - 2 * comb(n, 3))
        if total != expect and not args.ties:
            raise SystemExit("bench: MBD checksum mismatch %d != %d" % (total, expect))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    evals = n if relax else (args.nq or n)  # depth evaluations per step
    value = evals * args.steps / (ms / 1e3)
    e2e_value = evals * args.steps / (ms_e2e / 1e3)
    peak, peak_src = peaks()
    # roofline of the rank pipeline (all kernels of a step): algorithmic bytes = 8*n*T_local read once + 8*n written
    alg_bytes = 8.0 * n * Tl + 8.0 * n
    alg_bytes_placeholder = alg_bytes
    kern_s = kern_ns / 1e9 / args.steps
    achieved = alg_bytes / kern_s / 1e9
    dominant = max(phases, key=phases.get) if phases else None
    # DRAM traffic of one step: ncu --set full on the full cfg2 step (profiles/ncu_r01_v12_summary.md) measured
    # 276 MB (splitters) + 1 619 MB (partition) + 796 MB (rank) + ~75 MB (big parts, finish) = 2 766 MB
    # (sample read + one HBM round trip of the 8-byte part-list entries) = 3.37 bytes per algorithmic byte;
    # scaled to this rank's rows.  null for other shapes.
    traffic = 3.37 * alg_bytes_placeholder if n == 100_000 else None
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per kernel, round-1 v12 capture, scaled by rows",
            "peak_source": peak_src,
            "kernel": "all kernels of one step (per-rank); dominant phase: %s" % dominant,
            "algorithmic_bytes_per_step": alg_bytes, "kernel_ms_per_step": kern_s * 1e3,
            "phase_ms_per_step": {k: v / 1e6 / args.steps for k, v in phases.items()}}
    if dominant:  # the same algorithmic bytes over the dominant phase alone
        dom_s = phases[dominant] / 1e9 / args.steps
        roof["dominant_phase"] = {"name": dominant, "ms_per_step": dom_s * 1e3, "achieved": alg_bytes / dom_s / 1e9,
                                  "frac": alg_bytes / dom_s / 1e9 / peak}
    if not relax:
        # strict BD: int8-tensor-equivalent dense work 2*(2T)*C(n-1,2) ops per depth-eval
        ops = 2.0 * (2 * T) * comb(n - 1, 2) * nq_local
        i8_peak = eng.probe_int8_peak() / 1e12
        roof = {"bound": "tensor", "achieved": ops / kern_s / 1e12, "peak": i8_peak, "unit": "TOP/s (int8 dense-equivalent)",
                "frac": ops / kern_s / 1e12 / i8_peak, "traffic": None,
                "peak_source": "measured here: sd_probe_int8_peak (tcgen05.mma kind::i8 issue loop, operands resident in "
                               "shared memory; nominal dense is 4500).  Only --bd-impl gemm does this many operations; "
                               "for bits / match the figure is a dense-EQUIVALENT rate (they skip the work)",
                "kernel_ms_per_step": kern_s * 1e3,
                "phase_ms_per_step": {k: v / 1e6 / args.steps for k, v in phases.items()}}

    line = {
        "metric": "BD/MBD depth-evals/sec", "value": value, "unit": "depth-evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "curves": n, "time_points": T, "J": 2, "relax": relax,
                   "queries": evals, "bd_impl": args.bd_impl,
                   "sharding": ("time rows over ranks + int64 all-reduce" if relax else
                                "query curves over ranks + all-gather") if world > 1 else "single GPU",
                   "l2": "flushed between steps (256 MiB write)" if flush is not None else
                         "per-rank input (%.0f MB) exceeds L2" % (local_bytes / 1e6)},
        "e2e": {"value": e2e_value, "unit": "depth-evals/s", "h2d_bytes_per_step": int(local_bytes),
                "d2h_bytes_per_step": int(nq_local * 8), "ms_per_step": ms_e2e / args.steps,
                "api": "sd_band_depth_f64 (host buffers, pinned input) + float64 depth on the host"},
        "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
        "engine": {k: v for k, v in eng.timings().items() if k in ("fallback_rows", "bd_impl_used")},
        "depth_checksum": float(np.sum(depth)),
    }
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(n, T, relax, rows=args.cpu_sample_rows)
        cb.pop("seconds_sample", None)
        line["cpu_baseline"] = cb
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
