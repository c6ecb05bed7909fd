#!/usr/bin/env python
"""bench.py -- block-Lanczos mod-p hot path on B200: Lanczos iterations/s and SpMV G(nnz*n)/s.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg4|cfg4_small|cfg2|cfg3]

A "step" is one block-Lanczos iteration (two n-wide sparse products, block dot products,
semi-inverse, orthogonalize) of the named workload.  Default workload (BASELINE.json configs[3],
the one the metric's 1/2/4/8-GPU scaling and HBM-roofline targets are quoted on; it fits one
GPU): synthetic 50M x 50M power-law matrix, ~1.5e9 non-zeros, p = 2^31-1, n = 16, generated on
the device with a seeded torch generator.  With N > 1 (torchrun) the same matrix is row-sharded
over the ranks ("strong" scaling) and vector blocks are exchanged with NCCL.

One JSON line on stdout (rank 0).  `value` = iterations/s with everything resident in HBM,
device-timed with CUDA events on the library's stream, max over ranks.  `e2e` = the same through
the C ABI with host buffers: blk_set_state (H2D of the start block from pinned memory), K
blk_iterate(1) calls (each returns the iteration counter and stop flag to the host) and
blk_get_state (D2H of v), all inside the timed region.  `roofline` is for the SpMV kernel
(k_spmv), algorithmic bytes per SURVEY.md section 8(d) over the in-run CUDA-event time of its
launches.  `cpu_baseline` / `--impl reference` time the UNMODIFIED reference's OpenMP build
(oracle/_ref, compiled from /root/reference) on the host cores on a scaled-down twin.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_MERSENNE = 2147483647
WORKLOADS = {
    # name: rows, cols, mean nnz/row, n, p, right
    "cfg4": dict(rows=50_000_000, cols=50_000_000, mean=30.0, n=16, p=P_MERSENNE, right=False,
                 desc="BASELINE configs[3]: synthetic 50Mx50M power-law, ~1.5B nnz, p=2^31-1, n=16"),
    "cfg4_small": dict(rows=5_000_000, cols=5_000_000, mean=30.0, n=16, p=P_MERSENNE, right=False,
                       desc="1/10-scale twin of configs[3] (dev)"),
    "cfg4_tiny": dict(rows=200_000, cols=200_000, mean=30.0, n=16, p=P_MERSENNE, right=False,
                      desc="1/250-scale twin of configs[3] (CPU-sized)"),
}
METRIC = "lanczos_iters_per_s"
UNIT = "iterations/s"
CPU_TWIN = dict(rows=200_000, cols=200_000)          # bounded CPU sample of the same generator


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
def gen_device_coo(torch, w, dev, seed=1):
    """Power-law row degrees (Pareto shape 1.5, mean ~`mean`, cap 1e6), uniform columns, values in
    [1,100): SURVEY.md section 8(d).  Same generator family as synth.powerlaw_rows, on the device."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    N, Mc = w["rows"], w["cols"]
    u = torch.rand(N, device=dev, generator=g, dtype=torch.float64)
    dmin = max(1, round(w["mean"] / 3))
    d = torch.clamp((dmin * (1 - u) ** (-1 / 1.5)).floor().to(torch.int64), max=min(1_000_000, 8 * Mc))
    del u
    rows = torch.repeat_interleave(torch.arange(N, device=dev, dtype=torch.int32), d)
    del d
    nnz = rows.numel()
    cols = torch.randint(0, Mc, (nnz,), device=dev, generator=g, dtype=torch.int32)
    vals = torch.randint(1, 100, (nnz,), device=dev, generator=g, dtype=torch.int32)
    return rows, cols, vals, nnz


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------
def run_cpu_reference(w, iters, warmup, threads=None):
    """Reference OpenMP build on the host cores, on the CPU twin of workload w (see oracle/ref_runner.py)."""
    cores = os.cpu_count() or 1
    cand = [threads] if threads else sorted({min(cores, 16), cores})
    best = None
    for t in cand:
        env = dict(os.environ, OMP_NUM_THREADS=str(t), OMP_STACKSIZE="1G", OMP_PROC_BIND="false")
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "--lib", "omp",
               "--rows", str(CPU_TWIN["rows"]), "--cols", str(CPU_TWIN["cols"]), "--mean", str(w["mean"]),
               "--n", str(w["n"]), "--prime", str(w["p"]), "--right", str(int(w["right"])),
               "--iters", str(iters), "--warmup", str(warmup)]
        try:
            out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
            r = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception as e:                                   # pragma: no cover
            sys.stderr.write(f"[bench] CPU reference run failed with {t} threads: {e}\n")
            continue
        if best is None or r["iters_per_s"] > best["iters_per_s"]:
            best = r
    return best


def cpu_baseline_obj(r, w, nnz_full):
    """Scale the twin's measured rate to the full workload by the work ratio (both the sparse
    products, ~nnz*n, and the dense phases, ~N*n^2, are linear in the twin's scale factor)."""
    scale = r["nnz"] / float(nnz_full)
    return {
        "value": r["iters_per_s"] * scale, "unit": UNIT, "cores": r["threads"], "kind": r["kind"],
        "sample": (f"{r['lib']} build of the reference, {r['iters']} iterations on a {r['rows']}x{r['cols']} twin "
                   f"({r['nnz']} nnz, same generator, n={r['n']}, p={r['prime']}): {r['iters_per_s']:.3f} it/s "
                   f"= {r['gnnzn_per_s']:.3f} G(nnz*n)/s measured; value = that x nnz_twin/nnz_full ({scale:.3e}). "
                   "-DNDEBUG build: at p=2^31-1 the OpenMP variant's deferred modulo overflows u64 "
                   "(SURVEY F2), so it is a timing baseline only."),
        "measured_iters_per_s_on_sample": r["iters_per_s"], "spmv_gnnzn_per_s": r["gnnzn_per_s"],
    }


def expected_nnz(w):
    # mean of floor(10 * U^(-2/3)) capped at 1e6, measured on the generator: 29.33 per row
    return int(w["rows"] * 29.33 * (w["mean"] / 30.0))


# ------------------------------------------------------------------------------------------
def dense_roofline(phases_ms_per_step, rows_local, n_pad, peak):
    """Secondary kernels against the same HBM peak: algorithmic bytes of SURVEY 8(d) (block_dot_products reads v and
    Av: 2 blocks; orthogonalize reads v, Av, p and writes v, p: 5 blocks of rows x n_pad x 4 bytes) over the in-run
    phase times.  On one GPU the dots phase includes the fused n x n stage (one block, ~25 us)."""
    out = {}
    block = rows_local * n_pad * 4
    for key, blocks, kernel in (("dots", 2, "block_dot_products (+ n x n stage)"), ("ortho", 5, "orthogonalize")):
        ms = float(phases_ms_per_step.get(key, 0.0) or 0.0)
        if ms > 0 and block > 0:
            ach = blocks * block / (ms * 1e-3) / 1e9
            out[key] = {"kernel": kernel, "bound": "hbm", "bytes_per_launch": blocks * block, "ms_per_launch": ms,
                        "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("BLK_BENCH_WORKLOAD", "cfg4"), choices=sorted(WORKLOADS))
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    w = WORKLOADS[a.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{a.workload}: {w['desc']}", "rows": w["rows"], "cols": w["cols"], "n": w["n"],
              "prime": w["p"], "side": "right" if w["right"] else "left", "sharding": f"row-blocks x{world}",
              "l2": "inputs (matrix 12 GB/operator, 3.2 GB blocks) far larger than the 126 MB L2; no flush needed"
                    if a.workload == "cfg4" else "working set exceeds L2"}

    # ---------------- reference arm: the reference's own CPU implementation --------------
    if a.impl == "reference":
        if rank != 0:
            return 0
        r = run_cpu_reference(w, max(1, a.steps), max(0, a.warmup))
        nnz_full = expected_nnz(w)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "reference run failed on this host"}))
            return 0
        cb = cpu_baseline_obj(r, w, nnz_full)
        line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "impl": "reference", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u32 (u64 accumulate)", "data": "synthetic",
                "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------- our arm ---------------------------------------------------------------
    import numpy as np
    import torch
    import blk_lanczos_b200 as B
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        ident = [B.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        nccl_id = ident[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t0 = time.time()
    rows, cols, vals, nnz = gen_device_coo(torch, w, dev)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    stream = torch.cuda.Stream(device=dev)
    t0 = time.time()
    ctx = B.BlockLanczos(n=w["n"], prime=w["p"], right=w["right"], device=local_rank, rank=rank, world=world,
                         nccl_id=nccl_id, stream=stream.cuda_stream, chunk_len=a.chunk,
                         device_coo=(w["rows"], w["cols"], nnz, rows.data_ptr(), cols.data_ptr(), vals.data_ptr()))
    torch.cuda.synchronize()
    t_build = time.time() - t0
    del rows, cols, vals
    torch.cuda.empty_cache()
    info = ctx.info()
    n, p = w["n"], w["p"]
    N, Mc = info["N"], info["Mc"]

    # start block: pinned host memory, uniform residues (the reference's xoshiro start block is a
    # host-side sequential generator, sequential/lanczos_modp.c:624-625; for a throughput run any
    # full-rank start is equivalent)
    v_host = torch.empty(N * n, dtype=torch.int32, pin_memory=True)
    v_host.random_(0, p, generator=torch.Generator().manual_seed(7))
    v_np = v_host.numpy().view(np.uint32)
    out_host = torch.empty(ctx.pad, dtype=torch.int32, pin_memory=True)
    out_np = out_host.numpy().view(np.uint32)

    ctx.set_state(v_np)
    ctx.iterate(a.warmup)                       # warm-up steps (untimed)

    # -------- device-resident timed region: K steps, CUDA events on the library's stream
    sampler = ClockSampler(local_rank)
    ctx.set_profiling(True)
    launches0 = ctx.kernel_launches()
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        it, stopped = ctx.iterate(a.steps)
        e1.record(stream)
    barrier()
    clocks = sampler.finish()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.kernel_launches() - launches0
    phases = ctx.phase_times()
    ctx.set_profiling(False)
    assert not stopped, "the synthetic run hit the termination condition inside the timed region"
    value = a.steps / (ms_total / 1e3)

    # -------- end to end through the C ABI with host buffers
    e2e = None
    if not a.no_e2e:
        barrier()
        t0 = time.perf_counter()
        ctx.set_state(v_np)                                   # H2D of the start block (pinned)
        t1 = time.perf_counter()
        for _ in range(a.steps):
            ctx.iterate(1)                                    # returns (n_iterations, stopped) to the host
        t2 = time.perf_counter()
        ctx.L.blk_get_state(ctx.h, out_np.ctypes.data, None, None, None)     # D2H of v
        barrier()
        t3 = time.perf_counter()
        dt = max_over_ranks(t3 - t0)
        e2e = {"value": a.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(N * n * 4 / a.steps + 32), "d2h_bytes_per_step": int(ctx.pad * 4 / a.steps + 32),
               "seconds": {"set_state": t1 - t0, "iterate": t2 - t1, "get_state": t3 - t2},
               "note": "blk_set_state + K x blk_iterate(1) + blk_get_state(v); state copies amortised over K"}

    # -------- roofline of the SpMV kernel (both products), SURVEY 8(d) algorithmic bytes
    peak, peak_src = peaks()
    nnz1, nnz2 = info["nnz_local"]
    r1, r2 = info["local_M1"] - info["local_M0"], info["local_N1"] - info["local_N0"]
    b1 = 8 * nnz1 + 4 * (r1 + 1) + 4 * n * N + 4 * n * r1
    b2 = 8 * nnz2 + 4 * (r2 + 1) + 4 * n * Mc + 4 * n * r2
    g1 = nnz1 * (8 + 4 * n) + 4 * (r1 + 1) + 4 * n * r1
    g2 = nnz2 * (8 + 4 * n) + 4 * (r2 + 1) + 4 * n * r2
    t_spmv = (phases["spmv1"]["ms"] + phases["spmv2"]["ms"]) / a.steps          # ms per iteration, this rank
    achieved = (b1 + b2) / (t_spmv * 1e-3) / 1e9
    achieved_g = (g1 + g2) / (t_spmv * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("workload") == a.workload and tj.get("n_gpus", 1) == world:
            traffic = tj.get("dram_bytes_per_launch")
    dram = None
    if traffic:
        # both products move the same bytes to within 1%; `traffic` is one launch
        dram = {"bytes_per_launch": traffic, "achieved": 2 * traffic / (t_spmv * 1e-3) / 1e9,
                "frac": 2 * traffic / (t_spmv * 1e-3) / 1e9 / peak,
                "note": "ncu dram__bytes_read+write of one k_spmv launch (profiles/r01_spmv_cfg4_ncu_summary.txt): "
                        "every L2 miss fills a 128 B line, so a 64 B x-row gather costs 128 B of HBM traffic "
                        "(profiles/r01_gather_granularity.txt)"}
    roofline = {"kernel": "k_spmv (both products of one iteration)", "bound": "hbm", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "algorithmic_bytes_per_iteration": b1 + b2, "traffic": traffic, "dram": dram,
                "gather_model": {"bytes_per_iteration": g1 + g2, "achieved": achieved_g, "frac": achieved_g / peak,
                                 "note": "x block (3.2 GB) >> L2: every non-zero gathers its own 4n-byte x row"},
                "ms_per_launch_pair": t_spmv, "frac_of_nominal_8TBs": achieved / 8000.0}
    spmv_rate = (nnz1 + nnz2) * n / (t_spmv * 1e-3) / 1e9
    if world > 1:
        t = torch.tensor([spmv_rate], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        spmv_rate = float(t.item())

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 (u64 accumulate)", "data": "synthetic", "config": dict(config, nnz=nnz),
            "spmv_gnnzn_per_s": spmv_rate, "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline,
            "phases_ms_per_step": {k: v["ms"] / a.steps for k, v in phases.items()},
            "setup_s": {"generate": t_gen, "build_layout": t_build}, "device_bytes": info["device_bytes"]}
    try:
        line["roofline_dense"] = dense_roofline(line["phases_ms_per_step"], info["local_N1"] - info["local_N0"], info["n_pad"], peak)
    except Exception as exc:          # secondary information must never cost the bench line
        line["roofline_dense"] = {"error": str(exc)}
    if e2e:
        line["e2e"] = e2e

    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = run_cpu_reference(w, 3, 1)
        if r:
            line["cpu_baseline"] = cpu_baseline_obj(r, w, nnz)
    ctx.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
