#!/usr/bin/env python
"""bench.py -- headline benchmark: LD blocks fitted / second, genome-wide DBSLMM fit.

One "step" = one pass of the hot path (decode -> Gram -> block solve) over the whole
synthetic genome (BASELINE.json configs[2]: 22 chromosomes, ~1.1 M SNPs, 1,703 EUR LD
blocks, n_ref = 2,000, clumped large-effect SNPs).  With N GPUs the blocks are sharded by
the library's LPT cost model (no collective on the data path; rank 0 only gathers counts),
so the job size is fixed: "strong" scaling as BASELINE.json's config asks ("sharded at
1/2/4/8 GPUs"); `--scaling weak` gives every rank its own genome instead.

  value : blocks/s from the library's CUDA-event device time (bed, plan and z resident in HBM)
  e2e   : blocks/s through ONE C-ABI call per step from HOST buffers (fit_args.bed: batched panel H2D overlapped with the
          fit, plan/z H2D, kernels, beta D2H), wall clock around the synchronous calls
  --impl reference : the reference's CPU path on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (total_snps, chroms, cap, n_ref, n_obs)
    "c3": (1_100_000, list(range(1, 23)), 3000, 2000, 300_000),
    "c2": (90_000, [1], 3000, 500, 300_000),
    "tiny": (6_000, [22], 400, 400, 2400),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=list(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--missing", type=float, default=0.0)
    ap.add_argument("--solver", default="cholesky", choices=["cholesky", "pcg"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU work of the baseline sample")
    ap.add_argument("--seed", type=int, default=20240003)
    ap.add_argument("--emulate-shard", default="", help="R/N: time rank R's shard of an N-GPU run on one GPU (tuning aid)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# synthetic genome on the GPU (torch is plumbing here: RNG + pinned host buffers)
# ---------------------------------------------------------------------------------------------
def make_bed_cuda(torch, dev, n_snp, n_ref, seed, missing_rate=0.0, rho=0.9, chunk=256):
    """uint8 [n_snp, ceil(n_ref/4)] in pinned host memory.  AR(1) latent Gaussian per haplotype
    along the SNP axis (same model as dbslmm_b200.synth.make_genotypes, run as chunked scans)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    pitch = (n_ref + 3) // 4
    n4 = pitch * 4
    out = torch.empty((n_snp, pitch), dtype=torch.uint8, device=dev)
    s = (1.0 - rho * rho) ** 0.5
    i = torch.arange(chunk, device=dev, dtype=torch.float64)
    T = torch.tril(s * rho ** (i[:, None] - i[None, :]).clamp(min=0)).float()
    pw = (rho ** (i + 1.0)).float()[:, None]
    carry = torch.randn((1, 2 * n_ref), generator=g, device=dev)
    normal = torch.distributions.Normal(0.0, 1.0)
    lut = torch.tensor([3, 2, 0], dtype=torch.uint8, device=dev)       # allele count -> PLINK code
    shifts = torch.tensor([0, 2, 4, 6], dtype=torch.int32, device=dev)
    for c0 in range(0, n_snp, chunk):
        m = min(chunk, n_snp - c0)
        E = torch.randn((chunk, 2 * n_ref), generator=g, device=dev)
        L = T @ E + pw * carry
        carry = L[chunk - 1:chunk].clone()
        p = torch.rand((chunk, 1), generator=g, device=dev) * 0.45 + 0.05
        thr = normal.icdf(p)
        A = (L < thr)
        G = (A[:, :n_ref].to(torch.uint8) + A[:, n_ref:].to(torch.uint8))
        mono = (G.amin(dim=1) == G.amax(dim=1))
        if bool(mono.any()):                                             # re-draw monomorphic rows
            fix = (torch.rand((chunk, n_ref), generator=g, device=dev) < 0.3).to(torch.uint8)
            G = torch.where(mono[:, None], fix, G)
        code = lut[G.long()]
        if missing_rate > 0:
            miss = torch.rand((chunk, n_ref), generator=g, device=dev) < missing_rate
            code = torch.where(miss, torch.ones_like(code), code)
        full = torch.zeros((chunk, n4), dtype=torch.uint8, device=dev)
        full[:, :n_ref] = code
        packed = (full.view(chunk, pitch, 4).to(torch.int32) << shifts).sum(dim=2).to(torch.uint8)
        out[c0:c0 + m] = packed[:m]
    host = torch.empty((n_snp, pitch), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    host.copy_(out)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    del out
    return host


def build_workload(args, torch, dev, rank_seed):
    from dbslmm_b200 import synth
    total, chroms, cap, n_ref, n_obs = CONFIGS[args.config]
    sizes = synth.eur_block_sizes(total, cap, chroms=chroms)
    n_snp = int(sizes.sum())
    bed_t = make_bed_cuda(torch, dev, n_snp, n_ref, rank_seed, missing_rate=args.missing)
    rng = np.random.default_rng(rank_seed + 1)
    z, large = synth.make_sumstats(rng, sizes)
    s_off, s_pos, l_off, l_pos = synth.split_csr(sizes, large)
    return {"sizes": sizes, "n_snp": n_snp, "n_ref": n_ref, "n_obs": n_obs, "bed_t": bed_t,
            "bed": bed_t.numpy(), "z": z, "s_off": s_off, "s_pos": s_pos, "l_off": l_off, "l_pos": l_pos,
            "nsnp_total": n_snp}


def shard_workload(w, owner, rank, torch):
    """Compact per-rank problem (dbslmm_b200.multigpu.shard) with the .bed shard in pinned host memory."""
    from dbslmm_b200 import multigpu
    keep = []

    def pinned(shape):
        t = torch.empty(shape, dtype=torch.uint8, pin_memory=True)
        keep.append(t)
        return t.numpy()
    ww = dict(w)
    ww["s_z"] = w["z"][w["s_pos"]]
    ww["l_z"] = w["z"][w["l_pos"]]
    sh = multigpu.shard(ww, owner, rank, pinned_alloc=pinned)
    sh["_pin"] = keep
    return sh


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []          # (arrival time, text)
        self.proc = None
        self.t_mark = 0.0

    def mark(self):
        """Samples that arrive from now on count (nvidia-smi needs a moment to start: it is launched before the warm-up)."""
        self.t_mark = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if ts < self.t_mark:
                continue
            t = [x.strip() for x in ln.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1])); mx.append(float(t[2]))
            except ValueError:
                continue
            for nm, v in zip(names, t[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measure_fp64_peak(torch, dev):
    """cuBLAS DGEMM 8192^3, best of 5 (TFLOP/s): the FP64 tensor-pipe roofline denominator.
    MEASURED_PEAKS.json has no FP64 figure, so it is measured here, on this box, and labelled so."""
    n = 8192
    a = torch.randn((n, n), dtype=torch.float64, device=dev)
    b = torch.randn((n, n), dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    torch.cuda.empty_cache()
    return best


# ---------------------------------------------------------------------------------------------
# CPU arms (the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------
def cpu_sample(w, cpu_seconds, threads):
    """Every k-th block (keeps the size distribution), sized for ~cpu_seconds of host work."""
    sizes = w["sizes"].astype(np.float64)
    n = w["n_ref"]
    # calibrated on the GPU boxes' hosts: ~1.9 GFLOP/s per thread for the X'X loops, ~1 GFLOP/s for PCG mat-vecs
    est = (2.0 * n * sizes ** 2 / 1.9e9 + 60 * 2 * sizes ** 2 / 1.0e9 + 1e-4).sum() / max(threads, 1)   # seconds
    stride = max(1, int(np.ceil(est / cpu_seconds)))
    return np.arange(0, sizes.size, stride), stride


def run_cpu(w, blocks, threads):
    """CPU arm on a block sample.  Prefers the UNMODIFIED reference (oracle/_ref: IO::readSNPIm +
    SNPPROC::nomalizeVec + DBSLMMFIT::estBlock/PCG under the reference's batches-of-60 omp-dynamic
    schedule, compiled over oracle/shim) and falls back to the oracle port.  Returns (seconds, kind)."""
    from oracle import oracle as O
    from oracle import refharness as R
    s_off = np.zeros(blocks.size + 1, np.int32)
    l_off = np.zeros(blocks.size + 1, np.int32)
    sp, lp = [], []
    for i, b in enumerate(blocks):
        a = w["s_pos"][w["s_off"][b]:w["s_off"][b + 1]]
        c = w["l_pos"][w["l_off"][b]:w["l_off"][b + 1]]
        sp.append(a); lp.append(c)
        s_off[i + 1] = s_off[i] + a.size
        l_off[i + 1] = l_off[i] + c.size
    sp = np.concatenate(sp).astype(np.int64); lp = np.concatenate(lp).astype(np.int64)
    sigma_s = 0.5 / w["nsnp_total"]
    # compact .bed with only the sampled rows (the reference reads a FILE through an ifstream)
    rows = np.unique(np.concatenate([sp, lp]))
    remap = np.full(w["n_snp"], -1, np.int64); remap[rows] = np.arange(rows.size)
    sub = np.ascontiguousarray(w["bed"][rows])
    sp32, lp32 = remap[sp].astype(np.int32), remap[lp].astype(np.int32)
    if R.available():
        try:
            tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") else None
            import tempfile
            f = tempfile.NamedTemporaryFile(suffix=".bed", dir=tmpdir, delete=False); f.close()
            R.write_bed(sub, f.name)
            try:
                t = time.perf_counter()
                R.est_path(f.name, w["n_ref"], w["n_obs"], sigma_s, s_off, sp32, w["z"][sp], l_off, lp32, w["z"][lp], threads=threads)
                return time.perf_counter() - t, "reference"
            finally:
                os.unlink(f.name)
        except OSError:
            pass
    t = time.perf_counter()
    O.est(sub, w["n_ref"], w["n_obs"], sigma_s, s_off, sp32, w["z"][sp], l_off, lp32, w["z"][lp], threads=threads, mode=O.MODE_REF)
    return time.perf_counter() - t, "port"


KIND_TEXT = {"reference": "unmodified reference functions (readSNPIm + nomalizeVec + estBlock/PCG, batches of 60, omp dynamic) over oracle/shim",
             "port": "ref-mode oracle port (PCG tol 1e-7, batches of 60, omp dynamic)"}


def main():
    args = parse()
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    total, chroms, cap, n_ref, n_obs = CONFIGS[args.config]
    workload_name = {"c3": "genome-wide synthetic DBSLMM: 22 chr, ~1.1M SNPs, 1,703 EUR LD blocks, n_ref=2000, clumped large-effect SNPs",
                     "c2": "LMM-size synthetic chr1: ~90k SNPs, 133 EUR LD blocks, n_ref=500",
                     "tiny": "tiny synthetic chr22 slice"}[args.config]

    # ---------------- reference arm: CPU only, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        from dbslmm_b200 import synth
        from oracle import oracle as O
        dev = torch.device("cuda", local) if torch.cuda.is_available() else None
        if dev is None:
            print(json.dumps({"impl": "reference", "unavailable": "needs a CUDA device to synthesise the workload"}))
            return
        torch.cuda.set_device(dev)
        w = build_workload(args, torch, dev, args.seed)
        threads = min(os.cpu_count() or 1, 100)                 # reference caps -t at 100 (dbslmm.cpp:224)
        # every step is a bounded sample; the sample shrinks with the step count so the whole run stays within ~3 minutes
        per_step = max(1.5, min(args.cpu_seconds, 150.0 / (args.steps + min(args.warmup, 1))))
        blocks, stride = cpu_sample(w, per_step, threads)
        times = []
        kind = "port"
        for i in range(min(args.warmup, 1) + args.steps):       # one warm-up pass is enough on the CPU
            dt, kind = run_cpu(w, blocks, threads)
            if i >= min(args.warmup, 1):
                times.append(dt)
        dt = float(np.mean(times)) if times else float("nan")
        v = blocks.size / dt
        line = {"metric": "LD blocks fitted/sec genome-wide", "value": v, "unit": "blocks/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": {"workload": workload_name, "sample": f"every {stride}-th block ({blocks.size} of {w['sizes'].size})"},
                "cpu_baseline": {"value": v, "unit": "blocks/s", "cores": threads, "kind": kind,
                                 "sample": f"every {stride}-th LD block ({blocks.size} of {w['sizes'].size}), {KIND_TEXT[kind]}"},
                "e2e": {"value": v, "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------- B200 arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from dbslmm_b200 import _abi
    seed = args.seed if args.scaling == "strong" else args.seed + 1000 * rank
    w = build_workload(args, torch, dev, seed)
    nb_total = int(w["sizes"].size)
    eng = _abi.Engine(local)
    ms_blk = (w["s_off"][1:] - w["s_off"][:-1]).astype(np.int32)
    ml_blk = (w["l_off"][1:] - w["l_off"][:-1]).astype(np.int32)
    if args.emulate_shard:
        er, en = (int(x) for x in args.emulate_shard.split("/"))
        owner, _ = eng.plan_shards(ms_blk, ml_blk, n_ref, en)
        sh = shard_workload(w, owner, er, torch)
    elif args.scaling == "strong" and world > 1:
        owner, _ = eng.plan_shards(ms_blk, ml_blk, n_ref, world)
        sh = shard_workload(w, owner, rank, torch)
    else:
        owner = np.zeros(nb_total, np.int32) + (0 if args.scaling == "strong" else rank)
        sh = shard_workload(w, owner, owner[0], torch)
    my_blocks = int(sh["blocks"].size)
    my_snps = int(sh["s_pos"].size + sh["l_pos"].size)
    sigma_s = [0.5 / w["nsnp_total"]]
    solver = _abi.SOLVER_CHOLESKY if args.solver == "cholesky" else _abi.SOLVER_PCG
    fit_kw = dict(sigma_s=sigma_s, n_obs=n_obs, tau=0.8, solver=solver)
    csr = (sh["s_off"], sh["s_pos"], sh["s_z"], sh["l_off"], sh["l_pos"], sh["l_z"])

    fp64_peak = measure_fp64_peak(torch, dev) if rank == 0 else 0.0

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- device-resident throughput (value)
    eng.load_bed(sh["bed"], n_ref)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        r = eng.fit(*csr, **fit_kw)
    barrier()
    sampler.mark()
    dev_ms, tms = [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = eng.fit(*csr, flags=_abi.FLAG_PLAN_CACHED, **fit_kw)
        t = r["timing"]
        tms.append(t)
        dev_ms.append(t["decode_ms"] + t["gram_ms"] + t["solve_ms"])
    barrier()
    wall_resident = time.perf_counter() - t0
    clocks = sampler.stop()
    n_bad = int(r["n_bad"])

    # ---- end to end from host buffers (e2e): ONE C-ABI call per step takes the pinned host .bed shard and the CSR
    # block lists and returns the betas on the host -- the shape of the reference's DBSLMMFIT::est(bed_str, info, ...).
    # Inside the call the panel upload is cut into batches (big blocks first) and overlaps decode/Gram/Cholesky.
    for _ in range(min(args.warmup, 2)):
        eng.fit(*csr, bed=sh["bed"], n_ref=n_ref, reuse_outputs=True, **fit_kw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r2 = eng.fit(*csr, bed=sh["bed"], n_ref=n_ref, reuse_outputs=True, **fit_kw)
    barrier()
    wall_e2e = time.perf_counter() - t0
    assert np.array_equal(r2["beta_s"], r["beta_s"]) or np.abs(r2["beta_s"] - r["beta_s"]).max() <= 1e-12 * np.abs(r["beta_s"]).max()
    # the same work as two calls (upload everything, then fit): what the overlap buys
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 2)):
        eng.load_bed(sh["bed"], n_ref)
        eng.fit(*csr, **fit_kw)
    barrier()
    wall_two_call = (time.perf_counter() - t0) / max(1, args.steps // 2)

    dev_total = float(np.sum(dev_ms))
    stats = torch.tensor([dev_total, wall_e2e, wall_resident, float(my_blocks), float(my_snps)], dtype=torch.float64, device=dev)
    if dist is not None:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_total, wall_e2e, wall_resident = float(mx[0]), float(mx[1]), float(mx[2])
        blocks_all, snps_all = float(sm[3]), float(sm[4])
    else:
        blocks_all, snps_all = float(my_blocks), float(my_snps)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    K = args.steps
    value = blocks_all * K / (dev_total * 1e-3)
    e2e_v = blocks_all * K / wall_e2e
    avg = lambda k: float(np.mean([t[k] for t in tms]))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    chol_ms = avg("chol_ms")
    chol_flops = float(tms[-1]["solve_flops"])
    ach = chol_flops / (chol_ms * 1e-3) / 1e12 if chol_ms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "chol_traffic.json")))
        if world == 1 and args.config == "c3" and args.missing == 0.0:
            traffic = tj["dram_bytes_per_fit"]
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "chol_panel_kernel + chol_diag_kernel (all panel steps of one fit)",
                "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None,
                "traffic": traffic,
                "traffic_note": "dram__bytes_read+write summed over all chol_* launches of one fit (ncu, profiles/chol_traffic.json); "
                                "algorithmic flops / traffic = arithmetic intensity of the factorisation",
                "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "flops_per_step": chol_flops, "ms_per_step": chol_ms}
    dec_gbs = float(tms[-1]["decode_bytes"]) / (avg("decode_ms") * 1e-3) / 1e9 if avg("decode_ms") > 0 else 0.0
    gram_pops = float(tms[-1]["gram_ops"]) / (avg("gram_ms") * 1e-3) / 1e15 if avg("gram_ms") > 0 else 0.0
    sigma_bytes = float(np.sum((w["sizes"].astype(np.float64)) ** 2)) * 4.0     # lower triangle, 8 B
    other = {"decode": {"bound": "hbm", "achieved": dec_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": dec_gbs / hbm_peak,
                        "peak_source": hbm_src, "ms": avg("decode_ms")},
             "gram": {"bound": "hbm (Sigma write) / int8 tensor", "int8_Pops": gram_pops,
                      "sigma_write_GBs": sigma_bytes / (avg("gram_ms") * 1e-3) / 1e9 if world == 1 and avg("gram_ms") > 0 else None,
                      "ms": avg("gram_ms")},
             "solve_total_ms": avg("solve_ms"), "h2d_ms": avg("h2d_ms"), "d2h_ms": avg("d2h_ms"),
             "chol_class_ms": [float(np.mean([t["class_ms"][c] for t in tms])) for c in range(4)]}
    h2d = int(sh["bed"].nbytes + 8 * my_snps + 24 * my_snps)     # bed + z + plan rows (rank 0's share)
    d2h = int(8 * my_snps + 8 * my_blocks)
    line = {"metric": "LD blocks fitted/sec genome-wide", "value": value, "unit": "blocks/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": dev_total / K, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64 (int8 Gram, s32 accumulate)",
            "data": "synthetic", "impl": "b200",
            "config": {"workload": workload_name, "blocks": nb_total, "snps": int(w["n_snp"]), "n_ref": n_ref,
                       "n_obs": n_obs, "missing_rate": args.missing, "solver": args.solver,
                       "parallelism": f"blocks sharded over {world} GPU(s) by LPT cost model, no collective on the data path",
                       "l2": "inputs larger than L2 (codes 2.2 GB, Sigma 8+ GB per step)"},
            "snps_per_s": snps_all * K / (dev_total * 1e-3),
            "e2e": {"value": e2e_v, "unit": "blocks/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": wall_e2e / K * 1e3,
                    "how": "one dbslmm_b200_fit call per step with fit_args.bed = pinned host .bed: batched H2D of the panel "
                           "overlapped with decode/Gram/Cholesky, plan + z H2D, beta D2H; wall clock",
                    "upload_then_fit_ms_per_step": wall_two_call * 1e3},
            "resident_wall_ms_per_step": wall_resident / K * 1e3,
            "gpu_launches": int(sum(t["n_launches"] for t in tms)),
            "clocks": clocks, "roofline": roofline, "rooflines_other": other, "blocks_not_spd": n_bad}
    if not args.no_cpu_baseline and world == 1:
        try:
            threads = min(os.cpu_count() or 1, 100)
            blocks, stride = cpu_sample(w, args.cpu_seconds, threads)
            dt, kind = run_cpu(w, blocks, threads)
            line["cpu_baseline"] = {"value": blocks.size / dt, "unit": "blocks/s", "cores": threads, "kind": kind,
                                    "sample": f"every {stride}-th LD block ({blocks.size} of {nb_total}), {KIND_TEXT[kind]}, {dt:.1f} s"}
        except Exception as e:  # the checker must never take the product bench down
            line["cpu_baseline"] = {"value": None, "unit": "blocks/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
