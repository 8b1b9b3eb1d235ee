#!/usr/bin/env python
"""bench.py — block-QR TFLOP/s (Householder count 2mn^2 - 2n^3/3) on synthetic matrices.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU block QR

A "step" is one complete factorisation of the workload matrix (BASELINE.json configs[3]:
32768 x 32768, r = 128; it fits one B200, so it is the N = 1 workload too).  At N > 1 the
matrix is distributed 1-D column-block-cyclic and the SAME matrix is factored (strong scaling).
Inputs are generated on the device from a stateless hash (oracle-identical), a pristine copy
stays resident in HBM and is restored inside the timed region before each factorisation
(the algorithm is in place).  Rank 0 prints ONE JSON line.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (m, n, r)   — BASELINE.json configs
    "c4": (32768, 32768, 128),
    "c3": (4096, 16384, 64),
    "c2": (2048, 2048, 32),
    "c16k": (16384, 16384, 128),
    "c8k": (8192, 8192, 128),
    "c5": (1048576, 256, 128),   # tall-skinny: TSQR path (mpqr_tsqr_device / mpqr_mg_tsqr_device), row blocks over the GPUs
}
REF_SAMPLE = (640, 640, 32)  # bounded sample of the same workload family for the CPU reference arm


def householder_flops(m, n):
    m, n = float(m), float(n)
    return 2 * m * n * n - 2 * n ** 3 / 3 if m >= n else 2 * m * m * n - 2 * m ** 3 / 3


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tc_burst": d["bf16_tflops"], "tc_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tc_burst": 1590.0, "tc_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm
def run_reference_sample(steps, warmup):
    """Times the reference's own CPU implementation of the path (h_block_qr, Cuda/qr.cu:1275)
    from oracle/_ref when it was compiled, else the oracle port; single thread (the reference
    has no threading)."""
    import numpy as np
    import oracle
    m, n, r = REF_SAMPLE
    A = oracle.uniform_matrix(m, n, 640640)
    kind = "reference" if oracle.ref_available() else "port"
    fn = (lambda: oracle.ref_block_qr(A, r)) if kind == "reference" else (lambda: oracle.block_qr(A, r, dense=True))
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        P, Q = fn()
    dt = (time.perf_counter() - t0) / steps
    be = oracle.backward_error(A, oracle.strip_R(P), Q)
    tf = householder_flops(m, n) / dt / 1e12
    return {"value": tf, "unit": "TFLOP/s", "cores": 1, "kind": kind,
            "sample": f"{m}x{n} r={r} uniform[0,1) FP32, h_block_qr (Cuda/qr.cu:1275) incl. explicit Q, {dt:.2f} s/step",
            "backward_error": be, "seconds_per_step": dt}


def extra_cpu_baselines():
    """SURVEY 8(d) CPU lines beside the reference's own h_block_qr: (a) Eigen::HouseholderQR, the library call of the
    reference's CPU toy C++/main.cpp:54, prebuilt into oracle/_ref/eigen_qr against the vendored Eigen 3.4.0 (1 thread);
    (b) LAPACK sgeqrf through scipy on all host cores.  Bounded samples, reported as baselines only."""
    out = {"host_cores": os.cpu_count()}
    exe = os.path.join(ROOT, "oracle", "_ref", "eigen_qr")
    if os.path.exists(exe):
        try:
            m = n = 4096
            r = json.loads(subprocess.run([exe, str(m), str(n), "4096064", "1"], capture_output=True, text=True, timeout=300).stdout)
            out["eigen_householder_qr"] = {"value": householder_flops(m, n) / r["seconds"] / 1e12, "unit": "TFLOP/s", "cores": 1,
                                           "kind": "reference (Eigen::HouseholderQR<float>, C++/main.cpp:54, Eigen " + r["eigen"] + ")",
                                           "sample": f"{m}x{n} uniform[0,1) FP32, {r['seconds']:.2f} s"}
        except Exception as ex:   # noqa: BLE001
            out["eigen_householder_qr"] = {"unavailable": str(ex)[:200]}
    else:
        out["eigen_householder_qr"] = {"unavailable": "oracle/_ref/eigen_qr not built"}
    try:
        import numpy as np
        from scipy.linalg import lapack
        m = n = 8192
        rng = np.random.default_rng(8192)
        A = np.asfortranarray(rng.random((m, n), dtype=np.float32))
        lapack.sgeqrf(np.asfortranarray(A[:256, :256].copy()))          # warm the BLAS threads
        t0 = time.perf_counter()
        _, _, _, info = lapack.sgeqrf(A, overwrite_a=1)
        dt = time.perf_counter() - t0
        out["lapack_sgeqrf"] = {"value": householder_flops(m, n) / dt / 1e12, "unit": "TFLOP/s", "cores": os.cpu_count(), "kind": "LAPACK sgeqrf (scipy/OpenBLAS, all host cores)",
                                "sample": f"{m}x{n} uniform[0,1) FP32, {dt:.2f} s, info={info}"}
    except Exception as ex:   # noqa: BLE001
        out["lapack_sgeqrf"] = {"unavailable": str(ex)[:200]}
    return out


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = run_reference_sample(args.steps, args.warmup)
    m, n, r = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": "block-QR TFLOP/s (2mn^2-2n^3/3)", "value": cb["value"], "unit": "TFLOP/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"bounded sample {REF_SAMPLE[0]}x{REF_SAMPLE[1]} r={REF_SAMPLE[2]} of the {args.workload} family "
                                   f"(the reference's h_block_qr is O(m^2 n^2 / r): {m}x{n} r={r} itself is infeasible on a CPU)",
                       "timed_sample": cb["sample"], "same_config": False, "requested_workload": f"{m}x{n} r={r} block QR",
                       "host_cores": os.cpu_count()},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "backward_error": cb["backward_error"]}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- native arm
def sampled_backward_error(torch, A0, P, r, k=16):
    """||(A - QR) X||_F / (||A||_F sqrt(k)), Gaussian X, FP64, straight from the packed factor."""
    m, n = A0.shape
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(n, k, device="cuda", dtype=torch.float64, generator=g)
    AX = torch.zeros(m, k, device="cuda", dtype=torch.float64)
    Z = torch.zeros(m, k, device="cuda", dtype=torch.float64)
    anorm2 = 0.0
    step = 4096
    for i in range(0, m, step):   # chunked to bound FP64 temporaries
        blk = A0[i:i + step].double()
        AX[i:i + step] = blk @ X
        anorm2 += float((blk * blk).sum())
        Z[i:i + step] = torch.triu(P[i:min(i + step, m)].double(), diagonal=i) @ X   # (P has m + 1 rows)
    kmax = min(m, n)
    for lam in range(((kmax - 1) // r) * r, -1, -r):
        pw = min(r, kmax - lam)
        Y = torch.tril(P[lam + 1:m + 1, lam:lam + pw].double())
        Tinv = torch.triu(Y.T @ Y, 1) + 0.5 * torch.eye(pw, device="cuda", dtype=torch.float64)
        Z[lam:] -= Y @ torch.linalg.solve_triangular(Tinv, Y.T @ Z[lam:], upper=True)
    return float(torch.linalg.norm(AX - Z)) / (anorm2 ** 0.5 * k ** 0.5)


def sampled_backward_error_mg(torch, dist, A0loc, Ploc, plan, m, n, r, rank, world, k=16):
    """The same estimate for the 1-D column-block-cyclic layout: every rank holds its columns of A and of the packed
    factor; A X and R X are summed over the ranks, the panels are applied by their owners (last to first) and the
    running m x k product is handed on by a broadcast per outer block.  Outside the timed region."""
    nb = plan.nb
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(n, k, device="cuda", dtype=torch.float64, generator=g)
    dist.broadcast(X, src=0)
    nloc = plan.local_cols
    gcols = torch.tensor([(j // nb * world + rank) * nb + j % nb for j in range(nloc)], device="cuda", dtype=torch.long)
    Xl = X[gcols] if nloc else torch.zeros(0, k, device="cuda", dtype=torch.float64)
    AX = torch.zeros(m, k, device="cuda", dtype=torch.float64)
    Z = torch.zeros(m, k, device="cuda", dtype=torch.float64)
    an2 = torch.zeros(1, device="cuda", dtype=torch.float64)
    step = 4096
    for i in range(0, m, step):
        if nloc == 0:
            break
        blk = A0loc[i:i + step, :nloc].double()
        AX[i:i + step] = blk @ Xl
        an2 += (blk * blk).sum()
        rows = torch.arange(i, min(i + step, m), device="cuda").unsqueeze(1)
        Z[i:i + step] = (Ploc[i:min(i + step, m), :nloc].double() * (rows <= gcols.unsqueeze(0))) @ Xl
    dist.all_reduce(AX)
    dist.all_reduce(Z)
    dist.all_reduce(an2)
    kmax = min(m, n)
    nblk = (kmax + nb - 1) // nb
    for b in range(nblk - 1, -1, -1):
        owner = b % world
        if rank == owner:
            c0, c1 = b * nb, min((b + 1) * nb, kmax)
            l0 = (b // world) * nb
            for lam in range(c0 + ((c1 - c0 - 1) // r) * r, c0 - 1, -r):
                pw = min(r, c1 - lam)
                lc = l0 + lam - c0
                Y = torch.tril(Ploc[lam + 1:m + 1, lc:lc + pw].double())
                Tinv = torch.triu(Y.T @ Y, 1) + 0.5 * torch.eye(pw, device="cuda", dtype=torch.float64)
                Z[lam:] -= Y @ torch.linalg.solve_triangular(Tinv, Y.T @ Z[lam:], upper=True)
        dist.broadcast(Z, src=owner)
    return float(torch.linalg.norm(AX - Z)) / (float(an2.item()) ** 0.5 * k ** 0.5)


def main_native(args):
    import torch
    import mixedprecisionblockqr_b200 as pkg
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    m, n, r = WORKLOADS[args.workload]
    F = householder_flops(m, n)
    st = torch.cuda.current_stream().cuda_stream
    peaks = load_peaks()

    if world == 1:
        plan = pkg.BlockQR(m, n, r, nb=args.nb, precision=args.precision)
        nloc, lda = n, (n + 7) // 8 * 8
        A0 = torch.zeros(m, lda, device="cuda")
        pkg.fill_uniform(A0.data_ptr(), lda, n, 0, m, 0, n, args.seed, st)
    else:
        uid = [pkg.mg_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        plan = pkg.MultiGpuBlockQR(m, n, r, args.nb, rank, world, uid[0], precision=args.precision)
        nloc = plan.local_cols
        lda = (max(nloc, 8) + 7) // 8 * 8
        A0 = torch.zeros(m, lda, device="cuda")
        # every local block is generated in place from the global (row, col) hash: no transfers
        for lb in range((nloc + plan.nb - 1) // plan.nb):
            g0 = (lb * world + rank) * plan.nb
            w = min(plan.nb, n - g0)
            pkg.fill_uniform(A0.data_ptr() + 4 * lb * plan.nb, lda, n, 0, m, g0, w, args.seed, st)
    A = torch.zeros(m + 1, lda, device="cuda")

    def step():
        A[:m].copy_(A0)                        # restore the input (HBM -> HBM), inside the timed region
        plan.factor(A.data_ptr(), lda, st)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = plan.last_launches * args.steps   # kernels launched inside the timed region (counted by the library per factor call)
    # Per-kernel-class breakdown: ONE more step of the same workload with the library's per-launch
    # CUDA events switched on (they serialise the streams and cost ~2 us per launch, so they are
    # kept out of the region `value` is computed from); prof_ms is that step's own duration.
    plan.set_profiling(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record()
    step()
    p1.record()
    barrier()
    prof_ms = p0.elapsed_time(p1)
    prof = plan.profile()
    plan.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_per_step = ms / args.steps
    value = F / (ms_per_step * 1e-3) / 1e12

    # ---- accuracy of the last timed factorisation (outside the timed region)
    be = None
    if world == 1 and not args.no_check:
        be = sampled_backward_error(torch, A0[:, :n], A[:, :n], plan.r)
    elif not args.no_check:
        be = sampled_backward_error_mg(torch, dist, A0, A, plan, m, n, plan.r, rank, world)

    # ---- end to end through the host-pointer C-ABI (pinned host buffers, H2D + D2H inside)
    e2e = None
    esteps = max(1, min(args.steps, args.e2e_steps))
    if world == 1:
        host = torch.empty((m + 1, n), dtype=torch.float32, pin_memory=True)
        host[:m].copy_(A0[:, :n])
        host[m].zero_()
        src = host.clone().pin_memory()
        flags = {"fp16": pkg.MPQR_FP16, "bf16": pkg.MPQR_BF16, "fp32": pkg.MPQR_FP32}[args.precision]
        torch.cuda.synchronize()
        dt = 0.0
        for w in range(1 + esteps):   # one warm-up
            host.copy_(src)           # restore the caller's input (host memcpy): not part of the call, not timed
            t0 = time.perf_counter()
            pkg.check(pkg.lib().mpqr_block_qr_host(host.data_ptr(), None, m, n, r, flags), "mpqr_block_qr_host")
            if w >= 1:
                dt += time.perf_counter() - t0   # the call returns with the factor in `host` (device-synchronous)
        dt /= esteps
        nbytes = (m + 1) * n * 4
        e2e = {"value": F / dt / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "ms_per_step": dt * 1e3, "path": "mpqr_block_qr_host(A_host_pinned, Q=NULL): H2D + factor + D2H (plan cached by the warm-up call; first call adds ~150 ms)"}
        del host, src
    else:
        hostA = torch.empty((m + 1, lda), dtype=torch.float32, pin_memory=True)
        hostA[:m].copy_(A0)
        hostA[m].zero_()
        hostSrc = hostA.clone().pin_memory()
        barrier()
        dt = 0.0
        for w in range(1 + esteps):
            hostA.copy_(hostSrc)                      # restore the host input (host memcpy): not timed
            barrier()
            t0 = time.perf_counter()
            A.copy_(hostA, non_blocking=True)         # H2D of this rank's shard
            plan.factor(A.data_ptr(), lda, st)
            hostA.copy_(A, non_blocking=True)         # D2H of the packed factor shard
            barrier()
            if w >= 1:
                dt += time.perf_counter() - t0
        dt /= esteps
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        nbytes = (m + 1) * lda * 4
        e2e = {"value": F / dt / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world,
               "ms_per_step": dt * 1e3, "path": "per rank: pinned host shard -> H2D -> mpqr_mg_factor_device -> D2H"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline per kernel class (CUDA events recorded inside the library on the launch stream)
    # classes panel_* are the parts of "panel"; shares are taken over the leaf classes only
    leaves = [k for k in prof if k != "panel"] if any(prof[k]["launches"] for k in prof if k.startswith("panel_")) else list(prof)
    by_kernel, total_kernel_ms = {}, sum(prof[k]["ms"] for k in leaves) or 1.0
    # `traffic`: DRAM bytes of one launch from an ncu --set full capture of THIS workload (profiles/ncu_traffic.json:
    # {workload: {class: {"dram_bytes_per_launch": ..., "algorithmic_bytes_that_launch": ..., "shape": ...}}}); null when
    # no capture of the benchmarked workload exists (never a figure taken at another workload)
    traffic_file = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = {}
    if os.path.exists(traffic_file):
        try:
            traffic = json.load(open(traffic_file)).get(args.workload, {})
        except Exception:   # noqa: BLE001
            traffic = {}
    def traffic_of(name):
        """(DRAM bytes of the captured launch, its description): the capture is ONE launch of the benchmarked workload (its
        shape and its own algorithmic bytes are in `traffic_capture`), not the average launch `achieved` is computed from."""
        t = traffic.get(name)
        if not isinstance(t, dict):
            return None, None
        cap = {k: t[k] for k in ("shape", "algorithmic_bytes_that_launch", "tensor_pipe_active_pct", "tflops_that_launch", "dram_gbs_that_launch") if k in t}
        if t.get("algorithmic_bytes_that_launch"):
            cap["dram_over_algorithmic"] = t["dram_bytes_per_launch"] / t["algorithmic_bytes_that_launch"]
        return t.get("dram_bytes_per_launch"), cap

    for name in leaves:
        v = prof[name]
        if v["launches"] == 0:
            continue
        avg_ms = v["ms"] / v["launches"]
        tr, cap = traffic_of(name)
        if name in ("gemm_tn", "gemm_nn"):
            ach = v["flops"] / (v["ms"] * 1e-3) / 1e12
            by_kernel[name] = {"bound": "tensor", "achieved": ach, "peak": peaks["tc_sustained"], "unit": "TFLOP/s",
                               "frac": ach / peaks["tc_sustained"], "traffic": tr, "traffic_capture": cap,
                               "hbm_gbs_algorithmic": v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                               "hbm_frac_algorithmic": v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"]}
        else:
            ach = v["bytes"] / (v["ms"] * 1e-3) / 1e9
            by_kernel[name] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                               "frac": ach / peaks["hbm_gbs"], "traffic": tr, "traffic_capture": cap}
        by_kernel[name].update({"share_of_kernel_time": v["ms"] / total_kernel_ms, "avg_launch_ms": avg_ms,
                                "launches_per_step": v["launches"], "algorithmic_bytes_per_launch": v["bytes"] / v["launches"]})
    dominant = max(by_kernel, key=lambda k: by_kernel[k]["share_of_kernel_time"])
    roofline = dict(by_kernel[dominant])
    roofline["kernel"] = dominant
    roofline["peak_source"] = peaks["source"] + (", sustained bf16 cuBLAS" if roofline["bound"] == "tensor" else ", copy bandwidth")
    if dominant == "panel_block":
        roofline["note"] = ("latency-bound, not a bandwidth kernel: one cluster per panel runs 16 sequential reflector steps per register "
                            "block (~1.2 us each at 32768 rows); ncu: 12.5 % warps active, DRAM reads 1.02 x the block data "
                            "(profiles/r2_ncu_chain_kernel.txt); algorithmic bytes = 8*D*r per panel (SURVEY 8d)")
    # whole-QR roofline: every class at its own bound (SURVEY 8d T_roof)
    t_roof = sum(max(prof[k]["flops"] / (peaks["tc_sustained"] * 1e12) if k in ("gemm_tn", "gemm_nn") else 0.0,
                     prof[k]["bytes"] / (peaks["hbm_gbs"] * 1e9)) for k in leaves)
    whole = {"t_roof_ms": t_roof * 1e3, "frac": t_roof * 1e3 / ms_per_step,
             "frac_of_tensor_peak": value / peaks["tc_sustained"]}

    cpu_baseline, cpu_extra = None, None
    if world == 1 and not args.no_cpu_baseline:
        cb = run_reference_sample(1, 0)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu_baseline["host_cores"] = os.cpu_count()
        cpu_baseline["same_config"] = False
        cpu_extra = extra_cpu_baselines()

    line = {
        "metric": "block-QR TFLOP/s (2mn^2-2n^3/3)", "value": value, "unit": "TFLOP/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": f"{m}x{n} r={r} mixed-precision block QR ({args.workload})", "effective_r": plan.r, "outer_block_nb": plan.nb,
                   "parallelism": "single GPU" if world == 1 else f"1-D column-block-cyclic x{world}, NCCL broadcast of Y|W",
                   "seed": args.seed, "l2": "inputs (>= 4 GB at c4) are larger than L2; input restored HBM->HBM inside the timed region",
                   "flop_model": "2mn^2-2n^3/3 (m>=n) / 2m^2n-2m^3/3 (m<n)"},
        "backward_error_sampled": be, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "roofline": roofline, "roofline_by_kernel": by_kernel, "whole_qr_roofline": whole, "cpu_baseline": cpu_baseline,
        "profiled_step_ms": prof_ms, "cpu_baselines_extra": cpu_extra,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main_tsqr(args):
    """BASELINE config 5 (1048576 x 256, python/ca_qr.py TSQR path): R and the thin Q of the SAME matrix at every N,
    rows split evenly over the ranks (strong scaling), one ncclAllGather of the 256 x 256 R factors."""
    import torch
    import mixedprecisionblockqr_b200 as pkg
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    m, n, _ = WORKLOADS[args.workload]
    F = householder_flops(m, n)
    st = torch.cuda.current_stream().cuda_stream
    peaks = load_peaks()
    r0, r1 = m * rank // world, m * (rank + 1) // world
    mloc = r1 - r0
    A = torch.zeros(mloc, n, device="cuda")
    pkg.fill_uniform(A.data_ptr(), n, n, r0, mloc, 0, n, args.seed, st)
    Q = torch.zeros(mloc, n, device="cuda")
    R = torch.zeros(n, n, device="cuda")
    plan = None
    if world > 1:
        uid = [pkg.mg_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        plan = pkg.MultiGpuTSQR(rank, world, uid[0])

    def step():   # device-synchronous calls (they allocate their lanes and free them again)
        if plan is None:
            pkg.tsqr(A.data_ptr(), n, mloc, n, Q.data_ptr(), n, R.data_ptr(), n, st)
        else:
            plan.factor(A.data_ptr(), n, mloc, n, Q.data_ptr(), n, R.data_ptr(), n, st)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # accuracy (FP64 on the device): local residual and Gram pieces, reduced over the ranks
    Ad, Qd, Rd = A.double(), Q.double(), R.double()
    parts = torch.stack([((Ad - Qd @ Rd) ** 2).sum(), (Ad ** 2).sum()])
    G = Qd.T @ Qd
    # end to end: pinned host rows -> H2D -> TSQR -> D2H of the thin Q rows and R
    hostA = torch.empty((mloc, n), dtype=torch.float32, pin_memory=True)
    hostA.copy_(A)
    hostQ = torch.empty((mloc, n), dtype=torch.float32, pin_memory=True)
    hostR = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
    esteps = max(1, min(args.steps, args.e2e_steps))
    dt = 0.0
    for w in range(1 + esteps):
        barrier()
        t0 = time.perf_counter()
        A.copy_(hostA, non_blocking=True)
        step()
        hostQ.copy_(Q, non_blocking=True)
        hostR.copy_(R, non_blocking=True)
        barrier()
        if w >= 1:
            dt += time.perf_counter() - t0
    dt /= esteps
    if dist is not None:
        t = torch.tensor([ms, dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, dt = float(t[0].item()), float(t[1].item())
        dist.all_reduce(parts)
        dist.all_reduce(G)
    if rank == 0:
        ms_per_step = ms / args.steps
        value = F / (ms_per_step * 1e-3) / 1e12
        alg_bytes = 8.0 * m * n + 4.0 * n * n       # read A once, write the thin Q once (+ R): SURVEY 8d
        ach = alg_bytes / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": "block-QR TFLOP/s (2mn^2-2n^3/3)", "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"{m}x{n} tall-skinny TSQR ({args.workload}), R + thin Q",
                       "parallelism": "single GPU" if world == 1 else f"1-D row blocks x{world}, one ncclAllGather of the R factors",
                       "seed": args.seed, "l2": "input (1.07 GB) is larger than L2", "flop_model": "2mn^2-2n^3/3"},
            "backward_error": float((parts[0] / parts[1]).sqrt().item()),
            "orthogonality_fro": float((G - torch.eye(n, device="cuda", dtype=torch.float64)).norm().item()),
            "e2e": {"value": F / dt / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": 4 * m * n, "d2h_bytes_per_step": 4 * m * n + 4 * n * n,
                    "ms_per_step": dt * 1e3, "path": "per rank: pinned host rows -> H2D -> mpqr_(mg_)tsqr_device -> D2H of Q rows and R"},
            "gpu_launches": None, "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"] * world, "unit": "GB/s", "frac": ach / (peaks["hbm_gbs"] * world),
                         "traffic": None, "kernel": "whole TSQR call (per-call lane set-up included)", "peak_source": peaks["source"] + ", copy bandwidth x n_gpus"},
            "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    if plan is not None:
        plan.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--nb", type=int, default=0)
    ap.add_argument("--seed", type=int, default=32768128)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3   # timing rule: at least 3 warm-up steps
    if a.impl == "reference":
        main_reference(a)
    elif a.workload == "c5":
        main_tsqr(a)
    else:
        main_native(a)
