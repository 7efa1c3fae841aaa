#!/usr/bin/env python
"""bench.py — the hot path's headline measurement (BASELINE.json: "Ristretto MSM
Mpoints/s @1/2/4/8 B200").

A step = one multiscalar multiplication over synthetic inputs: every rank owns a
2^LG-point shard (default 2^20) of a fixed table resident in HBM as affine-Niels
plus that shard's uniform scalars, runs the whole Pippenger pipeline on it, the
ranks' 128-byte partial sums are all-gathered over NCCL and every rank adds and
encodes them.  At N=1 that is exactly the 2^20-point MSM of BASELINE.json's
config 3; at N ranks it is an N*2^20-point MSM ("scaling": "weak").

  python bench.py --gpus N --steps K --warmup W            # this framework
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference algorithm

One JSON line on stdout (rank 0).  See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import socket
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_REAL_STDOUT = None


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "ristretto255_msm_throughput"
UNIT = "Mpoints/s"
BASEPOINT = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lg-points", type=int, default=20, help="log2 of the points per GPU")
    ap.add_argument("--cpu-sample-lg", type=int, default=20, help="log2 of the CPU-baseline sample size (default: the whole per-GPU workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-combine", action="store_true", help="N>1: NCCL all-gather + combine launch instead of the fused peer exchange")
    ap.add_argument("--no-r1cs", action="store_true", help="skip the R1CS prove/verify timing at 2^16 multipliers")
    ap.add_argument("--no-varbase", action="store_true", help="skip the variable-base (ad-hoc points, no precomputed table) MSM timing")
    ap.add_argument("--r1cs-lg", type=int, default=16)
    return ap.parse_args()


# ---------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe), runs during the timed region
# ---------------------------------------------------------------------------------
class Clocks:
    Q = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL,
                text=True,
            )
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
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
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {
            "sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


ACCUM_SOURCES = ["msm_accum_kernels.cuh", "msm_sort_kernels.cuh", "ge.cuh", "fe.cuh", "msm_accum.cu"]


def kernel_source_hash() -> str:
    import hashlib

    h = hashlib.sha256()
    for name in ACCUM_SOURCES:
        with open(os.path.join(ROOT, "mpc_bulletproof_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def int32_peak():
    """wide multiply-add issue peak measured by tools/int32_peak (committed under profiles/)."""
    p = os.path.join(ROOT, "profiles", "int32_peak.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


# ---------------------------------------------------------------------------------
# reference arm: CPU restatement of the reference algorithm (oracle/c), host cores
# ---------------------------------------------------------------------------------
def ensure_oracle():
    so = os.path.join(ROOT, "oracle", "c", "libbp_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    from oracle import cbind

    return cbind


def host_threads() -> int:
    """every host core this process may run on -- NOT omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1 into its ranks, which would silently turn the CPU arm into a one-thread run"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


NSETS_ROT = 4  # scalar sets rotated through so that no step re-reads the previous step's scalars


def workload_config(args, world: int) -> dict:
    """identical in both arms: what one step computes"""
    n = 1 << args.lg_points
    return {
        "workload": f"msm_2^{args.lg_points}_points_per_gpu",
        "group": "ristretto255",
        "points_per_gpu": n,
        "points_total": world * n,
        "scalars": "uniform in [0, 2^252), 32 bytes each",
        "points": "k_i*B, k_i uniform",
        "l2": f"{NSETS_ROT} scalar sets of {n * 32 >> 20} MiB rotated per GPU; scalars + table + sort/bucket workspace are far above the 126 MB L2",
    }


def host_uniform_scalars(n: int, seed: int) -> bytes:
    import numpy as np

    g = np.random.Generator(np.random.PCG64(seed))
    a = g.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 31] &= 0x0F  # uniform in [0, 2^252) < l
    return a.tobytes()


def run_reference(args, rank, world):
    """The reference's CPU path for the same workload on the box's host cores: oracle/c (the C restatement of
    the reference algorithm; the Rust crate cannot be built here) with every host thread, one step = one
    MSM over the same number of points per GPU as this framework's arm.  Rank 0 only."""
    if rank != 0:
        return
    cbind = ensure_oracle()
    threads = host_threads()
    n = 1 << args.lg_points
    pts = cbind.basepoint_mul(host_uniform_scalars(n, 0xB2000003), threads)
    dec = cbind.DecodedPoints(pts, threads)
    sc = [host_uniform_scalars(n, 0xB2000100 + i) for i in range(NSETS_ROT)]
    for i in range(args.warmup):
        dec.msm(sc[i % NSETS_ROT], threads=threads)
    t0 = time.perf_counter()
    for i in range(args.steps):
        dec.msm(sc[i % NSETS_ROT], threads=threads)
    dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    val = n / ms / 1e3
    sample = (f"one 2^{args.lg_points}-point variable-base MSM per step, the whole per-GPU workload (dalek-style radix-2^8 "
              f"Pippenger, digit columns over {threads} OpenMP threads; points pre-decoded, scalars in host memory)")
    r1cs = None
    if not args.no_r1cs:
        r1cs = cpu_r1cs_baseline(cbind, threads, args.r1cs_lg)
        r1cs["config1_shuffle_k64"] = cpu_shuffle_baseline(cbind, threads)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": val,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u64 limbs (radix 2^51)",
        "data": "synthetic",
        "config": workload_config(args, max(args.gpus, 1)),
        "impl_config": {"algorithm": "variable-base Pippenger, radix 2^8 (32 bucket additions per point), no precomputed table",
                        "host_threads": threads, "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "r1cs": r1cs,
        "gpu_launches": 0,
    }
    emit(line)


def cpu_r1cs_baseline(cbind, threads: int, lg: int) -> dict:
    """The first half of BASELINE.json's metric on the host cores: the group work of `Prover::prove` and
    `Verifier::verify` at 2^lg multipliers as the reference performs it (src/r1cs/prover.rs:465-494, 627-631,
    699-708; src/r1cs/verifier.rs:516-547), from oracle/c: the three phase-1 commitment MSMs (2n+1, n+1, 2n+1
    terms), `InnerProductProof::create` at n = 2^lg with factors (2n sequential scalar multiplications, then
    lg n rounds of two MSMs and a threaded two-term-MSM fold), and for verify the single 2n+13+2 lg n term MSM.
    The O(n) scalar-field work around them (flattening, polynomials) is NOT included: this is a lower bound
    of the CPU time, which makes the GPU/CPU ratio conservative."""
    n = 1 << lg
    G = cbind.basepoint_mul(host_uniform_scalars(n, 0xB2000011), threads)
    H = cbind.basepoint_mul(host_uniform_scalars(n, 0xB2000012), threads)
    Q = cbind.basepoint_mul(host_uniform_scalars(1, 0xB2000013), threads)
    s = [host_uniform_scalars(2 * n + 16 + 2 * lg, 0xB2000020 + i) for i in range(4)]
    dec = cbind.DecodedPoints(G + H, threads)
    t0 = time.perf_counter()
    dec.msm(s[0][: 64 * n], 0, 2 * n, threads=threads)  # A_I
    dec.msm(s[1][: 32 * n], 0, n, threads=threads)      # A_O
    dec.msm(s[2][: 64 * n], 0, 2 * n, threads=threads)  # S
    t_commit = time.perf_counter() - t0
    us = host_uniform_scalars(lg, 0xB2000030)
    t0 = time.perf_counter()
    cbind.ipp_create(Q, s[0][: 32 * n], s[1][: 32 * n], G, H, s[2][: 32 * n], s[3][: 32 * n], us, threads)
    t_ipp = time.perf_counter() - t0
    t0 = time.perf_counter()
    dec.msm(s[3][: 64 * n], 0, 2 * n, threads=threads)  # the verifier's mega-MSM (its 13 + 2 lg n ad-hoc terms are negligible)
    t_verify = time.perf_counter() - t0
    return {
        "multipliers": n,
        "prove_ms": (t_commit + t_ipp) * 1e3,
        "verify_ms": t_verify * 1e3,
        "prove_parts_ms": {"commitment_msms": t_commit * 1e3, "ipp_create": t_ipp * 1e3},
        "cores": threads,
        "kind": "port",
        "what": "group work only (commitment MSMs + InnerProductProof::create; verify: the mega-MSM), points pre-decoded for the MSMs; a lower bound of the CPU prove/verify time",
    }


# ---------------------------------------------------------------------------------
# this framework
# ---------------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from mpc_bulletproof_b200 import Comb, Context, Table
    from mpc_bulletproof_b200.api import dev_sum_encode

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    n = 1 << args.lg_points

    def uniform_scalars(count, seed):
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        s = torch.randint(-(2**31), 2**31, (count, 8), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
        s[:, 7] &= 0x0FFFFFFF  # uniform in [0, 2^252) < l
        return s.contiguous()

    with torch.cuda.stream(stream):
        # synthetic table: point i of this rank = k_i * B, k_i uniform (product fixed-base path)
        comb = Comb(ctx, BASEPOINT)
        gen_k = uniform_scalars(n, 0xB2000003 + 1000 * rank)
        pts_bytes = torch.empty(n * 32, dtype=torch.uint8, device=dev)
        comb.dev_mul(gen_k.data_ptr(), n, pts_bytes.data_ptr())
        # resident generator-style table: window multiples precomputed once at upload
        table = Table(ctx, dev_ptr=pts_bytes.data_ptr(), n=n).set_windows(0)
        table_windows = (255 + table.window - 1) // table.window
        scal = [uniform_scalars(n, 0xB2000100 + 1000 * rank + i) for i in range(NSETS_ROT)]
        part = torch.zeros(32, dtype=torch.int32, device=dev)  # this rank's partial sum (extended point)
        parts = torch.zeros(world * 32, dtype=torch.int32, device=dev)
        result = torch.zeros(32, dtype=torch.uint8, device=dev)
    stream.synchronize()

    # N>1: the ranks' 128-byte partial sums meet in peer-mapped buffers; ONE kernel per rank pushes,
    # waits, adds and encodes (mpc_bulletproof_b200.multi.PeerExchange).  --nccl-combine keeps the
    # all-gather + combine-launch form for comparison.
    peer = None
    if world > 1 and not args.nccl_combine:
        from mpc_bulletproof_b200.multi import PeerExchange

        peer = PeerExchange(ctx, max_sets=1)
    combine = "single GPU" if world == 1 else ("NCCL all_gather + combine kernel" if peer is None else "fused peer exchange (P2P stores + flags, one kernel)")

    def step(i):
        table.dev_msm(scal[i % NSETS_ROT].data_ptr(), 1, part.data_ptr())
        if peer is not None:
            peer.exchange_sum_encode(part.data_ptr(), 1, result.data_ptr())
        elif world > 1:
            dist.all_gather_into_tensor(parts, part)
            dev_sum_encode(ctx, parts.data_ptr(), world, 1, result.data_ptr())
        else:
            dev_sum_encode(ctx, part.data_ptr(), 1, 1, result.data_ptr())

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput ("value") ----------------------------------
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step(i)
    barrier()
    ctx.profile(True)
    ctx.profile_reset()
    launches0 = ctx.launches
    clocks = Clocks(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(args.steps):
            step(i)
        e1.record(stream)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launches - launches0
    prof = ctx.profile_read()
    ctx.profile(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * n / ms_step / 1e3
    result_hex = bytes(result.cpu().tolist()).hex()

    # ---- in-run parity of the timed computation, every N: the table's points are k_i*B, so the result of the
    # last timed step must be (sum over all ranks of sum_i s_i k_i mod l) * B -- formed here from each rank's
    # exact integer dot product (host big integers) and ONE fixed-base multiplication
    L_ORDER = 2**252 + 27742317777372353535851937790883648493
    last = scal[(args.steps - 1) % NSETS_ROT] if args.steps > 0 else scal[0]
    if args.steps == 0:
        with torch.cuda.stream(stream):
            step(0)
        barrier()
        result_hex = bytes(result.cpu().tolist()).hex()

    def ints(t):
        raw = t.cpu().numpy().tobytes()
        return [int.from_bytes(raw[i : i + 32], "little") for i in range(0, len(raw), 32)]

    gen_k_ints = ints(gen_k)

    def expected_for(scalars_t):
        dot = sum(a * k for a, k in zip(ints(scalars_t), gen_k_ints)) % L_ORDER
        if world > 1:
            mine = torch.tensor(list(dot.to_bytes(32, "little")), dtype=torch.uint8, device=dev)
            all_dots = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(all_dots, mine)
            dot = sum(int.from_bytes(bytes(d.cpu().tolist()), "little") for d in all_dots) % L_ORDER
        return comb.mul(dot.to_bytes(32, "little")).hex()

    expected_hex = expected_for(last)
    result_ok = expected_hex == result_hex

    # ---- end to end through the host-buffer C ABI ("e2e") ----------------------
    # every step: its scalars start in page-locked HOST memory and go host -> device, the MSM runs, the 32-byte
    # result comes back device -> host; all of it inside the timed region.  The steps are PIPELINED two deep, as
    # a caller with a stream of MSMs would run them: N = 1 through bpg_msm_table_submit / bpg_msm_job_wait (the
    # library's own double-buffered staging), N > 1 with the same two-slot scheme around bpg_dev_msm_table and
    # the exchange.  `e2e_sync` is the one-call-at-a-time form (bpg_msm_table), `e2e_pageable` the pipelined
    # form with ordinary (pageable) host memory, which is what a plain Vec<u8> caller passes.
    import ctypes

    from mpc_bulletproof_b200._lib import check, lib

    host_sc = [s.cpu().pin_memory() for s in scal[:2]]
    host_out = [torch.empty(32, dtype=torch.uint8).pin_memory() for _ in range(2)]
    out_buf = ctypes.create_string_buffer(32)

    def e2e_sync_step(i):
        check(lib().bpg_msm_table(ctx._h, table._h, 0, n, ctypes.c_void_p(host_sc[i % 2].data_ptr()), 1, out_buf))

    if world == 1:
        jobs = {}

        def submit(i, src=None):
            src = host_sc if src is None else src
            j = ctypes.c_void_p()
            check(lib().bpg_msm_table_submit(ctx._h, table._h, 0, n, ctypes.c_void_p(src[i % 2].data_ptr()), 1, ctypes.byref(j)))
            jobs[i] = j

        def wait(i):
            check(lib().bpg_msm_job_wait(jobs.pop(i), out_buf))
    else:
        copy_stream = torch.cuda.Stream(device=dev)
        d_in = [torch.empty_like(scal[0]) for _ in range(2)]
        res2 = [torch.zeros(32, dtype=torch.uint8, device=dev) for _ in range(2)]
        ev_up = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]

        def submit(i, src=None):
            src = host_sc if src is None else src
            k = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_done[k])  # slot k was last read two submissions ago
                d_in[k].copy_(src[k], non_blocking=True)
                ev_up[k].record(copy_stream)
            with torch.cuda.stream(stream):
                stream.wait_event(ev_up[k])
                table.dev_msm(d_in[k].data_ptr(), 1, part.data_ptr())
                if peer is not None:
                    peer.exchange_sum_encode(part.data_ptr(), 1, res2[k].data_ptr())
                else:
                    dist.all_gather_into_tensor(parts, part)
                    dev_sum_encode(ctx, parts.data_ptr(), world, 1, res2[k].data_ptr())
                host_out[k].copy_(res2[k], non_blocking=True)
                ev_done[k].record(stream)

        def wait(i):
            ev_done[i % 2].synchronize()

    def pipelined(steps, src=None):
        submit(0, src)
        for i in range(1, steps):
            submit(i, src)
            wait(i - 1)
        wait(steps - 1)

    def timed(fn, steps):
        barrier()
        t0 = time.perf_counter()
        fn(steps)
        barrier()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    pipelined(max(2, min(3, args.warmup)))
    e2e_ms = timed(pipelined, max(args.steps, 2))
    e2e_val = world * n / e2e_ms / 1e3
    e2e_expected_hex = expected_for(scal[(max(args.steps, 2) - 1) % 2])
    e2e_last = bytes(out_buf.raw).hex() if world == 1 else bytes(host_out[(max(args.steps, 2) - 1) % 2].tolist()).hex()
    e2e_sync_ms = e2e_page_ms = None
    if world == 1:
        for i in range(2):
            e2e_sync_step(i)
        e2e_sync_ms = timed(lambda k: [e2e_sync_step(i) for i in range(k)], max(args.steps, 2))
        page_sc = [torch.from_numpy(s.numpy().copy()) for s in host_sc]  # ordinary pageable memory
        pipelined(2, page_sc)
        e2e_page_ms = timed(lambda k: pipelined(k, page_sc), max(args.steps, 2))

    # ---- rank 0: roofline, CPU baseline, JSON ----------------------------------
    if rank == 0:
        peaks, peaks_kind = measured_peaks()
        acc_ms, acc_n = prof["accum"]
        acc_ms = acc_ms / max(acc_n, 1)
        alg_bytes = 128.0 * n  # 32 B scalar + 96 B Niels entry per point
        achieved = alg_bytes / (acc_ms * 1e-3) / 1e9 if acc_ms else None
        # DRAM bytes per launch of the dominant kernel come from an ncu capture (profiles/accum_traffic.json);
        # the capture records a hash of the kernel's sources, and a capture of other code reads as null
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "accum_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            if tj.get("kernel_source_sha256") == kernel_source_hash():
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = tj.get("capture")
        roof = {
            "bound": "hbm",
            "kernel": "k_accum (bucket accumulation)",
            "achieved": achieved,
            "peak": peaks["hbm_gbs"],
            "peak_kind": f"{peaks_kind} (MEASURED_PEAKS.json hbm_gbs)" if peaks_kind == "measured" else "fallback",
            "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"] if achieved else None,
            "traffic": traffic,
            "traffic_capture": traffic_src,
            "kernel_ms": acc_ms,
            "kernel_share_of_step": acc_ms / ms_step if ms_step else None,
            "note": "the kernel is INT32-issue-bound, not HBM-bound; see int32",
        }
        ip = int32_peak()
        if ip and acc_ms:
            wide = 16 * 504.0 * n  # 16 mixed adds per point x 7 fe_mul x 72 wide MADs
            roof["int32"] = {
                "achieved_wide_mad_per_s": wide / (ms_step * 1e-3),
                # what k_accum really executes: one mixed addition per (point, window) of the table
                "kernel_wide_mad_per_s": (n * table_windows * 504.0) / (acc_ms * 1e-3),
                "kernel_frac_of_peak": (n * table_windows * 504.0) / (acc_ms * 1e-3) / (ip["imad_wide_Tops"] * 1e12),
                "kernel_frac_of_chained_peak": (n * table_windows * 504.0) / (acc_ms * 1e-3) / (ip["imad_wide_x_Tops"] * 1e12),
                "peak_wide_mad_per_s": ip["imad_wide_Tops"] * 1e12,
                "frac_of_step": wide / (ms_step * 1e-3) / (ip["imad_wide_Tops"] * 1e12),
                "peak_kind": "measured (tools/int32_peak, profiles/int32_peak.json)",
            }
        phases = {k: v[0] / args.steps for k, v in prof.items() if v[1]}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(args, pts_bytes, scal[0], n)
        r1cs = None
        if world == 1 and not args.no_r1cs:
            r1cs = r1cs_timing(ctx, comb, args.r1cs_lg, dev)
        varbase = None
        if world == 1 and not args.no_varbase:
            varbase = variable_base_timing(ctx, pts_bytes, host_sc[0], n, min(args.steps, 5))
        table_bytes = n * table.entry_bytes * table_windows
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u32 limbs (8x32-bit, GF(2^255-19))",
            "data": "synthetic",
            "config": workload_config(args, world),
            "impl_config": {
                "algorithm": f"fixed-base Pippenger over a resident windowed table (one gathered multiple per window, {table_windows} bucket additions per point, no doublings)",
                "table": f"windowed affine-Niels, c={table.window}, {table_bytes >> 20} MiB resident",
                "points_from": "device fixed-base comb",
                "gpoint_ops_per_s_eq": 16 * world * n / (ms_step * 1e-3) / 1e9,
                "combine": combine,
            },
            "e2e": {
                "value": e2e_val,
                "unit": UNIT,
                "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": n * 32,
                "d2h_bytes_per_step": 32,
                "api": "bpg_msm_table_submit / bpg_msm_job_wait (page-locked host scalars, two jobs in flight)" if world == 1
                else f"two-slot pipeline: pinned H2D on a copy stream + bpg_dev_msm_table + {combine} + D2H",
                "result_ok": e2e_last == e2e_expected_hex,
                "sync_ms_per_step": e2e_sync_ms,
                "sync_value": (n / e2e_sync_ms / 1e3) if e2e_sync_ms else None,
                "pageable_ms_per_step": e2e_page_ms,
                "pageable_value": (n / e2e_page_ms / 1e3) if e2e_page_ms else None,
            },
            "gpu_launches": int(launches),
            "phases_ms": phases,
            "roofline": roof,
            "cpu_baseline": cpu,
            "r1cs": r1cs,
            "variable_base": varbase,
            "clocks": clk,
            "result": result_hex,
            "result_ok": result_ok and e2e_last == e2e_expected_hex,
            "result_check": "encode((sum_ranks sum_i s_i k_i mod l) * B) from host big integers + one fixed-base multiplication",
        }
        emit(line)
    if peer is not None:
        if not peer.ok():
            print("peer exchange timed out on rank", rank, file=sys.stderr)
        barrier()
        peer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not (result_ok and e2e_last == e2e_expected_hex):
        print(f"rank {rank}: RESULT MISMATCH: timed {result_hex} / e2e {e2e_last} vs expected {expected_hex} / {e2e_expected_hex}", file=sys.stderr)
        sys.exit(3)


def r1cs_timing(ctx, comb, lg, dev):
    """BASELINE.json's other half: R1CS prove / verify wall-clock at 2^lg multipliers through the
    host mirror's C ABI (bpg_prover_prove / bpg_verifier_verify), on the reference's own benchmark
    circuit (benches/r1cs.rs:24-32: a chain of squarings).  Host buffers in, proof bytes out."""
    import torch

    from mpc_bulletproof_b200 import protocol as P

    n = 1 << lg

    def synth(count, seed):
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        k = torch.randint(-(2**31), 2**31, (count, 8), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
        k[:, 7] &= 0x0FFFFFFF
        out = torch.empty(count * 32, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)
        comb.dev_mul(k.data_ptr(), count, out.data_ptr())
        ctx.sync()
        return bytes(out.cpu().numpy().tobytes())

    gens = P.Gens(ctx, synth(n, 11), synth(n, 12), BASEPOINT, synth(1, 13))
    val = 0x1234567890ABCDEF1234567890ABCDEF
    pm, vm = [], []
    proof = None
    for it in range(6):
        p = P.Prover(gens, P.Transcript(b"bench r1cs"))
        p.square_chain(p.commit_public(val), n)
        t0 = time.perf_counter()
        proof = p.prove(1234 + it)
        pm.append((time.perf_counter() - t0) * 1e3)
        v = P.Verifier(gens, P.Transcript(b"bench r1cs"))
        v.square_chain(v.commit_public(val), n)
        t0 = time.perf_counter()
        v.verify(proof)
        vm.append((time.perf_counter() - t0) * 1e3)
    v = P.Verifier(gens, P.Transcript(b"bench r1cs"))
    v.square_chain(v.commit_public(val + 1), n)
    rejected = False
    try:
        v.verify(proof)
    except P.VerificationError:
        rejected = True
    gens.close()
    pm, vm = sorted(pm[1:]), sorted(vm[1:])  # first run warms the pools
    shuffle = shuffle_timing(ctx, synth)
    return {
        "config1_shuffle_k64": shuffle,
        "multipliers": n,
        "circuit": "chain of squarings (reference benches/r1cs.rs:24-32)",
        "prove_ms": pm[len(pm) // 2],
        "verify_ms": vm[len(vm) // 2],
        "prove_ms_min": pm[0],
        "verify_ms_min": vm[0],
        "proof_bytes": len(proof),
        "false_statement_rejected": rejected,
        "timed": "bpg_prover_prove / bpg_verifier_verify wall-clock, median of 5 after one warm-up",
    }


def variable_base_timing(ctx, pts_bytes_dev, host_scalars, n, reps):
    """BASELINE.json config 3 as the reference calls it for points it has never seen (src/r1cs/verifier.rs:516-547
    for ad-hoc points): `bpg_msm` on n COMPRESSED points and n scalars, both in host memory -- upload, decode to
    affine Niels, a plain (one multiple per point) table, Pippenger with per-window bucket arrays and Horner;
    nothing precomputed, nothing resident.  Wall clock per call, best of `reps`."""
    import ctypes

    from mpc_bulletproof_b200._lib import check, lib

    pts = pts_bytes_dev.cpu().pin_memory()
    out = ctypes.create_string_buffer(32)
    best, res = 1e9, None
    for _ in range(max(reps, 2)):
        t0 = time.perf_counter()
        check(lib().bpg_msm(ctx._h, ctypes.c_void_p(host_scalars.data_ptr()), ctypes.c_void_p(pts.data_ptr()), n, out))
        best = min(best, time.perf_counter() - t0)
        res = out.raw.hex()
    return {
        "api": "bpg_msm (compressed points + scalars in host memory; decode + plain table + Pippenger with Horner per call)",
        "points": n,
        "ms": best * 1e3,
        "value": n / best / 1e6,
        "unit": UNIT,
        "h2d_bytes": n * 64,
        "result": res,
        "note": "the headline `value` uses a table precomputed once (fixed-base, 13 additions per point); this is the no-precomputation form, the same algorithm class as the CPU arm",
    }


def shuffle_timing(ctx, synth):
    """BASELINE.json config 1: the reference's shuffle bench (benches/shuffle.rs:147-236) at k = 64 through the
    product's gadget module -- 128 committed values, two phases (n1 = 0, n2 = 126 multipliers, padded to 128)."""
    import random

    from mpc_bulletproof_b200 import gadgets as PG
    from mpc_bulletproof_b200 import protocol as P

    k = 64
    gens = P.Gens(ctx, synth(128, 21), synth(128, 22), BASEPOINT, synth(1, 23))
    r = random.Random(64)
    inp = [r.randrange(2**64) for _ in range(k)]
    outp = inp[:]
    r.shuffle(outp)
    blinds = [r.randrange(P.L) for _ in range(2 * k)]
    pm, vm = [], []
    for it in range(6):
        t0 = time.perf_counter()
        proof, ic, oc = PG.shuffle_prove(gens, P.Transcript, b"ShuffleBenchmark", inp, outp, blinds, rng_seed=100 + it)
        pm.append((time.perf_counter() - t0) * 1e3)
        t0 = time.perf_counter()
        PG.shuffle_verify(gens, P.Transcript, b"ShuffleBenchmark", proof, ic, oc)
        vm.append((time.perf_counter() - t0) * 1e3)
    gens.close()
    pm, vm = sorted(pm[1:]), sorted(vm[1:])
    return {"k": k, "prove_ms": pm[len(pm) // 2], "verify_ms": vm[len(vm) // 2], "proof_bytes": len(proof),
            "timed": "constraint-system construction (128 commitments, gadget) + prove / verify through the Python face of the C ABI"}


def cpu_shuffle_baseline(cbind, threads: int) -> dict:
    """config 1 on the host cores, group work only (as cpu_r1cs_baseline): 128 Pedersen commitments, the phase-2
    commitment MSMs (253, 127, 253 terms), five T commitments and InnerProductProof::create at n = 128; verify:
    one MSM of 2 * 128 + 128 + 13 + 14 terms."""
    n, k = 128, 64
    G = cbind.basepoint_mul(host_uniform_scalars(n, 0xB2000041), threads)
    H = cbind.basepoint_mul(host_uniform_scalars(n, 0xB2000042), threads)
    Q = cbind.basepoint_mul(host_uniform_scalars(1, 0xB2000043), threads)
    s = [host_uniform_scalars(512, 0xB2000050 + i) for i in range(4)]
    t0 = time.perf_counter()
    for i in range(2 * k + 5):  # V_j and T_i: two-term commitments
        cbind.msm(s[0][64 * (i % 8) : 64 * (i % 8) + 64], G[:64], 1)
    cbind.msm(s[0][: 32 * 253], (G + H)[: 32 * 253], threads)
    cbind.msm(s[1][: 32 * 127], G[: 32 * 127], threads)
    cbind.msm(s[2][: 32 * 253], (G + H)[: 32 * 253], threads)
    t_commit = time.perf_counter() - t0
    us = host_uniform_scalars(7, 0xB2000060)
    t0 = time.perf_counter()
    cbind.ipp_create(Q, s[0][: 32 * n], s[1][: 32 * n], G, H, s[2][: 32 * n], s[3][: 32 * n], us, threads)
    t_ipp = time.perf_counter() - t0
    t0 = time.perf_counter()
    cbind.msm((s[0] + s[1])[: 32 * 411], ((G + H) * 2)[: 32 * 411], threads)
    t_verify = time.perf_counter() - t0
    return {"k": k, "prove_ms": (t_commit + t_ipp) * 1e3, "verify_ms": t_verify * 1e3, "cores": threads, "kind": "port",
            "what": "group work only (points decoded inside the calls)"}


def cpu_baseline(args, pts_bytes_dev, scal_dev, n):
    """The C restatement of the reference CPU algorithm on a bounded sample, host cores."""
    try:
        cbind = ensure_oracle()
    except Exception as e:  # build tools missing on the box
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}
    m = min(n, 1 << args.cpu_sample_lg)
    pts = bytes(pts_bytes_dev[: m * 32].cpu().numpy().tobytes())
    sc = bytes(scal_dev[:m].cpu().numpy().tobytes())
    threads = host_threads()
    dec = cbind.DecodedPoints(pts, threads)
    dec.msm(sc, threads=threads)
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        out = dec.msm(sc, threads=threads)
    dt = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    out1 = dec.msm(sc[: 32 * (m // 4)], 0, m // 4, threads=1)
    dt1 = time.perf_counter() - t0
    # parity of the sample against the GPU through the C ABI
    from mpc_bulletproof_b200 import Context, Table

    ctx2 = Context(0)
    gpu = Table(ctx2, pts).msm(sc)[0]
    ctx2.close()
    ipp = cpu_ipp_baseline(cbind, threads, pts, sc, 14)
    r1cs_cpu = None if args.no_r1cs else cpu_r1cs_baseline(cbind, threads, args.r1cs_lg)
    if r1cs_cpu is not None:
        r1cs_cpu["config1_shuffle_k64"] = cpu_shuffle_baseline(cbind, threads)
    return {
        "value": m / dt / 1e6,
        "unit": UNIT,
        "cores": threads,
        "kind": "port",
        "ipp_create": ipp,
        "r1cs": r1cs_cpu,
        "sample": f"first 2^{m.bit_length() - 1} points+scalars of the workload, dalek-style radix-2^8 Pippenger, points pre-decoded, {reps} runs",
        "single_thread_value": (m // 4) / dt1 / 1e6,
        "bytes_equal_gpu": gpu == out,
    }


def cpu_ipp_baseline(cbind, threads, pts, sc, lg):
    """BASELINE.json config 2 beside the GPU numbers: the C restatement of the reference's CPU
    `InnerProductProof::create` (src/inner_product_proof.rs:49-193: factor multiplications one after
    the other, two-term MSM folds on all threads above the threshold) at n = 2^lg, and the product's
    `bpg_ipp_create` on the same inputs; the final a, b must agree byte for byte (the challenges fed
    to the CPU run are the ones the product's transcript produced)."""
    import ctypes

    from mpc_bulletproof_b200 import Context, Table
    from mpc_bulletproof_b200 import protocol as P

    n = 1 << lg
    Gb, Hb, Q = pts[: 32 * n], pts[32 * n : 64 * n], pts[64 * n : 64 * n + 32]
    a, b, Gf, Hf = sc[: 32 * n], sc[32 * n : 64 * n], sc[64 * n : 96 * n], sc[96 * n : 128 * n]
    ctx = Context(0)
    tab = Table(ctx, Gb + Hb).set_windows(0)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        proof = P.InnerProductProof.create(ctx, P.Transcript(b"bench ipp"), Q, Gf, Hf, tab, tab, a, b, h_off=n)
        best = min(best, time.perf_counter() - t0)
    # replay the transcript on the host mirror to recover the challenges u_j
    tr = P.Transcript(b"bench ipp")
    tr.append_message(b"dom-sep", b"ipp v1")
    tr.append_u64(b"n", n)
    us = b""
    for L, R in zip(proof.L_vec, proof.R_vec):
        tr.append_message(b"L", L)
        tr.append_message(b"R", R)
        us += P.sc_bytes(tr.challenge_scalar(b"u"))
    t0 = time.perf_counter()
    lr, fa, fb = cbind.ipp_create(Q, Gf, Hf, Gb, Hb, a, b, us, threads)
    cpu_s = time.perf_counter() - t0
    same = [x for pair in lr for x in pair] == [x for pair in zip(proof.L_vec, proof.R_vec) for x in pair]
    same = same and fa == P.sc_bytes(proof.a) and fb == P.sc_bytes(proof.b)
    tab.close()
    ctx.close()
    return {"n": n, "cpu_ms": cpu_s * 1e3, "cores": threads, "gpu_ms": best * 1e3, "kind": "port",
            "bytes_equal_gpu": same, "note": "points decoded inside the CPU run; GPU generators resident in one windowed table"}


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl == "b200":
        # launched bare: re-launch under torchrun, one rank per GPU (the ranks inherit this stdout)
        cmd = [
            sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
            "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.abspath(__file__),
        ] + sys.argv[1:]
        sys.exit(subprocess.call(cmd, stdout=sys.stdout.fileno()))
    # Only the JSON line may reach stdout (NCCL and friends print banners there): route fd 1 to
    # stderr for the whole run and keep the real stdout for the final line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
