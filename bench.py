#!/usr/bin/env python
"""bench.py — headline measurement of the zkb200 hot path (see DESIGN.md "Measurement").

Step   : one 2^24-point BN254 G1 MSM (ParamsKZG::commit shape: uniform Fr scalars against an SRS resident in
         HBM) per GPU.  At N GPUs the MSM is sharded by SRS point range, one 2^24-point shard per rank (weak
         scaling; at N=4 this is the 2^26-point MSM of BASELINE.json's sweep), the 96-byte partial results are
         folded on the host.
value  : points/s, whole job, scalars already in HBM, timed with CUDA events on the launch stream.
e2e    : the same through the host-buffer C ABI call (zkb_msm_g1_srs): pinned host scalars -> H2D -> kernels ->
         window sums D2H -> host fold, wall clock (the host fold is part of the call).
ntt    : secondary object: batched 2^22 Fr NTT (16 columns) elements/s and its HBM roofline.
quotient: secondary object: evaluate_h's custom-gate pass (halo2-base gate on 4 advice columns, 2^24 extended rows resident in HBM)
         rows/s and its HBM roofline.  At N > 1 `sharded_quotient` is the same pass sharded by rows (halo exchange over NCCL) and
         `sharded_ntt` one 2^26 NTT sharded over the ranks.
`--impl reference` times the CPU restatement of halo2's best_multiexp (oracle/, all host threads) on a bounded
sample of the same workload; the reference itself is Rust with un-vendored dependencies and cannot be built
in this image (DESIGN.md "Oracle").
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LOG_N_MSM = int(os.environ.get("ZKB_BENCH_LOG_N", "24"))
NTT_LOG_N = int(os.environ.get("ZKB_BENCH_NTT_LOG_N", "22"))
NTT_COLS = int(os.environ.get("ZKB_BENCH_NTT_COLS", "16"))
QUOT_LOG_N = int(os.environ.get("ZKB_BENCH_QUOT_LOG_N", "24"))   # extended domain of the wrapper circuit (k = 22, extended_k = 24)
QUOT_COLS = int(os.environ.get("ZKB_BENCH_QUOT_COLS", "4"))
SHARDED_LOG_N = int(os.environ.get("ZKB_BENCH_SHARDED_LOG_N", "26"))
CPU_SAMPLE_LOG_N = int(os.environ.get("ZKB_BENCH_CPU_LOG_N", "24"))   # cpu_baseline leg: the full workload once, ~10 s on 16 threads
REF_SAMPLE_LOG_N = int(os.environ.get("ZKB_BENCH_REF_LOG_N", "22"))   # --impl reference: bounded sample per step
METRIC = "BN254 G1 MSM throughput (2^%d points per GPU, SRS resident)" % LOG_N_MSM


def random_field(n, seed):
    from util import random_field as rf

    return rf(n, seed)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe).  Started before the warm-up (nvidia-smi needs a
    moment to come up), rows are time-stamped and only those inside the timed window(s) are used."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.windows = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(pred):
            sm, mx, reasons = [], [], set()
            for t, r in self.rows:
                if len(r) < 9 or not pred(t):
                    continue
                try:
                    sm.append(float(r[2])); mx.append(float(r[3]))
                except ValueError:
                    continue
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons

        # a row printed at wall time t describes the ~50 ms before it
        sm, mx, reasons = collect(lambda t: any(t0 <= t <= t1 + 0.1 for t0, t1 in self.windows))
        scope = "timed windows"
        if not sm:
            sm, mx, reasons = collect(lambda t: True)
            scope = "whole run (no sample fell inside the timed windows)"
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "scope": scope, "reasons": sorted(reasons)}


def cpu_baseline_msm(log_n: int, repeats: int = 1):
    """Oracle restatement of best_multiexp on all host threads; returns (pts/s, cores, seconds)."""
    from oracle import coracle

    coracle.build()
    n = 1 << log_n
    s = random_field(n, 0x5EED0000 + log_n)
    # bases: cheap synthetic curve points for the CPU arm — multiples of G by small random scalars (CPU fixed-base)
    b = coracle.g1_fixed_base_mul(random_field(min(n, 4096), 7))
    bases = np.ascontiguousarray(np.tile(b, (n // b.shape[0] + 1, 1))[:n])
    cores = coracle.num_threads()
    best = None
    for _ in range(repeats):
        t = time.perf_counter()
        coracle.best_multiexp(s, bases, 0)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return n / best, cores, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ts = []
    cores = 0
    for i in range(args.warmup + args.steps):
        v, cores, dt = cpu_baseline_msm(REF_SAMPLE_LOG_N)
        if i >= args.warmup:
            ts.append(dt)
    n = 1 << REF_SAMPLE_LOG_N
    value = n * len(ts) / sum(ts)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pts/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(ts) / len(ts),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64x4 (254-bit Montgomery)",
        "data": "synthetic",
        "config": {"workload": "msm_g1_2^%d_uniform" % LOG_N_MSM, "sample": "2^%d points per step" % REF_SAMPLE_LOG_N},
        "cpu_baseline": {"value": value, "unit": "pts/s", "cores": cores, "kind": "port",
                         "sample": "best_multiexp restatement (C, pthreads; not rayon) on 2^%d uniform points per step"
                                   % REF_SAMPLE_LOG_N},
        "e2e": {"value": value, "unit": "pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="zkb200")
    ap.add_argument("--skip-ntt", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    wd = int(os.environ.get("ZKB_BENCH_WATCHDOG", "0"))
    if wd:  # debugging aid: dump all Python stacks and exit if the run exceeds `wd` seconds
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)

    def note(msg):
        if os.environ.get("ZKB_BENCH_VERBOSE"):
            print("[bench] " + msg, file=sys.stderr, flush=True)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    zkb = importlib.import_module("zksnap-circuits-halo2_b200")
    zkb.init(local_rank)
    lib = zkb.lib()
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic inputs: this rank's point-range shard -------------------------------------------------------------
    n = 1 << LOG_N_MSM
    seed = 0x5EED0000 + LOG_N_MSM + 1000 * rank
    scal_np = random_field(n, seed)
    h_scal = torch.from_numpy(scal_np.view(np.int64)).pin_memory()
    d_scal = h_scal.to(dev, non_blocking=False)
    dlog = random_field(n, seed + 7)
    bases = zkb.g1_fixed_base_mul(dlog)           # [b_i]G on the GPU, known discrete logs
    params = zkb.ParamsKZG(LOG_N_MSM, bases)
    del bases
    out = np.zeros(12, dtype=np.uint64)
    outp = out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
    c_bits, n_win, chunk = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    lib.zkb_msm_get_params(n, ctypes.byref(c_bits), ctypes.byref(n_win), ctypes.byref(chunk))
    # the commit runs through the SRS window table, whose window width is chosen by its own cost model: report that one
    t_bits, t_bytes = ctypes.c_uint32(), ctypes.c_uint64()
    lib.zkb_srs_precompute(params.handle_g, ctypes.byref(t_bits), ctypes.byref(t_bytes))
    if t_bits.value:
        c_bits.value = t_bits.value
        n_win.value = (255 + t_bits.value - 1) // t_bits.value

    def step_dev():
        rc = lib.zkb_msm_g1_srs_dev(params.handle_g, 0, ctypes.c_void_p(d_scal.data_ptr()), n, outp, sptr)
        if rc != 0:
            raise RuntimeError(lib.zkb_last_error().decode())

    def step_e2e():
        rc = lib.zkb_msm_g1_srs(params.handle_g, ctypes.cast(h_scal.data_ptr(), ctypes.POINTER(ctypes.c_uint64)), n, outp)
        if rc != 0:
            raise RuntimeError(lib.zkb_last_error().decode())

    zdist = importlib.import_module("zksnap-circuits-halo2_b200.distributed")

    def fold(result):
        """host fold of the per-rank partial sums (96 B each) — the only inter-GPU exchange of a sharded MSM"""
        if world == 1:
            return result
        return zkb.g1_sum(zdist.all_gather_g1(result, device=dev))

    note("inputs ready")
    clocks = ClockSampler(local_rank)
    clocks.start()
    # ---- warm-up + correctness of the measured configuration (rank-local known-dlog check on the first step) ------------
    for _ in range(args.warmup):
        step_dev()
    # checker leg (the oracle, like the cpu_baseline leg below; outside every timed region): the bases are [b_i]G, so the
    # measured configuration must return [sum s_i b_i]G
    from oracle import coracle

    ip = coracle.fr_inner_product(scal_np, dlog)
    want = coracle.g1_mul(coracle.g1_generator(), ip)
    parity = bool((out[:8] == want).all())

    note("warm-up + parity done")
    # ---- timed: device-resident ------------------------------------------------------------------------------------------
    zkb.prof.enable(True)
    zkb.prof.reset()
    launches0 = zkb.launch_count()
    barrier()
    t_win0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
        total = fold(out)
    e1.record(stream)
    barrier()
    clocks.window(t_win0, time.time())
    ms = e0.elapsed_time(e1)
    launches = zkb.launch_count() - launches0
    acc_ms, acc_calls = zkb.prof.get("msm_accumulate")
    sort_ms, _ = zkb.prof.get("msm_sort")
    dig_ms, _ = zkb.prof.get("msm_digits")
    red_ms, _ = zkb.prof.get("msm_reduce")
    zkb.prof.enable(False)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * n / (ms_per_step * 1e-3)

    note("device-resident timing done")
    # ---- timed: end to end through the host-buffer ABI ---------------------------------------------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    t_win0 = time.time()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
        total = fold(out)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks.window(t_win0, time.time())
    clock_info = clocks.stop()
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n * args.steps / e2e_s

    note("e2e timing done")
    # ---- integer-pipe peak (measured here) and the roofline of the dominant kernel -----------------------------------------
    peak = ctypes.c_double(0)
    lib.zkb_measure_imad_peak.argtypes = [ctypes.POINTER(ctypes.c_double)]
    lib.zkb_measure_imad_peak(ctypes.byref(peak))
    alg_mac = n * n_win.value * 10 * 128          # n*W mixed adds x 10 Fq mul x 128 32-bit MACs
    acc_launch_ms = acc_ms / max(acc_calls, 1)
    achieved = alg_mac / (acc_launch_ms * 1e-3) / 1e9 if acc_launch_ms else 0.0
    roofline = {"bound": "imad", "kernel": "msm_accumulate_kernel<level0> (+partial levels, bucket memset)",
                "achieved": achieved, "peak": peak.value / 1e9, "unit": "GMAC/s (32x32+64 wide MACs)",
                "frac": achieved / (peak.value / 1e9) if peak.value else None,
                "traffic": 28.87e9 if LOG_N_MSM == 24 else None,
                "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of the level-0 launch at 2^24 (profiles/r1c_msm_accumulate_ncu.txt)",
                "peak_source": "measured in this run: unrolled independent mad.wide.u32 chains (zkb_measure_imad_peak)",
                "hbm_view": ({"achieved": 28.87e9 / (acc_launch_ms * 1e-3) / 1e9, "unit": "GB/s",
                              "note": "measured DRAM traffic of the same launch / its duration: the kernel is not HBM bound"}
                             if (LOG_N_MSM == 24 and acc_launch_ms) else None),
                "ms_per_launch": acc_launch_ms, "window_bits": c_bits.value, "windows": n_win.value,
                "share_of_step": acc_launch_ms / ms_per_step if ms_per_step else None,
                "other_ms": {"digits": dig_ms / max(acc_calls, 1), "sort": sort_ms / max(acc_calls, 1), "reduce": red_ms / max(acc_calls, 1)}}

    # ---- secondary: batched NTT ----------------------------------------------------------------------------------------------
    ntt_obj = None
    if not args.skip_ntt:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, src = 6650.0, "fallback"
        if os.path.exists(peaks_path):
            hbm_peak, src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        N = 1 << NTT_LOG_N
        cols = NTT_COLS
        a_np = random_field(N * cols, 0xF0F0 + NTT_LOG_N + rank)
        h_a = torch.from_numpy(a_np.view(np.int64)).pin_memory()
        d_a = h_a.to(dev)
        d_s = torch.empty_like(d_a)
        w = zkb.omega(NTT_LOG_N)
        wp = w.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))

        def ntt_step():
            rc = lib.zkb_ntt_fr_dev(ctypes.c_void_p(d_a.data_ptr()), ctypes.c_void_p(d_s.data_ptr()), cols, wp, NTT_LOG_N, sptr)
            if rc != 0:
                raise RuntimeError(lib.zkb_last_error().decode())

        note("ntt inputs ready")
        for _ in range(args.warmup):
            ntt_step()
        barrier()
        launches1 = zkb.launch_count()
        e0.record(stream)
        for _ in range(args.steps):
            ntt_step()
        e1.record(stream)
        barrier()
        nms = e0.elapsed_time(e1) / args.steps
        launches += zkb.launch_count() - launches1
        alg_bytes = 64.0 * N * cols
        gbs = alg_bytes / (nms * 1e-3) / 1e9
        note("ntt device timing done")
        # e2e: host columns through zkb_ntt_fr_batch (H2D + kernels + D2H)
        cols_np = [a_np[i * N:(i + 1) * N] for i in range(cols)]
        ptrs = (ctypes.POINTER(ctypes.c_uint64) * cols)(*[ctypes.cast(h_a.data_ptr() + i * N * 32, ctypes.POINTER(ctypes.c_uint64)) for i in range(cols)])
        lib.zkb_ntt_fr_batch(ptrs, cols, wp, NTT_LOG_N)
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, args.steps // 2)):
            lib.zkb_ntt_fr_batch(ptrs, cols, wp, NTT_LOG_N)
        barrier()
        ntt_e2e_s = (time.perf_counter() - t0) / max(1, args.steps // 2)
        ntt_obj = {"workload": "best_fft 2^%d x %d columns (batched, in HBM)" % (NTT_LOG_N, cols),
                   "value": world * N * cols / (nms * 1e-3), "unit": "elems/s", "ms_per_step": nms,
                   "e2e": {"value": world * N * cols / ntt_e2e_s, "unit": "elems/s", "h2d_bytes_per_step": N * cols * 32,
                           "d2h_bytes_per_step": N * cols * 32},
                   "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                "traffic": 12.9e9 if (NTT_LOG_N, NTT_COLS) == (22, 16) else None, "peak_source": src + " (MEASURED_PEAKS.json hbm_gbs)",
                                "note": "64 B algorithmic bytes per element; the kernel is integer-issue bound, see DESIGN.md"}}
        del cols_np

    # ---- secondary: quotient evaluation on resident cosets (GraphEvaluator row loop) ------------------------
    quot_obj = None
    if not args.skip_ntt:
        ev = importlib.import_module("zksnap-circuits-halo2_b200.evaluation")
        rows, qc, rs = 1 << QUOT_LOG_N, QUOT_COLS, 4
        gr = ev.GraphEvaluator()
        parts = []
        for i in range(qc):   # halo2-base: q_i * (a + b * c - d), a..d = advice column i at rotations 0..3
            a_, b_, c_, d_ = (("advice", i, r) for r in range(4))
            parts.append(gr.add_expression(("prod", ("fixed", i, 0), ("sum", ("sum", a_, ("prod", b_, c_)), ("neg", d_)))))
        gr.add_horner(ev.ValueSource(ev.PREVIOUS), parts, ev.ValueSource(ev.Y))
        adv_np, sel_np = random_field(rows, 0x9A7E + rank), random_field(rows, 0x5E1 + rank)
        y_np = random_field(1, 0x77)[0]
        adv = [zkb.Polynomial(adv_np) for _ in range(qc)]   # same values, distinct HBM buffers: the traffic is real
        sel = [zkb.Polynomial(sel_np) for _ in range(qc)]
        vals = zkb.Polynomial(np.zeros((rows, 4), dtype=np.uint64))
        note("quotient inputs ready")
        gr.evaluate(vals, fixed=sel, advice=adv, y=y_np, rot_scale=rs)
        # parity of the measured configuration on sampled rows, Python integers (previous value 0):
        # value = gate * (y^(qc-1) + ... + 1), gate = q (a + b c - d)
        FRM = ev.FR
        Rinv = pow(1 << 256, FRM - 2, FRM)
        li = lambda v: sum(int(v[j]) << (64 * j) for j in range(4)) * Rinv % FRM  # noqa: E731
        got = vals.to_host()
        yv = li(y_np)
        ysum = sum(pow(yv, j, FRM) for j in range(qc)) % FRM
        q_ok = True
        for r in [0, 1, rows - 1, rows - 5, rows // 3, rows // 2 + 7]:
            av = [li(adv_np[(r + j * rs) % rows]) for j in range(4)]
            q_ok &= li(got[r]) == li(sel_np[r]) * (av[0] + av[1] * av[2] - av[3]) % FRM * ysum % FRM
        del got
        for _ in range(args.warmup):
            gr.evaluate(vals, fixed=sel, advice=adv, y=y_np, rot_scale=rs)
        barrier()
        launches3 = zkb.launch_count()
        zkb.prof.enable(True)
        zkb.prof.reset()
        for _ in range(args.steps):
            gr.evaluate(vals, fixed=sel, advice=adv, y=y_np, rot_scale=rs)
        barrier()
        qms, qk = zkb.prof.get("graph_evaluate")   # CUDA events around the kernel on the library stream
        qms /= max(qk, 1)
        zkb.prof.enable(False)
        launches += zkb.launch_count() - launches3
        info = gr.last_info()
        qgbs = info["bytes_per_row"] * rows / (qms * 1e-3) / 1e9
        quot_obj = {"workload": "evaluate_h custom gates: halo2-base gate q(a+bc-d) on %d advice columns, 2^%d extended rows, rot_scale %d, resident in HBM" % (qc, QUOT_LOG_N, rs),
                    "value": world * rows / (qms * 1e-3), "unit": "rows/s", "ms_per_step": qms, "parity_checked": bool(q_ok),
                    "lowered": info, "modmul_per_row": 3 * qc,
                    "roofline": {"bound": "hbm", "achieved": qgbs, "peak": hbm_peak, "unit": "GB/s", "frac": qgbs / hbm_peak,
                                 "traffic": 5.362e9 if (QUOT_LOG_N, QUOT_COLS) == (24, 4) else None,   # ncu, profiles/r1d_graph_ncu.txt
                                 "note": "32 B x (polynomials read + previous value + result) per row; integer-issue bound, see DESIGN.md"}}
        parity = parity and bool(q_ok)
        for p_ in adv + sel + [vals]:
            p_.free()
        del adv_np, sel_np
        note("quotient done")

    # ---- N > 1: one 2^26 NTT sharded over the ranks (exchange fused into the NTT passes over NVLink peer memory) ----------------
    sharded_obj = None
    if world > 1 and not args.skip_ntt and (world & (world - 1)) == 0 and world <= 8:
        k = SHARDED_LOG_N
        sh = zdist.ShardedNtt(k, device=dev)
        off, ln = zdist.ntt_slice(k, rank, world)
        w = zkb.omega(k)
        wp = w.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
        # parity of the measured configuration, size-independent: NTT(delta_0) = (1, 1, ..., 1)
        d_in = torch.zeros(ln * 4, dtype=torch.int64, device=dev)
        one = torch.from_numpy(np.array([0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f],
                                        dtype=np.uint64).view(np.int64)).to(dev)
        if rank == 0:
            d_in[:4] = one
        d_out = torch.empty_like(d_in)
        rc = lib.zkb_dist_ntt_fr_dev(ctypes.c_void_p(d_in.data_ptr()), ctypes.c_void_p(d_out.data_ptr()), wp, k, sptr)
        if rc != 0 or lib.zkb_dist_status(sptr) != 0:
            raise RuntimeError(lib.zkb_last_error().decode())
        sh_ok = bool((d_out.view(-1, 4) == one).all().item())
        g = torch.Generator(device=dev)
        g.manual_seed(0xD157 + rank)
        d_in = torch.randint(0, 1 << 60, (ln * 4,), dtype=torch.int64, device=dev, generator=g)
        lib.zkb_dist_ntt_fr_dev(ctypes.c_void_p(d_in.data_ptr()), None, wp, k, sptr)  # loads the symmetric input slice
        for _ in range(args.warmup):
            lib.zkb_dist_ntt_fr_dev(None, None, wp, k, sptr)
        barrier()
        launches2 = zkb.launch_count()
        e0.record(stream)
        for _ in range(args.steps):
            lib.zkb_dist_ntt_fr_dev(None, None, wp, k, sptr)
        e1.record(stream)
        if lib.zkb_dist_status(sptr) != 0:
            raise RuntimeError(lib.zkb_last_error().decode())
        barrier()
        sms = e0.elapsed_time(e1) / args.steps
        launches += zkb.launch_count() - launches2
        t = torch.tensor([sms, 0.0 if sh_ok else 1.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sms, sh_ok = float(t[0].item()), t[1].item() == 0.0
        N = 1 << k
        sharded_obj = {"workload": "one best_fft of 2^%d sharded over %d GPUs (contiguous slices in, contiguous slices out)" % (k, world),
                       "value": N / (sms * 1e-3), "unit": "elems/s", "ms_per_step": sms, "parity_checked": sh_ok,
                       "exchange": "fused into NTT pass 0 (peer loads+stores) and the final pass (peer stores) over NVLink; device-side barriers",
                       "nvlink_bytes_per_gpu_per_step": int(3 * (world - 1) / world * ln * 32),
                       "roofline": {"bound": "hbm", "achieved": 64.0 * N / (sms * 1e-3) / 1e9, "peak": hbm_peak * world, "unit": "GB/s",
                                    "frac": 64.0 * N / (sms * 1e-3) / 1e9 / (hbm_peak * world)}}
        parity = parity and sh_ok
        sh.close()
        del d_in, d_out
    # ---- N > 1: the same quotient evaluation sharded by rows (ring halo exchange over NCCL, then the row window on every rank) ------
    sq_obj = None
    if world > 1 and not args.skip_ntt and (world & (world - 1)) == 0:
        try:
            srows = (1 << QUOT_LOG_N) // world
            sq = zdist.ShardedQuotient(gr, 4)
            gen = torch.Generator(device=dev)
            gen.manual_seed(0x51AB)   # the same rows on every rank: the domain is periodic with period srows, so the sharded result
            base = torch.randint(0, 1 << 60, (srows, 4), dtype=torch.int64, device=dev, generator=gen)  # must equal a wrapping evaluation of one period
            bufs = []
            for _ in range(2 * QUOT_COLS):
                buf, view = sq.alloc_column(srows, dev)
                view.copy_(base)
                bufs.append(buf)
            sel_b, adv_b = bufs[:QUOT_COLS], bufs[QUOT_COLS:]
            vals_s = torch.zeros_like(base)
            sq.run_padded(vals_s, sel_b, adv_b, [], y=y_np)
            ref_s = torch.zeros_like(base)
            plain = [base.clone() for _ in range(2)]
            gr.evaluate_dev(ref_s.data_ptr(), srows, [plain[0].data_ptr()] * QUOT_COLS, [plain[1].data_ptr()] * QUOT_COLS, y=y_np, rot_scale=4,
                            stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            sq_ok = bool(torch.equal(vals_s, ref_s))
            for _ in range(args.warmup):
                sq.run_padded(vals_s, sel_b, adv_b, [], y=y_np)
            barrier()
            launches4 = zkb.launch_count()
            e0.record(stream)
            for _ in range(args.steps):
                sq.run_padded(vals_s, sel_b, adv_b, [], y=y_np)
            e1.record(stream)
            barrier()
            qsms = e0.elapsed_time(e1) / args.steps
            launches += zkb.launch_count() - launches4
            t = torch.tensor([qsms, 0.0 if sq_ok else 1.0], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            qsms, sq_ok = float(t[0].item()), t[1].item() == 0.0
            sq_obj = {"workload": "the quotient workload above, 2^%d rows sharded by rows over %d GPUs" % (QUOT_LOG_N, world),
                      "value": (1 << QUOT_LOG_N) / (qsms * 1e-3), "unit": "rows/s", "ms_per_step": qsms, "parity_checked": sq_ok,
                      "exchange": "ring halo exchange, %d + %d rows per column over NCCL point-to-point, then zkb_graph_evaluate_dev on the row window" % (sq.halo_lo, sq.halo_hi),
                      "halo_bytes_per_rank": (sq.halo_lo + sq.halo_hi) * 32 * 2 * QUOT_COLS}
            parity = parity and sq_ok
            del bufs, vals_s, ref_s, plain, base
        except Exception as exc:  # a secondary object must not take the headline line down
            sq_obj = {"error": repr(exc)[:300]}
    note("ntt done")
    # ---- CPU baseline on this box (rank 0, N=1 only) ---------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        v, cores, dt = cpu_baseline_msm(CPU_SAMPLE_LOG_N)
        cpu = {"value": v, "unit": "pts/s", "cores": cores, "kind": "port",
               "sample": "best_multiexp restatement (C, pthreads; not rayon) on 2^%d uniform points, %.1f s" % (CPU_SAMPLE_LOG_N, dt)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "pts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (254-bit Montgomery integers)", "data": "synthetic",
            "config": {"workload": "msm_g1_2^%d_uniform_per_gpu" % LOG_N_MSM, "sharding": "srs_point_range_per_rank, host fold",
                       "l2": "inputs_exceed_l2 (512 MiB scalars + 1 GiB bases per step)", "window_bits": c_bits.value,
                       "windows": n_win.value, "srs_window_table_bytes": int(t_bytes.value), "chunk": chunk.value},
            "e2e": {"value": e2e_value, "unit": "pts/s", "h2d_bytes_per_step": n * 32 * world,
                    "d2h_bytes_per_step": n_win.value * 128 * world, "timer": "wall clock around the C-ABI call (includes host fold)"},
            "gpu_launches": int(launches), "parity_checked": parity, "roofline": roofline, "cpu_baseline": cpu,
            "clocks": clock_info, "ntt": ntt_obj, "quotient": quot_obj, "sharded_ntt": sharded_obj, "sharded_quotient": sq_obj,
        }
        print(json.dumps(line))
    params.close()
    if world > 1:
        dist.destroy_process_group()
    return 0 if parity else 1


if __name__ == "__main__":
    sys.exit(main())
