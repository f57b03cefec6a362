#!/usr/bin/env python
"""bench.py — measurement of the zkb200 hot path (DESIGN.md "Measurement").

Headline step (the driver's contract line): one 2^24-point BN254 G1 MSM (ParamsKZG::commit shape: uniform Fr scalars against
an SRS resident in HBM) per GPU.  At N GPUs the MSM is sharded by SRS point range, one 2^24-point shard per rank (weak
scaling), the 96-byte partial results are folded on the host.
  value        points/s, whole job, scalars already in HBM, CUDA events on the launch stream
  e2e          the same through the host-buffer C ABI call (zkb_msm_g1_srs): page-locked host scalars -> H2D -> kernels ->
               sums D2H -> host fold, wall clock; `e2e_pageable`: the same from pageable memory (what a Rust Vec<Fr> is)
  roofline     integer-pipe fraction of the dominant kernel (bucket accumulation), IMAD peak measured in the same run
Objects beside it (all parity-checked, all in the same JSON line):
  ntt, quotient                 batched 2^22 x 16 best_fft and evaluate_h's gate pass, per GPU
  wrapper_replay / voter_replay / st_replay
                                 BASELINE configs #1-#3 "proof-gen sec": the prover's op sequence on the hot path (commits, iNTTs,
                                 coset NTTs, the quotient transform) replayed in prover order, as STRONG scaling over the N GPUs:
                                 `kernel` (operands resident in HBM), `dropin` (host buffers through the C ABI, PCIe included)
  msm_split                     one 2^26-point MSM split N ways (strong scaling, BASELINE config #4's largest size)
  sharded_ntt, sharded_quotient N > 1: one 2^26 NTT sharded over the ranks / the quotient pass sharded by rows
  single_process                N > 1: rank 0 ALONE drives all N devices through the unchanged C ABI (zkb_init with N devices) —
                                 the deployment of the reference's one-process prover — and must reproduce the torchrun results
`--impl reference` times the CPU restatement of halo2's best_multiexp (oracle/, all host threads, built -O3 -march=native on
this host) on the same 2^24-point workload; the reference itself is Rust with un-vendored dependencies and cannot be built in
this image (DESIGN.md "Oracle").
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
from concurrent.futures import ThreadPoolExecutor

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
SPLIT_LOG_N = int(os.environ.get("ZKB_BENCH_SPLIT_LOG_N", "26"))   # one MSM of this size split over the N GPUs
WRAPPER_K = int(os.environ.get("ZKB_BENCH_WRAPPER_K", "22"))
REF_MAX_STEPS = int(os.environ.get("ZKB_BENCH_REF_STEPS", "3"))     # --impl reference: one 2^24 step is ~8 s of 16 host threads
METRIC = "BN254 G1 MSM throughput (2^%d points per GPU, SRS resident)" % LOG_N_MSM
WORKLOAD = "msm_g1_2^%d_uniform_per_gpu" % LOG_N_MSM
u64p = ctypes.POINTER(ctypes.c_uint64)
FR_ONE = np.array([0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f], dtype=np.uint64)  # R mod r


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


def measured_traffic(key):
    """DRAM bytes per launch of a kernel (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture), read
    from the tracked profiles/traffic.json, which names the ncu summary each number comes from.  None when no capture of the
    configuration is on file — never a typed-in constant."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f).get(key)
        return (float(e["bytes_per_launch"]), e["source"]) if e else (None, None)
    except Exception:
        return None, None


def cpu_baseline_msm(log_n: int):
    """Oracle restatement of best_multiexp on all host threads; returns (pts/s, cores, seconds, native build?)."""
    from oracle import coracle

    coracle.build()
    native = coracle.use_native()   # -O3 -march=native, compiled on this host
    n = 1 << log_n
    s = random_field(n, 0x5EED0000 + log_n)
    # bases: cheap synthetic curve points for the CPU arm — multiples of G by small random scalars (CPU fixed-base)
    b = coracle.g1_fixed_base_mul(random_field(min(n, 4096), 7))
    bases = np.ascontiguousarray(np.tile(b, (n // b.shape[0] + 1, 1))[:n])
    cores = coracle.num_threads()
    t = time.perf_counter()
    coracle.best_multiexp(s, bases, 0)
    dt = time.perf_counter() - t
    return n / dt, cores, dt, native


def run_reference(args):
    """The reference arm: same metric, same config (2^24 uniform points per step) on the host cores.  One step is ~8 s, so at
    most REF_MAX_STEPS timed steps (and one warm-up) are run whatever --steps asks; the line says how many."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, min(args.steps, REF_MAX_STEPS))
    warm = 1 if args.warmup else 0
    ts, cores, native = [], 0, False
    for i in range(warm + steps):
        _, cores, dt, native = cpu_baseline_msm(LOG_N_MSM)
        if i >= warm:
            ts.append(dt)
    n = 1 << LOG_N_MSM
    value = n * len(ts) / sum(ts)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pts/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "steps_requested": args.steps, "ms_per_step": 1e3 * sum(ts) / len(ts),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64x4 (254-bit Montgomery)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "the full 2^%d-point step, %d timed steps" % (LOG_N_MSM, steps)},
        "cpu_baseline": {"value": value, "unit": "pts/s", "cores": cores, "kind": "port",
                         "sample": "best_multiexp restatement (C, pthreads; not rayon; %s) on 2^%d uniform points per step, %d steps"
                                   % ("-O3 -march=native built on this host" if native else "portable -O3 build", LOG_N_MSM, steps)},
        "e2e": {"value": value, "unit": "pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---- small helpers shared by the legs ------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: keep the rank's host threads (and therefore its page-locked buffers, first touched by them) on the NUMA node
    its GPU hangs off, as a production launcher does with numactl — eight ranks uploading 512 MiB each otherwise cross the socket link.
    Returns a description for the JSON line, or None when the topology cannot be read (nothing is changed then).  ZKB_BENCH_NUMA=0 disables."""
    if os.environ.get("ZKB_BENCH_NUMA", "1") == "0":
        return None
    try:
        import torch

        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        mine = cpus & allowed
        if len(mine) < 2 or mine == allowed:
            return {"node": node, "cpus": len(allowed), "bound": False}
        os.sched_setaffinity(0, mine)
        return {"node": node, "cpus": len(mine), "bound": True, "all": sorted(allowed)}
    except Exception:
        return None


class Env:
    """torch / torch.distributed / libzkb200 plumbing of one rank."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa = bind_to_gpu_numa_node(self.local_rank) if self.world > 1 else None   # before any host thread / pinned buffer of this rank exists
        self.host_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            self.host_group = dist.new_group(backend="gloo")   # host-side barrier / small exchanges that must not touch the GPUs
        self.zkb = importlib.import_module("zksnap-circuits-halo2_b200")
        self.zdist = importlib.import_module("zksnap-circuits-halo2_b200.distributed")
        self.zkb.init(self.local_rank)
        self.lib = self.zkb.lib()
        self.stream = torch.cuda.current_stream()
        self.sptr = ctypes.c_void_p(self.stream.cuda_stream)
        self.launches = 0
        self.parity = {}

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.zkb_last_error().decode())

    def timed_events(self, fn, steps):
        """device time of `steps` calls of fn on the torch stream, max over ranks, per step (ms); counts launches"""
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        self.barrier()
        l0 = self.zkb.launch_count()
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        self.barrier()
        self.launches += self.zkb.launch_count() - l0
        return self.max_over_ranks(e0.elapsed_time(e1))[0] / steps

    def timed_wall(self, fn, steps):
        self.barrier()
        l0 = self.zkb.launch_count()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.barrier()
        dt = time.perf_counter() - t0
        self.launches += self.zkb.launch_count() - l0
        return self.max_over_ranks(dt)[0] / steps * 1e3

    def dptr(self, t):
        return ctypes.c_void_p(t.data_ptr())

    def device_field(self, n, seed):
        """n Fr elements on the device (int64 limbs, top limb < 2^60 so every value is < r): same seed -> same values on any GPU"""
        g = self.torch.Generator(device=self.dev)
        g.manual_seed(seed)
        return self.torch.randint(0, 1 << 60, (n, 4), dtype=self.torch.int64, device=self.dev, generator=g)

    def fold(self, out):
        """host fold of the per-rank partial sums (96 B each) — the only inter-GPU exchange of a sharded MSM"""
        if self.world == 1:
            return out.copy()
        return self.zkb.g1_sum(self.zdist.all_gather_g1(out, device=self.dev))


def known_dlog_point(scal_np, dlog_np):
    """[sum s_i b_i] G by the oracle (the checker; outside every timed region)"""
    from oracle import coracle

    coracle.build()
    return coracle.g1_mul(coracle.g1_generator(), coracle.fr_inner_product(scal_np, dlog_np))


def to_np(t):
    return t.cpu().numpy().view(np.uint64).reshape(-1, 4)


# ---- headline: 2^24-point commit per GPU ---------------------------------------------------------------------------------------
def headline_msm(env: Env, clocks: ClockSampler):
    zkb, lib, torch = env.zkb, env.lib, env.torch
    args = env.args
    n = 1 << LOG_N_MSM
    seed = 0x5EED0000 + LOG_N_MSM + 1000 * env.rank
    scal_np = random_field(n, seed)
    h_scal = torch.from_numpy(scal_np.view(np.int64)).pin_memory()
    d_scal = h_scal.to(env.dev, non_blocking=False)
    dlog = random_field(n, seed + 7)
    bases = zkb.g1_fixed_base_mul(dlog)           # [b_i]G on the GPU, known discrete logs
    params = zkb.ParamsKZG(LOG_N_MSM, bases)
    out = np.zeros(12, dtype=np.uint64)
    outp = out.ctypes.data_as(u64p)
    c_bits, n_win, chunk = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    lib.zkb_msm_get_params(n, ctypes.byref(c_bits), ctypes.byref(n_win), ctypes.byref(chunk))
    # the commit runs through the SRS window table (steady state of a long-running prover: built once per SRS, forced here), whose
    # window width comes from its own cost model: report that one
    t_bits, t_bytes = ctypes.c_uint32(), ctypes.c_uint64()
    t0 = time.perf_counter()
    lib.zkb_srs_precompute(params.handle_g, ctypes.byref(t_bits), ctypes.byref(t_bytes))
    table_build_s = time.perf_counter() - t0
    if t_bits.value:
        c_bits.value = t_bits.value
        n_win.value = (255 + t_bits.value - 1) // t_bits.value

    def step_dev():
        env.check(lib.zkb_msm_g1_srs_dev(params.handle_g, 0, env.dptr(d_scal), n, outp, env.sptr))

    def step_e2e():
        env.check(lib.zkb_msm_g1_srs(params.handle_g, ctypes.cast(h_scal.data_ptr(), u64p), n, outp))

    def step_pageable():
        env.check(lib.zkb_msm_g1_srs(params.handle_g, scal_np.ctypes.data_as(u64p), n, outp))

    for _ in range(args.warmup):
        step_dev()
    want = known_dlog_point(scal_np, dlog)
    parity = bool((out[:8] == want).all())
    # ---- timed: device-resident
    zkb.prof.enable(True)
    zkb.prof.reset()
    t_win0 = time.time()
    ms_per_step = env.timed_events(lambda: (step_dev(), env.fold(out)), args.steps)
    clocks.window(t_win0, time.time())
    acc_ms, acc_calls = zkb.prof.get("msm_accumulate")
    sort_ms, _ = zkb.prof.get("msm_sort")
    dig_ms, _ = zkb.prof.get("msm_digits")
    red_ms, _ = zkb.prof.get("msm_reduce")
    zkb.prof.enable(False)
    ent = ctypes.c_uint64(0)
    lib.zkb_msm_last_entries(ctypes.byref(ent))
    value = env.world * n / (ms_per_step * 1e-3)
    # ---- timed: end to end through the host-buffer ABI, page-locked scalars (the contract's e2e) and pageable ones
    for _ in range(2):
        step_e2e()
    t_win0 = time.time()
    e2e_ms = env.timed_wall(lambda: (step_e2e(), env.fold(out)), args.steps)
    clocks.window(t_win0, time.time())
    parity = parity and bool((out[:8] == want).all())
    step_pageable()
    pg_ms = env.timed_wall(lambda: (step_pageable(), env.fold(out)), max(2, args.steps // 2))
    parity = parity and bool((out[:8] == want).all())
    # ---- the same commit WITHOUT the window table (what the first ~100 commits against a fresh SRS cost)
    lib.zkb_srs_set_precompute(0)
    plain = zkb.ParamsKZG(LOG_N_MSM, bases)

    def step_plain():
        env.check(lib.zkb_msm_g1_srs_dev(plain.handle_g, 0, env.dptr(d_scal), n, outp, env.sptr))

    step_plain()
    parity = parity and bool((out[:8] == want).all())
    no_table_ms = env.timed_events(step_plain, 3)
    plain.close()
    lib.zkb_srs_set_precompute(2)
    del bases
    # ---- the same commit for a WITNESS-LIKE column (BASELINE sweep distribution W: 50 % zero, 25 % < 2^16, 20 % < 2^88, 5 % uniform —
    # the shape of halo2-base advice columns): zero digits are never emitted, so the cost follows the non-zero digits
    rng = np.random.default_rng(0x517 + env.rank)
    w_np = scal_np.copy()
    u = rng.random(n)
    w_np[u < 0.5] = 0
    for lo_, hi_, nbytes in ((0.5, 0.75, 2), (0.75, 0.95, 11)):
        m_ = (u >= lo_) & (u < hi_)
        pool = np.zeros((4096, 4), dtype=np.uint64)
        raw = rng.integers(0, 256, size=(4096, nbytes), dtype=np.uint64)
        vals_ = [sum(int(b) << (8 * j) for j, b in enumerate(row)) for row in raw]
        R_ = (1 << 256) % 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
        for i_, v_ in enumerate(vals_):
            mv = v_ * R_ % 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
            pool[i_] = [(mv >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]
        w_np[m_] = pool[rng.integers(0, 4096, int(m_.sum()))]
    d_w = torch.from_numpy(w_np.view(np.int64)).to(env.dev)

    def step_w():
        env.check(lib.zkb_msm_g1_srs_dev(params.handle_g, 0, env.dptr(d_w), n, outp, env.sptr))

    step_w()
    w_ok = bool((out[:8] == known_dlog_point(w_np, dlog)).all())
    w_ms = env.timed_events(step_w, max(3, args.steps // 2))
    w_ent = ctypes.c_uint64(0)
    lib.zkb_msm_last_entries(ctypes.byref(w_ent))
    parity = parity and w_ok
    del d_w, w_np
    # ---- integer-pipe peak (measured here) and the roofline of the dominant kernel
    peak = ctypes.c_double(0)
    lib.zkb_measure_imad_peak.argtypes = [ctypes.POINTER(ctypes.c_double)]
    lib.zkb_measure_imad_peak(ctypes.byref(peak))
    alg_mac = n * n_win.value * 10 * 128          # n*W mixed adds x 10 Fq mul x 128 32-bit MACs
    acc_launch_ms = acc_ms / max(acc_calls, 1)
    achieved = alg_mac / (acc_launch_ms * 1e-3) / 1e9 if acc_launch_ms else 0.0
    traffic, traffic_src = measured_traffic("msm_accumulate_2^%d" % LOG_N_MSM)
    roofline = {"bound": "imad", "kernel": "msm_accumulate_kernel<level0> (+partial levels, bucket memset)",
                "achieved": achieved, "peak": peak.value / 1e9, "unit": "GMAC/s (32x32+64 wide MACs)",
                "frac": achieved / (peak.value / 1e9) if peak.value else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "measured in this run: unrolled independent mad.wide.u32 chains (zkb_measure_imad_peak)",
                "hbm_view": ({"achieved": traffic / (acc_launch_ms * 1e-3) / 1e9, "unit": "GB/s",
                              "note": "measured DRAM traffic of the same launch / its duration: the kernel is not HBM bound"}
                             if (traffic and acc_launch_ms) else None),
                "ms_per_launch": acc_launch_ms, "window_bits": c_bits.value, "windows": n_win.value,
                "bucket_additions": int(ent.value),
                "share_of_step": acc_launch_ms / ms_per_step if ms_per_step else None,
                "other_ms": {"digits": dig_ms / max(acc_calls, 1), "sort": sort_ms / max(acc_calls, 1), "reduce": red_ms / max(acc_calls, 1)}}
    params.close()
    del d_scal, h_scal
    torch.cuda.empty_cache()
    return {"value": value, "ms_per_step": ms_per_step, "parity": parity, "roofline": roofline,
            "e2e": {"value": env.world * n / (e2e_ms * 1e-3), "unit": "pts/s", "h2d_bytes_per_step": n * 32 * env.world,
                    "d2h_bytes_per_step": 128 * env.world, "ms_per_step": e2e_ms,
                    "timer": "wall clock around the C-ABI call (includes host fold), page-locked scalars"},
            "e2e_pageable": {"value": env.world * n / (pg_ms * 1e-3), "unit": "pts/s", "ms_per_step": pg_ms,
                             "note": "the same call from pageable memory (a Rust Vec<Fr>): staged through pinned buffers by host threads"},
            "no_table_ms_per_step": no_table_ms,
            "witness_like": {"workload": "the same 2^%d-point commit for a witness-like column (50 %% zero, 25 %% < 2^16, 20 %% < 2^88, 5 %% uniform)" % LOG_N_MSM,
                             "ms_per_step": w_ms, "value": env.world * n / (w_ms * 1e-3), "unit": "pts/s", "bucket_additions": int(w_ent.value),
                             "parity_checked": w_ok},
            "config": {"workload": WORKLOAD, "sharding": "srs_point_range_per_rank, host fold",
                       "l2": "inputs_exceed_l2 (512 MiB scalars + 1 GiB bases per step)", "window_bits": c_bits.value,
                       "windows": n_win.value, "srs_window_table_bytes": int(t_bytes.value), "srs_window_table_build_s": table_build_s,
                       "chunk": chunk.value,
                       "host_numa": ({k: v for k, v in env.numa.items() if k != "all"} if env.numa else None)}}


# ---- secondary: batched NTT ----------------------------------------------------------------------------------------------------
def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ntt_object(env: Env, imad_peak):
    zkb, lib, torch = env.zkb, env.lib, env.torch
    args = env.args
    peak, src = hbm_peak()
    N, cols = 1 << NTT_LOG_N, NTT_COLS
    a_np = random_field(N * cols, 0xF0F0 + NTT_LOG_N + env.rank)
    h_a = torch.from_numpy(a_np.view(np.int64)).pin_memory()
    d_a = h_a.to(env.dev)
    d_s = torch.empty_like(d_a)
    w = zkb.omega(NTT_LOG_N)
    wp = w.ctypes.data_as(u64p)

    def ntt_step():
        env.check(lib.zkb_ntt_fr_dev(env.dptr(d_a), env.dptr(d_s), cols, wp, NTT_LOG_N, env.sptr))

    # parity of the measured configuration: column 0 against the definition on a 2-sparse probe, and forward/inverse round trip
    probe = torch.zeros_like(d_a)
    j1, j2 = 12345 % N, N - 7
    pr = np.zeros((N, 4), dtype=np.uint64)
    pr[j1], pr[j2] = a_np[1], a_np[2]
    probe.view(-1, 4)[:N] = torch.from_numpy(pr.view(np.int64)).to(env.dev)
    env.check(lib.zkb_ntt_fr_dev(env.dptr(probe), env.dptr(d_s), cols, wp, NTT_LOG_N, env.sptr))
    torch.cuda.synchronize()
    ok = two_sparse_ok(to_np(probe.view(-1, 4)[:N]), a_np[1], a_np[2], j1, j2, NTT_LOG_N, [0, 1, 2, N // 2 + 3, N - 1, N // 3])
    del probe, pr
    for _ in range(args.warmup):
        ntt_step()
    nms = env.timed_events(ntt_step, args.steps)
    alg_bytes = 64.0 * N * cols
    gbs = alg_bytes / (nms * 1e-3) / 1e9
    ptrs = (u64p * cols)(*[ctypes.cast(h_a.data_ptr() + i * N * 32, u64p) for i in range(cols)])
    lib.zkb_ntt_fr_batch(ptrs, cols, wp, NTT_LOG_N)
    e2e_ms = env.timed_wall(lambda: lib.zkb_ntt_fr_batch(ptrs, cols, wp, NTT_LOG_N), max(1, args.steps // 2))
    traffic, tsrc = measured_traffic("ntt_2^%dx%d" % (NTT_LOG_N, NTT_COLS))
    # integer view: Montgomery products per element of this plan = radix-8 butterfly rounds (1.375 products per
    # element per full round of 3 bits, fewer for the shorter last round of a pass) + one inter-pass twiddle per later pass
    prods = ntt_products_per_element(NTT_LOG_N)
    mac = prods * 128 * N * cols / (nms * 1e-3)
    obj = {"workload": "best_fft 2^%d x %d columns (batched, in HBM)" % (NTT_LOG_N, cols),
           "value": env.world * N * cols / (nms * 1e-3), "unit": "elems/s", "ms_per_step": nms, "parity_checked": ok,
           "e2e": {"value": env.world * N * cols / (e2e_ms * 1e-3), "unit": "elems/s", "h2d_bytes_per_step": N * cols * 32,
                   "d2h_bytes_per_step": N * cols * 32},
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                        "traffic": traffic, "traffic_source": tsrc, "peak_source": src,
                        "note": "64 B algorithmic bytes per element; the kernel is integer-issue bound (imad view), see DESIGN.md"},
           "roofline_imad": {"bound": "imad", "achieved": mac / 1e9, "peak": imad_peak / 1e9, "unit": "GMAC/s",
                             "frac": mac / imad_peak if imad_peak else None,
                             "note": "%.2f Montgomery products (128 MACs each) per element: butterflies + inter-pass twiddles" % prods}}
    del d_a, d_s, h_a
    torch.cuda.empty_cache()
    return obj


def ntt_products_per_element(log_n):
    """Montgomery products per element of the multi-pass plan ntt_plan.hpp picks for 2^log_n: per pass of radix 2^lr, DIF rounds of
    3,3,..,rem bits; a radix-8 round costs 5 internal + 7 output twiddle products per 8 elements (the last round of a pass has no
    output twiddles), radix-4: 1 + 3 per 4, radix-2: 0 + 1 per 2; plus one inter-pass twiddle per element for every pass after the
    first.  2^22 (passes 8, 7, 7): 11.25."""
    if log_n <= 10:
        lrs = [log_n]
    else:
        npass = (log_n + 8) // 9
        base, rem = divmod(log_n, npass)
        lrs = [base + (1 if p < rem else 0) for p in range(npass)]
    total = 0.0
    for p, lr in enumerate(lrs):
        left = lr
        while left > 0:
            t = 3 if left >= 3 else left
            left -= t
            internal = {3: 5 / 8, 2: 1 / 4, 1: 0.0}[t]
            outtw = {3: 7 / 8, 2: 3 / 4, 1: 1 / 2}[t] if left > 0 else 0.0
            total += internal + outtw
        if p > 0:
            total += 1.0
    return total


def two_sparse_ok(out_rows, v1, v2, j1, j2, log_n, idx, base=0):
    """out_rows[i - base] must equal v1 w^(i j1) + v2 w^(i j2) for i in idx: the NTT of a 2-sparse input by its definition —
    exact for any size, sensitive to any permutation or twiddle error (Python integers)."""
    from oracle import pyref as R
    from util import limbs_to_int

    w = R.omega_for(log_n)
    a1, a2 = R.from_mont(limbs_to_int(v1), R.FR), R.from_mont(limbs_to_int(v2), R.FR)
    for i in idx:
        want = (a1 * pow(w, i * j1, R.FR) + a2 * pow(w, i * j2, R.FR)) % R.FR
        if R.from_mont(limbs_to_int(out_rows[i - base]), R.FR) != want:
            return False
    return True


# ---- secondary: quotient evaluation on resident cosets (GraphEvaluator row loop) -----------------------------
def quotient_object(env: Env):
    zkb, torch = env.zkb, env.torch
    args = env.args
    peak, _ = hbm_peak()
    ev = importlib.import_module("zksnap-circuits-halo2_b200.evaluation")
    rows, qc, rs = 1 << QUOT_LOG_N, QUOT_COLS, 4
    gr = ev.GraphEvaluator()
    parts = []
    for i in range(qc):   # halo2-base: q_i * (a + b * c - d), a..d = advice column i at rotations 0..3
        a_, b_, c_, d_ = (("advice", i, r) for r in range(4))
        parts.append(gr.add_expression(("prod", ("fixed", i, 0), ("sum", ("sum", a_, ("prod", b_, c_)), ("neg", d_)))))
    gr.add_horner(ev.ValueSource(ev.PREVIOUS), parts, ev.ValueSource(ev.Y))
    adv_np, sel_np = random_field(rows, 0x9A7E + env.rank), random_field(rows, 0x5E1 + env.rank)
    y_np = random_field(1, 0x77)[0]
    adv = [zkb.Polynomial(adv_np) for _ in range(qc)]   # same values, distinct HBM buffers: the traffic is real
    sel = [zkb.Polynomial(sel_np) for _ in range(qc)]
    vals = zkb.Polynomial(np.zeros((rows, 4), dtype=np.uint64))
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
    env.barrier()
    l0 = zkb.launch_count()
    zkb.prof.enable(True)
    zkb.prof.reset()
    for _ in range(args.steps):
        gr.evaluate(vals, fixed=sel, advice=adv, y=y_np, rot_scale=rs)
    env.barrier()
    qms, qk = zkb.prof.get("graph_evaluate")   # CUDA events around the kernel on the library stream
    qms /= max(qk, 1)
    zkb.prof.enable(False)
    env.launches += zkb.launch_count() - l0
    info = gr.last_info()
    qgbs = info["bytes_per_row"] * rows / (qms * 1e-3) / 1e9
    traffic, tsrc = measured_traffic("graph_evaluate_2^%dx%d" % (QUOT_LOG_N, QUOT_COLS))
    obj = {"workload": "evaluate_h custom gates: halo2-base gate q(a+bc-d) on %d advice columns, 2^%d extended rows, rot_scale %d, resident in HBM" % (qc, QUOT_LOG_N, rs),
           "value": env.world * rows / (qms * 1e-3), "unit": "rows/s", "ms_per_step": qms, "parity_checked": bool(q_ok),
           "lowered": info, "modmul_per_row": 3 * qc,
           "roofline": {"bound": "hbm", "achieved": qgbs, "peak": peak, "unit": "GB/s", "frac": qgbs / peak,
                        "traffic": traffic, "traffic_source": tsrc,
                        "note": "32 B x (polynomials read + previous value + result) per row; integer-issue bound, see DESIGN.md"}}
    for p_ in adv + sel + [vals]:
        p_.free()
    return obj, gr, y_np


# ---- N > 1: one 2^26 NTT sharded over the ranks (exchange fused into the NTT passes over NVLink peer memory) --------------------
def sharded_ntt_object(env: Env):
    zkb, lib, torch, zdist = env.zkb, env.lib, env.torch, env.zdist
    args = env.args
    peak, _ = hbm_peak()
    k = SHARDED_LOG_N
    rank, world = env.rank, env.world
    sh = zdist.ShardedNtt(k, device=env.dev)
    off, ln = zdist.ntt_slice(k, rank, world)
    w = zkb.omega(k)
    wp = w.ctypes.data_as(u64p)
    N = 1 << k
    # (1) the measured size, by the definition: a 2-sparse input (one non-zero in the first rank's slice, one in the last rank's)
    #     must give out[i] = v1 w^(i j1) + v2 w^(i j2) at sampled positions of EVERY rank's output slice
    v = random_field(3, 0xD157)
    j1, j2 = 0x2345677 % ln, N - 11
    d_in = torch.zeros(ln * 4, dtype=torch.int64, device=env.dev)
    if off <= j1 < off + ln:
        d_in.view(-1, 4)[j1 - off] = torch.from_numpy(v[1].view(np.int64)).to(env.dev)
    if off <= j2 < off + ln:
        d_in.view(-1, 4)[j2 - off] = torch.from_numpy(v[2].view(np.int64)).to(env.dev)
    d_out = torch.empty_like(d_in)
    rc = lib.zkb_dist_ntt_fr_dev(env.dptr(d_in), env.dptr(d_out), wp, k, env.sptr)
    if rc != 0 or lib.zkb_dist_status(env.sptr) != 0:
        raise RuntimeError(lib.zkb_last_error().decode())
    rng = np.random.default_rng(77 + rank)
    idx = [off, off + 1, off + ln - 1, off + ln // 2 + 3] + [off + int(x) for x in rng.integers(0, ln, 12)]
    got = d_out.view(-1, 4)
    rows = {i: got[i - off].cpu().numpy().view(np.uint64) for i in idx}
    sparse_ok = all(two_sparse_ok([rows[i]], v[1], v[2], j1, j2, k, [i], base=i) for i in idx)
    # (2) a random vector at 2^22: this rank's output slice against the single-GPU zkb_ntt_fr_dev of the whole vector, all elements
    k2 = min(22, k)
    N2 = 1 << k2
    full = env.device_field(N2, 0xD1570)        # same vector on every rank
    off2, ln2 = zdist.ntt_slice(k2, rank, world)
    w2 = zkb.omega(k2)
    w2p = w2.ctypes.data_as(u64p)
    d_in2 = full[off2:off2 + ln2].contiguous()
    d_out2 = torch.empty_like(d_in2)
    rc = lib.zkb_dist_ntt_fr_dev(env.dptr(d_in2), env.dptr(d_out2), w2p, k2, env.sptr)
    if rc != 0 or lib.zkb_dist_status(env.sptr) != 0:
        raise RuntimeError(lib.zkb_last_error().decode())
    scr = torch.empty_like(full)
    env.check(lib.zkb_ntt_fr_dev(env.dptr(full), env.dptr(scr), 1, w2p, k2, env.sptr))
    torch.cuda.synchronize()
    slice_ok = bool(torch.equal(d_out2, full[off2:off2 + ln2]))
    del full, scr, d_in2, d_out2
    # timing at the measured size
    d_in = env.device_field(ln, 0xD157 + rank).view(-1)
    lib.zkb_dist_ntt_fr_dev(env.dptr(d_in), None, wp, k, env.sptr)  # loads the symmetric input slice
    for _ in range(args.warmup):
        lib.zkb_dist_ntt_fr_dev(None, None, wp, k, env.sptr)
    sms = env.timed_events(lambda: lib.zkb_dist_ntt_fr_dev(None, None, wp, k, env.sptr), args.steps)
    if lib.zkb_dist_status(env.sptr) != 0:
        raise RuntimeError(lib.zkb_last_error().decode())
    ok = env.max_over_ranks(0.0 if (sparse_ok and slice_ok) else 1.0)[0] == 0.0
    obj = {"workload": "one best_fft of 2^%d sharded over %d GPUs (contiguous slices in, contiguous slices out)" % (k, world),
           "value": N / (sms * 1e-3), "unit": "elems/s", "ms_per_step": sms, "parity_checked": ok,
           "parity": "2-sparse input against the definition at 16 positions of every rank's output slice (2^%d); random vector at 2^%d: "
                     "every rank's whole output slice equals the single-GPU transform" % (k, k2),
           "exchange": "fused into NTT pass 0 (peer loads+stores) and the final pass (peer stores) over NVLink; device-side barriers",
           "nvlink_bytes_per_gpu_per_step": int(3 * (world - 1) / world * ln * 32),
           "roofline": {"bound": "hbm", "achieved": 64.0 * N / (sms * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                        "frac": 64.0 * N / (sms * 1e-3) / 1e9 / (peak * world)}}
    sh.close()
    del d_in, d_out
    torch.cuda.empty_cache()
    return obj


def sharded_quotient_object(env: Env, gr, y_np):
    torch, zdist = env.torch, env.zdist
    args = env.args
    try:
        srows = (1 << QUOT_LOG_N) // env.world
        sq = zdist.ShardedQuotient(gr, 4)
        base = env.device_field(srows, 0x51AB)   # the same rows on every rank: the domain is periodic with period srows, so the
        bufs = []                                # sharded result must equal a wrapping evaluation of one period
        for _ in range(2 * QUOT_COLS):
            buf, view = sq.alloc_column(srows, env.dev)
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
        qsms = env.timed_events(lambda: sq.run_padded(vals_s, sel_b, adv_b, [], y=y_np), args.steps)
        sq_ok = env.max_over_ranks(0.0 if sq_ok else 1.0)[0] == 0.0
        obj = {"workload": "the quotient workload above, 2^%d rows sharded by rows over %d GPUs" % (QUOT_LOG_N, env.world),
               "value": (1 << QUOT_LOG_N) / (qsms * 1e-3), "unit": "rows/s", "ms_per_step": qsms, "parity_checked": sq_ok,
               "exchange": "ring halo exchange, %d + %d rows per column over NCCL point-to-point, then zkb_graph_evaluate_dev on the row window" % (sq.halo_lo, sq.halo_hi),
               "halo_bytes_per_rank": (sq.halo_lo + sq.halo_hi) * 32 * 2 * QUOT_COLS}
        del bufs, vals_s, ref_s, plain, base
        torch.cuda.empty_cache()
        return obj
    except Exception as exc:  # a secondary object must not take the headline line down
        return {"error": repr(exc)[:300], "parity_checked": False}


# ---- one MSM of 2^26 points split over the N GPUs (strong scaling) ------------------------------------------------------------------
SPLIT_SEED = 0x26A0


def split_shard(env: Env, shard, parts, dev=None):
    """shard `shard` of `parts` of the 2^SPLIT_LOG_N-point workload: (scalars, discrete logs) on the device, seeded per shard so
    that any process / device regenerates the same values"""
    n = (1 << SPLIT_LOG_N) // parts
    torch = env.torch
    d = dev if dev is not None else env.dev
    g = torch.Generator(device=d)
    g.manual_seed(SPLIT_SEED + 2 * shard + 64 * parts)
    s = torch.randint(0, 1 << 60, (n, 4), dtype=torch.int64, device=d, generator=g)
    g.manual_seed(SPLIT_SEED + 2 * shard + 1 + 64 * parts)
    b = torch.randint(0, 1 << 60, (n, 4), dtype=torch.int64, device=d, generator=g)
    return s, b


def msm_split_object(env: Env):
    """rank r commits shard r (2^26 / N points) against its SRS range; all_gather + host fold.  Returns the object and the folded
    point (compared bit for bit with the single-process run of the same workload)."""
    zkb, lib, torch = env.zkb, env.lib, env.torch
    args = env.args
    n = (1 << SPLIT_LOG_N) // env.world
    d_s, d_b = split_shard(env, env.rank, env.world)
    s_np, b_np = to_np(d_s), to_np(d_b)
    bases = zkb.g1_fixed_base_mul(b_np)
    h = ctypes.c_uint64(0)
    env.check(lib.zkb_srs_register(bases.ctypes.data_as(u64p), n, ctypes.byref(h)))
    del bases
    lib.zkb_srs_precompute(h, None, None)
    out = np.zeros(12, dtype=np.uint64)
    outp = out.ctypes.data_as(u64p)

    def step():
        env.check(lib.zkb_msm_g1_srs_dev(h, 0, env.dptr(d_s), n, outp, env.sptr))
        return env.fold(out)

    total = step()
    # parity: [sum over all shards of <s, b>] G; every rank contributes its inner product through the host group
    from oracle import coracle

    coracle.build()
    ip = coracle.fr_inner_product(s_np, b_np)
    if env.world > 1:
        t = torch.from_numpy(ip.view(np.int64).copy())
        parts = [torch.empty_like(t) for _ in range(env.world)]
        env.dist.all_gather(parts, t, group=env.host_group)
        ips = np.stack([p.numpy().view(np.uint64) for p in parts])
        ip = ips[0]
        for x in ips[1:]:
            ip = coracle.vec_op("fr", "add", ip.reshape(1, 4), x.reshape(1, 4))[0]
    want = coracle.g1_mul(coracle.g1_generator(), ip)
    ok = bool((total[:8] == want).all())
    step()
    ms = env.timed_events(step, max(2, min(args.steps, 5)))
    lib.zkb_srs_release(h)
    del d_s, d_b
    torch.cuda.empty_cache()
    obj = {"workload": "one MSM of 2^%d uniform points split over %d GPU(s) by SRS point range, scalars resident in HBM, host fold per commit"
                       % (SPLIT_LOG_N, env.world),
           "scaling": "strong", "value": (1 << SPLIT_LOG_N) / (ms * 1e-3), "unit": "pts/s", "ms_per_step": ms, "parity_checked": ok}
    return obj, total


# ---- prover op-sequence replays (BASELINE configs #1-#3: "proof-gen sec") ------------------------------------------------------------
def wrapper_replay_object(env: Env):
    """Wrapper circuit, k = 22 (create_proof at /root/reference/aggregator/src/wrapper.rs:129-137 as driven by gen_recursion_snark,
    wrapper.rs:869-902): 22 commitments of 2^22 scalars, 13 lagrange_to_coeff, 16 coeff_to_extended (2^22 -> 2^24),
    one 2^24 inverse transform for h(X) — in prover order, every commitment folded on the host BEFORE the next op (the transcript
    needs it).  Multi-process sharding: every MSM by SRS point range, independent columns round-robin, the 2^24 transform through
    the sharded NTT.  STRONG scaling: the work is fixed, N GPUs share it."""
    zkb, lib, torch, zd = env.zkb, env.lib, env.torch, env.zdist
    rank, world = env.rank, env.world
    k, ek = WRAPPER_K, WRAPPER_K + 2
    n, N = 1 << k, 1 << ek
    n_msm, n_intt, n_c2e = 22, 13, 16
    off, ln = zd.point_range(n, rank, world)
    dlog = random_field(n, 0xB45E)                      # same on every rank
    bases = zkb.g1_fixed_base_mul(dlog[off:off + ln])   # this rank's SRS range
    h = ctypes.c_uint64(0)
    env.check(lib.zkb_srs_register(bases.ctypes.data_as(u64p), ln, ctypes.byref(h)))
    lib.zkb_srs_precompute(h, None, None)
    cols = random_field(4 * n, 0x22).reshape(4, n, 4)   # four distinct columns reused round-robin
    d_slices = [torch.from_numpy(np.ascontiguousarray(cols[c, off:off + ln]).view(np.int64)).to(env.dev) for c in range(4)]
    out = np.zeros(12, dtype=np.uint64)
    outp = out.ctypes.data_as(u64p)
    mine_intt = zd.columns_for_rank(n_intt, rank, world)
    mine_c2e = zd.columns_for_rank(n_c2e, rank, world)
    d_col = torch.from_numpy(np.ascontiguousarray(cols[0]).view(np.int64)).reshape(-1).to(env.dev)
    d_work = torch.empty(n * 4, dtype=torch.int64, device=env.dev)
    d_ext = torch.empty(N * 4, dtype=torch.int64, device=env.dev)
    d_scr = torch.empty(N * 4, dtype=torch.int64, device=env.dev)
    sharded = world > 1 and (world & (world - 1)) == 0 and world <= 8
    sh = zd.ShardedNtt(ek, device=env.dev) if sharded else None
    w_inv = zkb.EvaluationDomain(4, k).get_extended_omega()   # any 2^24-th root times the transform the same
    wp = w_inv.ctypes.data_as(u64p)
    if sh is not None:
        d_in = env.device_field(N // world, 7 + rank).view(-1)
        env.check(lib.zkb_dist_ntt_fr_dev(env.dptr(d_in), None, wp, ek, env.sptr))
        del d_in

    def replay():
        first = None
        for i in range(n_msm):
            env.check(lib.zkb_msm_g1_srs_dev(h, 0, env.dptr(d_slices[i % 4]), ln, outp, env.sptr))
            c = env.fold(out)
            if first is None:
                first = c
        for _ in mine_intt:
            d_work.copy_(d_col)
            env.check(lib.zkb_lagrange_to_coeff_dev(env.dptr(d_work), env.dptr(d_scr), 1, k, env.sptr))
        for _ in mine_c2e:
            env.check(lib.zkb_coeff_to_extended_dev(env.dptr(d_col), env.dptr(d_ext), env.dptr(d_scr), 1, k, ek, env.sptr))
        if sh is not None:
            env.check(lib.zkb_dist_ntt_fr_dev(None, None, wp, ek, env.sptr))
            env.check(lib.zkb_dist_status(env.sptr))
        else:
            env.check(lib.zkb_extended_to_coeff_dev(env.dptr(d_ext), env.dptr(d_scr), 1, k, ek, env.sptr))
        return first

    first = replay()
    ok = bool((first[:8] == known_dlog_point(np.ascontiguousarray(cols[0]), dlog)).all())
    # one coset of the replay against Horner at sampled points of the extended domain (rank 0's last coeff_to_extended output)
    if mine_c2e:
        from oracle import coracle
        from oracle import pyref as R
        from util import ints_to_limbs

        env.check(lib.zkb_coeff_to_extended_dev(env.dptr(d_col), env.dptr(d_ext), env.dptr(d_scr), 1, k, ek, env.sptr))
        torch.cuda.synchronize()
        wext = R.omega_for(ek)
        for i in (0, 5, N // 2 + 1, N - 1):
            x = R.FR_ZETA * pow(wext, i, R.FR) % R.FR
            wantv = coracle.fr_eval_polynomial(np.ascontiguousarray(cols[0]), ints_to_limbs([R.to_mont(x, R.FR)])[0])
            ok = ok and bool((d_ext.view(-1, 4)[i].cpu().numpy().view(np.uint64) == wantv).all())
    ok = env.max_over_ranks(0.0 if ok else 1.0)[0] == 0.0
    best = min(env.timed_events(replay, 1) for _ in range(3))
    lib.zkb_srs_release(h)
    if sh is not None:
        sh.close()
    del d_slices, d_col, d_work, d_ext, d_scr
    torch.cuda.empty_cache()
    return {"workload": "wrapper circuit k = %d prover op sequence: %d MSM 2^%d + %d iNTT 2^%d + %d coset NTT 2^%d->2^%d + 1 iNTT 2^%d, prover order, "
                        "host fold after every commitment" % (k, n_msm, k, n_intt, k, n_c2e, k, ek, ek),
            "mode": "kernel (operands resident in HBM)", "scaling": "strong", "n_gpus": world,
            "proof_gen_hot_path_ms": best, "value": 1e3 / best, "unit": "proofs/s (hot path only)", "parity_checked": ok,
            "parity": "first commitment == [<s, b>]G (oracle); one coset == Horner at 4 points of the extended domain",
            "sharding": "MSM by SRS point range + all_gather/fold per commitment; columns round-robin; the 2^%d transform through the sharded NTT" % ek}


def dropin_replays(env: Env, pinned: bool):
    """The same op sequences through the host-buffer C ABI (what the patched halo2 calls): PCIe and staging included.  With
    several devices bound in this process the library shards every call itself.  Returns {wrapper, voter, st} objects."""
    zkb, lib, torch = env.zkb, env.lib, env.torch

    def host(a):
        if not pinned:
            return a, a.ctypes.data
        t = torch.from_numpy(a.view(np.int64)).pin_memory()
        return t, t.data_ptr()

    res = {}
    mem = "page-locked" if pinned else "pageable"
    # ---- wrapper k = 22
    k, ek = WRAPPER_K, WRAPPER_K + 2
    n, N = 1 << k, 1 << ek
    n_msm, n_intt, n_c2e = 22, 13, 16
    dlog = random_field(n, 0xB45E)
    bases = zkb.g1_fixed_base_mul(dlog)
    params = zkb.ParamsKZG(k, bases)
    lib.zkb_srs_precompute(params.handle_g, None, None)
    cols_np = random_field(4 * n, 0x22)
    hold, base = host(cols_np)
    work = [host(np.empty((n, 4), dtype=np.uint64)) for _ in range(4)]      # the prover's own column Vecs (mutated in place)
    exts = [host(np.empty((N, 4), dtype=np.uint64)) for _ in range(4)]
    out = np.zeros(12, dtype=np.uint64)
    outp = out.ctypes.data_as(u64p)
    in4 = (u64p * 4)(*[ctypes.cast(base + c * n * 32, u64p) for c in range(4)])
    work4 = (u64p * 4)(*[ctypes.cast(w[1], u64p) for w in work])
    ext4 = (u64p * 4)(*[ctypes.cast(e[1], u64p) for e in exts])

    def refill():
        for c in range(4):
            ctypes.memmove(work[c][1], base + c * n * 32, n * 32)

    def wrapper():
        first = None
        for i in range(n_msm):
            env.check(lib.zkb_msm_g1_srs(params.handle_g, ctypes.cast(base + (i % 4) * n * 32, u64p), n, outp))
            if first is None:
                first = out.copy()
        for i in range(0, n_intt, 4):
            env.check(lib.zkb_lagrange_to_coeff_batch(work4, min(4, n_intt - i), k))
        for i in range(0, n_c2e, 4):
            env.check(lib.zkb_coeff_to_extended_batch(in4, ext4, 4, k, ek))
        env.check(lib.zkb_extended_to_coeff(ctypes.cast(exts[0][1], u64p), k, ek))
        return first

    refill()
    first = wrapper()
    ok = bool((first[:8] == known_dlog_point(cols_np[:n], dlog)).all())
    best = 1e30
    for _ in range(2):
        refill()
        t0 = time.perf_counter()
        wrapper()
        best = min(best, time.perf_counter() - t0)
    res["wrapper"] = {"workload": "wrapper k = %d op sequence through the host-buffer C ABI (%s operands)" % (k, mem),
                      "proof_gen_hot_path_ms": best * 1e3, "parity_checked": ok,
                      "h2d_bytes": n_msm * n * 32 + n_intt * n * 32 + n_c2e * n * 32 + N * 32,
                      "d2h_bytes": n_intt * n * 32 + n_c2e * N * 32 + N * 32}
    params.close()
    del hold, work, exts, bases
    # ---- voter (k = 13, 256 advice columns) and state_transition (k = 15, 8 columns): batched column calls
    for name, kk, ncols in (("voter", 13, 256), ("st", 15, 8)):
        nn, NN = 1 << kk, 1 << (kk + 2)
        dl = random_field(nn, 0x700 + kk)
        g = zkb.g1_fixed_base_mul(dl)
        p = zkb.ParamsKZG(kk, g, g)
        lib.zkb_srs_precompute(p.handle_g, None, None)
        lib.zkb_srs_precompute(p.handle_g_lagrange, None, None)
        c_np = random_field(nn * ncols, 77 + kk)
        hc, cb = host(c_np)
        he, eb = host(np.empty((NN * ncols, 4), dtype=np.uint64))
        ptrs = (u64p * ncols)(*[ctypes.cast(cb + i * nn * 32, u64p) for i in range(ncols)])
        eptrs = (u64p * ncols)(*[ctypes.cast(eb + i * NN * 32, u64p) for i in range(ncols)])
        outs = np.zeros((ncols, 12), dtype=np.uint64)
        saved = c_np.copy()

        def shape():
            env.check(lib.zkb_msm_g1_srs_batch(p.handle_g_lagrange, ptrs, ncols, nn, outs.ctypes.data_as(u64p)))   # advice commits
            env.check(lib.zkb_lagrange_to_coeff_batch(ptrs, ncols, kk))
            env.check(lib.zkb_coeff_to_extended_batch(ptrs, eptrs, ncols, kk, kk + 2))
            env.check(lib.zkb_extended_to_coeff(ctypes.cast(eb, u64p), kk, kk + 2))                                  # h(X)
            for i in range(8):                                                                                        # h pieces + openings
                env.check(lib.zkb_msm_g1_srs(p.handle_g, ctypes.cast(cb + (i % ncols) * nn * 32, u64p), nn, outp))

        def reset():
            ctypes.memmove(cb, saved.ctypes.data, saved.nbytes)

        reset()
        shape()
        from oracle import coracle

        coracle.build()
        okc = bool((outs[0] == coracle.best_multiexp(saved[:nn], g)).all()) and bool((outs[ncols - 1] == coracle.best_multiexp(saved[(ncols - 1) * nn:], g)).all())
        bestc = 1e30
        for _ in range(3):
            reset()
            t0 = time.perf_counter()
            shape()
            bestc = min(bestc, time.perf_counter() - t0)
        res[name] = {"workload": "%s circuit shape: k = %d, %d columns: commit_lagrange batch + lagrange_to_coeff batch + coeff_to_extended batch + "
                                 "extended_to_coeff + 8 commits, host-buffer C ABI (%s operands)" % (name, kk, ncols, mem),
                     "proof_gen_hot_path_ms": bestc * 1e3, "parity_checked": okc}
        p.close()
        del hc, he
    return res


# ---- BASELINE configs #4 / #5: the MSM and NTT sweeps through the host-buffer C ABI ---------------------------------------------------
def sweep_object(env: Env):
    """Standalone G1 MSM 2^16..2^24 (2^26: msm_split) and Fr NTT 2^18..2^26 x {1, 16} columns (capped at 2 GiB per call), each as ONE
    host-buffer call from page-locked memory (PCIe included).  With several devices bound in this process (single_process leg)
    the library shards every call itself, so the same rows at N = 1, 2, 4, 8 are the strong-scaling curve of the drop-in."""
    zkb, lib, torch = env.zkb, env.lib, env.torch
    rows = []
    kmax = min(LOG_N_MSM, 24)
    dl = random_field(1 << kmax, 0x5EE9)
    bases = zkb.g1_fixed_base_mul(dl)
    sc = random_field(1 << kmax, 0x5EEA)
    h_s = torch.from_numpy(sc.view(np.int64)).pin_memory()
    out = np.zeros(12, dtype=np.uint64)
    outp = out.ctypes.data_as(u64p)
    for k in range(16, kmax + 1, 2):
        n = 1 << k
        h = ctypes.c_uint64(0)
        env.check(lib.zkb_srs_register(bases.ctypes.data_as(u64p), n, ctypes.byref(h)))
        lib.zkb_srs_precompute(h, None, None)
        sp = ctypes.cast(h_s.data_ptr(), u64p)
        env.check(lib.zkb_msm_g1_srs(h, sp, n, outp))
        ok = bool((out[:8] == known_dlog_point(sc[:n], dl[:n])).all())
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            env.check(lib.zkb_msm_g1_srs(h, sp, n, outp))
            best = min(best, time.perf_counter() - t0)
        rows.append({"op": "msm", "log_n": k, "e2e_ms": best * 1e3, "pts_per_s": n / best, "parity": ok})
        lib.zkb_srs_release(h)
    del bases, h_s
    cap = 1 << 26    # elements per call (2 GiB)
    buf = torch.empty((cap, 4), dtype=torch.int64).pin_memory()
    rnd = env.device_field(1 << 24, 0x5EEB).cpu()
    for k in (18, 20, 22, 24, 26):
        N = 1 << k
        w = zkb.omega(k)
        wp = w.ctypes.data_as(u64p)
        for cols in (1, 16):
            if N * cols > cap:
                cols = cap // N
                if cols <= 1 and any(r["op"] == "ntt" and r["log_n"] == k and r["cols"] == 1 for r in rows):
                    continue
            ptrs = (u64p * cols)(*[ctypes.cast(buf.data_ptr() + i * N * 32, u64p) for i in range(cols)])
            # parity: a 2-sparse column 0 against the definition
            buf[:N * cols].zero_()
            j1, j2 = 12345 % N, N - 7
            v = random_field(3, 0x5EEC + k)
            buf[j1] = torch.from_numpy(v[1].view(np.int64))
            buf[j2] = torch.from_numpy(v[2].view(np.int64))
            env.check(lib.zkb_ntt_fr_batch(ptrs, cols, wp, k))
            got = buf[:N].numpy().view(np.uint64).reshape(-1, 4)
            ok = two_sparse_ok(got, v[1], v[2], j1, j2, k, [0, 1, N // 2 + 3, N - 1, N // 3])
            for off in range(0, N * cols, 1 << 24):
                m = min(1 << 24, N * cols - off)
                buf[off:off + m] = rnd[:m]
            best = 1e30
            for _ in range(3):
                t0 = time.perf_counter()
                env.check(lib.zkb_ntt_fr_batch(ptrs, cols, wp, k))
                best = min(best, time.perf_counter() - t0)
            rows.append({"op": "ntt", "log_n": k, "cols": cols, "e2e_ms": best * 1e3, "elems_per_s": N * cols / best, "parity": ok})
    del buf
    return {"workload": "BASELINE configs #4 / #5 through the host-buffer C ABI, page-locked operands, devices bound in this process: %d"
                        % max(1, len(zkb.bound_devices())),
            "rows": rows, "parity_checked": all(r["parity"] for r in rows)}


def sp_wrapper_kernel_replay(env: Env, world: int):
    """wrapper_replay_object's sequence with operands resident in HBM, driven by ONE process: one host thread per device (bound with
    zkb_thread_bind_device) plays the part a torchrun rank plays — its SRS point range of every commit, its round-robin columns, its
    slice of the sharded 2^24 transform — and the main thread folds the partial sums (no NCCL, no torch.distributed)."""
    zkb, lib, torch, zd = env.zkb, env.lib, env.torch, env.zdist
    k, ek = WRAPPER_K, WRAPPER_K + 2
    n, N = 1 << k, 1 << ek
    n_msm, n_intt, n_c2e = 22, 13, 16
    dlog = random_field(n, 0xB45E)
    bases = zkb.g1_fixed_base_mul(dlog)
    h = ctypes.c_uint64(0)
    env.check(lib.zkb_srs_register(bases.ctypes.data_as(u64p), n, ctypes.byref(h)))     # replicated to every device
    lib.zkb_srs_precompute(h, None, None)
    cols = random_field(4 * n, 0x22).reshape(4, n, 4)
    env.check(lib.zkb_dist_create_inprocess(ek))
    w_inv = zkb.EvaluationDomain(4, k).get_extended_omega()
    wp = w_inv.ctypes.data_as(u64p)
    st = []
    for r in range(world):
        dev = torch.device("cuda", r)
        off, ln = zd.point_range(n, r, world)
        g = torch.Generator(device=dev)
        g.manual_seed(7 + r)
        st.append({
            "off": off, "ln": ln, "stream": torch.cuda.Stream(device=dev),
            "slices": [torch.from_numpy(np.ascontiguousarray(cols[c, off:off + ln]).view(np.int64)).to(dev) for c in range(4)],
            "col": torch.from_numpy(np.ascontiguousarray(cols[0]).view(np.int64)).reshape(-1).to(dev),
            "work": torch.empty(n * 4, dtype=torch.int64, device=dev), "ext": torch.empty(N * 4, dtype=torch.int64, device=dev),
            "scr": torch.empty(N * 4, dtype=torch.int64, device=dev),
            "din": torch.randint(0, 1 << 60, ((N // world) * 4,), dtype=torch.int64, device=dev, generator=g),
            "dout": torch.empty((N // world) * 4, dtype=torch.int64, device=dev),
            "out": np.zeros(12, dtype=np.uint64), "intt": zd.columns_for_rank(n_intt, r, world), "c2e": zd.columns_for_rank(n_c2e, r, world)})
    pool = ThreadPoolExecutor(world)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

    def chk(rc):
        if rc != 0:
            raise RuntimeError(lib.zkb_last_error().decode())

    def commit(args_):
        r, i = args_
        s_ = st[r]
        chk(lib.zkb_thread_bind_device(r))
        chk(lib.zkb_msm_g1_srs_dev(h, s_["off"], vp(s_["slices"][i % 4]), s_["ln"], s_["out"].ctypes.data_as(u64p), ctypes.c_void_p(s_["stream"].cuda_stream)))

    def tail(r):
        s_ = st[r]
        sp = ctypes.c_void_p(s_["stream"].cuda_stream)
        chk(lib.zkb_thread_bind_device(r))
        with torch.cuda.stream(s_["stream"]):
            for _ in s_["intt"]:
                s_["work"].copy_(s_["col"])
                chk(lib.zkb_lagrange_to_coeff_dev(vp(s_["work"]), vp(s_["scr"]), 1, k, sp))
            for _ in s_["c2e"]:
                chk(lib.zkb_coeff_to_extended_dev(vp(s_["col"]), vp(s_["ext"]), vp(s_["scr"]), 1, k, ek, sp))
        chk(lib.zkb_dist_ntt_fr_dev(vp(s_["din"]), vp(s_["dout"]), wp, ek, sp))
        chk(lib.zkb_dist_status(sp))     # synchronises this device's stream

    def replay():
        first = None
        for i in range(n_msm):
            list(pool.map(commit, [(r, i) for r in range(world)]))
            c = zkb.g1_sum(np.stack([s_["out"] for s_ in st]))
            if first is None:
                first = c
        list(pool.map(tail, range(world)))
        return first

    first = replay()
    ok = bool((first[:8] == known_dlog_point(np.ascontiguousarray(cols[0]), dlog)).all())
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        replay()
        best = min(best, time.perf_counter() - t0)
    pool.shutdown()
    lib.zkb_dist_destroy()
    lib.zkb_srs_release(h)
    del st
    torch.cuda.empty_cache()
    return {"workload": "the wrapper_replay kernel sequence (operands resident in HBM) driven by ONE process, one host thread per device",
            "proof_gen_hot_path_ms": best * 1e3, "parity_checked": ok, "timer": "wall clock (every device's stream synchronised at the end)"}


# ---- N > 1: rank 0 alone drives all N devices through the unchanged C ABI --------------------------------------------------------------
def single_process_object(env: Env, mp_split, mp_split_point, mp_wrapper=None):
    """zkb_init(all N devices) in ONE process — the deployment of the reference's prover (create_proof is one process,
    /root/reference/aggregator/src/wrapper.rs:129-137).  Must reproduce the torchrun results: the 2^26 MSM split N ways bit for
    bit and within a few percent of the multi-process time; drop-in replays are reported beside the N = 1 ones."""
    zkb, lib, torch = env.zkb, env.lib, env.torch
    world = env.world
    obj = {"devices": world}
    zkb.shutdown()
    if env.numa and env.numa.get("bound"):   # one process drives every device now: back to all the host's cores
        os.sched_setaffinity(0, set(env.numa["all"]))
    zkb.init(list(range(world)))
    try:
        # ---- the 2^26 MSM of msm_split, sharded by the library: same per-shard seeds, so the same points and scalars
        n = 1 << SPLIT_LOG_N
        per = n // world
        h_s = torch.empty((n, 4), dtype=torch.int64).pin_memory()
        b_np = np.empty((n, 4), dtype=np.uint64)
        for r in range(world):
            s, b = split_shard(env, r, world, dev=torch.device("cuda", 0))
            h_s[r * per:(r + 1) * per].copy_(s)
            b_np[r * per:(r + 1) * per] = to_np(b)
            del s, b
        torch.cuda.synchronize()
        bases = zkb.g1_fixed_base_mul(b_np)
        h = ctypes.c_uint64(0)
        env.check(lib.zkb_srs_register(bases.ctypes.data_as(u64p), n, ctypes.byref(h)))   # replicated to every device
        del bases, b_np
        lib.zkb_srs_precompute(h, None, None)
        out = np.zeros(12, dtype=np.uint64)
        outp = out.ctypes.data_as(u64p)
        sp = ctypes.cast(h_s.data_ptr(), u64p)
        env.check(lib.zkb_msm_g1_srs(h, sp, n, outp))
        identical = bool((out == mp_split_point).all())
        # device-resident, one host thread per device (what a torchrun rank does, without processes or NCCL): shard r on device r
        d_sh = [h_s[r * per:(r + 1) * per].to(torch.device("cuda", r)) for r in range(world)]
        outs = [np.zeros(12, dtype=np.uint64) for _ in range(world)]
        streams = [torch.cuda.Stream(device=r) for r in range(world)]
        pool = ThreadPoolExecutor(world)

        def one(r):
            rc = lib.zkb_msm_g1_srs_dev(h, r * per, ctypes.c_void_p(d_sh[r].data_ptr()), per, outs[r].ctypes.data_as(u64p),
                                        ctypes.c_void_p(streams[r].cuda_stream))
            if rc != 0:
                raise RuntimeError(lib.zkb_last_error().decode())

        def resident():
            list(pool.map(one, range(world)))
            return zkb.g1_sum(np.stack(outs))

        tot = resident()
        identical = identical and bool((tot == mp_split_point).all())
        resident()
        for r in range(world):
            torch.cuda.synchronize(r)
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            resident()
        res_ms = (time.perf_counter() - t0) / reps * 1e3
        env.check(lib.zkb_msm_g1_srs(h, sp, n, outp))
        t0 = time.perf_counter()
        for _ in range(3):
            env.check(lib.zkb_msm_g1_srs(h, sp, n, outp))
        e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
        pool.shutdown()
        obj["msm_split"] = {"workload": "the msm_split workload (one 2^%d MSM over %d devices) driven by ONE process" % (SPLIT_LOG_N, world),
                            "resident_ms_per_step": res_ms, "multiprocess_ms_per_step": mp_split["ms_per_step"],
                            "resident_vs_multiprocess": res_ms / mp_split["ms_per_step"],
                            "timer": "wall clock (the call returns the folded result), one host thread per device",
                            "e2e_ms_per_step": e2e_ms, "e2e_value": n / (e2e_ms * 1e-3), "unit": "pts/s",
                            "e2e_note": "zkb_msm_g1_srs from page-locked host scalars: every device uploads its own share over its own PCIe link",
                            "bit_identical_to_multiprocess": identical}
        lib.zkb_srs_release(h)
        del d_sh, h_s
        torch.cuda.empty_cache()
        # ---- the resident wrapper replay, one host thread per device, against the torchrun run of the same sequence
        if world & (world - 1) == 0:
            wk = sp_wrapper_kernel_replay(env, world)
            if mp_wrapper and "proof_gen_hot_path_ms" in mp_wrapper:
                wk["multiprocess_ms"] = mp_wrapper["proof_gen_hot_path_ms"]
                wk["vs_multiprocess"] = wk["proof_gen_hot_path_ms"] / mp_wrapper["proof_gen_hot_path_ms"]
            obj["wrapper_replay_kernel"] = wk
            identical = identical and wk["parity_checked"]
        # ---- drop-in replays with all devices behind the same calls
        rep = dropin_replays(env, pinned=False)
        obj["wrapper_replay_dropin_pageable"] = rep["wrapper"]
        obj["voter_replay_dropin_pageable"] = rep["voter"]
        obj["st_replay_dropin_pageable"] = rep["st"]
        obj["sweep"] = sweep_object(env)
        obj["parity_checked"] = identical and all(v["parity_checked"] for v in rep.values()) and obj["sweep"]["parity_checked"]
        obj["launches"] = int(zkb.launch_count())
    finally:
        zkb.shutdown()
    return obj


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="zkb200")
    ap.add_argument("--skip-ntt", action="store_true", help="headline MSM only (skips every secondary object)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-replay", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    wd = int(os.environ.get("ZKB_BENCH_WATCHDOG", "0"))
    if wd:  # debugging aid: dump all Python stacks and exit if the run exceeds `wd` seconds
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)

    def note(msg):
        if os.environ.get("ZKB_BENCH_VERBOSE"):
            print("[bench %.1f] %s" % (time.time() - t_start, msg), file=sys.stderr, flush=True)

    t_start = time.time()
    env = Env(args)
    rank, world = env.rank, env.world
    clocks = ClockSampler(env.local_rank)
    clocks.start()
    note("init done")
    head = headline_msm(env, clocks)
    clock_info = clocks.stop()
    note("headline done")
    parity = {"msm_2^%d_known_dlog" % LOG_N_MSM: head["parity"]}
    peak = ctypes.c_double(0)
    env.lib.zkb_measure_imad_peak(ctypes.byref(peak))

    def guarded(name, fn):
        """a secondary object must not take the headline line down; its failure is reported and fails parity"""
        try:
            o = fn()
            note(name + " done")
            return o
        except Exception as exc:
            note(name + " FAILED: " + repr(exc))
            return {"error": repr(exc)[:400], "parity_checked": False}

    ntt_obj = quot_obj = sharded_obj = sq_obj = wrap_obj = split_obj = sp_obj = sweep_obj = None
    dropin = {}
    gr = y_np = None
    split_point = None
    if not args.skip_ntt:
        ntt_obj = guarded("ntt", lambda: ntt_object(env, peak.value))

        def _q():
            nonlocal gr, y_np
            o, gr, y_np = quotient_object(env)
            return o

        quot_obj = guarded("quotient", _q)
        if world > 1 and (world & (world - 1)) == 0 and world <= 8:
            sharded_obj = guarded("sharded_ntt", lambda: sharded_ntt_object(env))
            if gr is not None:
                sq_obj = sharded_quotient_object(env, gr, y_np)
    if not args.skip_ntt and not args.skip_replay:
        wrap_obj = guarded("wrapper_replay", lambda: wrapper_replay_object(env))

        def _s():
            nonlocal split_point
            o, split_point = msm_split_object(env)
            return o

        split_obj = guarded("msm_split", _s)
        if world == 1:   # drop-in (host-buffer) replays of the three circuits on one GPU; at N > 1 the single process runs them
            dropin = guarded("dropin_replays", lambda: {"pageable": dropin_replays(env, False), "page_locked": dropin_replays(env, True)})
            sweep_obj = guarded("sweep", lambda: sweep_object(env))
    # ---- N > 1: every rank releases its GPU, then rank 0 alone drives all N devices
    if world > 1 and not args.skip_ntt and not args.skip_replay and split_point is not None:
        env.torch.cuda.synchronize()
        env.zkb.shutdown()
        env.torch.cuda.empty_cache()
        env.dist.barrier(group=env.host_group)      # host-side: an NCCL barrier would keep the other GPUs spinning
        if rank == 0:
            sp_obj = guarded("single_process", lambda: single_process_object(env, split_obj, split_point, wrap_obj))
        env.dist.barrier(group=env.host_group)
        env.zkb.init(env.local_rank)
    for name, o in (("ntt", ntt_obj), ("quotient", quot_obj), ("sharded_ntt", sharded_obj), ("sharded_quotient", sq_obj),
                    ("wrapper_replay", wrap_obj), ("msm_split", split_obj), ("single_process", sp_obj), ("sweep", sweep_obj)):
        if o is not None:
            parity[name] = bool(o.get("parity_checked", False))
    if dropin:
        if "error" in dropin:
            parity["dropin_replays"] = False
        else:
            for memk, d in dropin.items():
                for nm, o in d.items():
                    parity["dropin_%s_%s" % (nm, memk)] = bool(o.get("parity_checked", False))
    # ---- CPU baseline on this box (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        v, cores, dt, native = cpu_baseline_msm(LOG_N_MSM)
        cpu = {"value": v, "unit": "pts/s", "cores": cores, "kind": "port",
               "sample": "best_multiexp restatement (C, pthreads; not rayon; %s) on the full 2^%d uniform points once, %.1f s"
                         % ("-O3 -march=native built on this host" if native else "portable -O3 build", LOG_N_MSM, dt)}
        note("cpu baseline done")
    all_ok = all(parity.values())
    if rank == 0:
        replay = None
        if wrap_obj is not None:
            replay = {"kernel": wrap_obj}
            if dropin and "error" not in dropin:
                replay["dropin_pageable"] = dropin["pageable"]["wrapper"]
                replay["dropin_page_locked"] = dropin["page_locked"]["wrapper"]
        line = {
            "metric": METRIC, "value": head["value"], "unit": "pts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (254-bit Montgomery integers)", "data": "synthetic",
            "config": head["config"], "e2e": head["e2e"], "e2e_pageable": head["e2e_pageable"],
            "no_table_ms_per_step": head["no_table_ms_per_step"], "witness_like": head["witness_like"],
            "gpu_launches": int(env.launches), "parity_checked": all_ok, "parity": parity, "roofline": head["roofline"],
            "cpu_baseline": cpu, "clocks": clock_info, "ntt": ntt_obj, "quotient": quot_obj,
            "sharded_ntt": sharded_obj, "sharded_quotient": sq_obj,
            "wrapper_replay": replay, "msm_split": split_obj,
            "voter_replay": ({"dropin_pageable": dropin["pageable"]["voter"], "dropin_page_locked": dropin["page_locked"]["voter"]}
                             if dropin and "error" not in dropin else None),
            "st_replay": ({"dropin_pageable": dropin["pageable"]["st"], "dropin_page_locked": dropin["page_locked"]["st"]}
                          if dropin and "error" not in dropin else None),
            "single_process": sp_obj, "sweep": sweep_obj,
            "bench_wall_s": time.time() - t_start,
        }
        print(json.dumps(line))
    if world > 1:
        env.dist.destroy_process_group()
    return 0 if all_ok else 1


if __name__ == "__main__":
    sys.exit(main())
