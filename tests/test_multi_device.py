"""One process driving several GPUs through the unchanged C ABI (zkb_init with a device list) — the deployment the
reference's single-process prover needs (create_proof, /root/reference/aggregator/src/wrapper.rs:129-137, chained by
gen_recursion_snark, wrapper.rs:869-902; BASELINE config #3 "MSMs and column NTTs sharded over 8xB200").

Every multi-device result must be bit-identical to the oracle (small sizes) and to the single-device result of the same
library (large sizes).  Needs >= 2 visible GPUs: on a 1-GPU box these tests skip and `bench.py --gpus N` (rank 0,
`single_process` object) runs the same checks instead.
"""
import ctypes
import importlib

import numpy as np
import pytest

from util import random_field

zkb = importlib.import_module("zksnap-circuits-halo2_b200")

pytestmark = pytest.mark.gpu


def _ndev():
    try:
        return zkb.device_count()
    except Exception:
        return 0


@pytest.fixture(scope="module")
def multi():
    """Rebinds the library to all visible devices (<= 8) for this module, with small sharding thresholds so that oracle-sized
    inputs take the multi-device paths; restores the single-device binding afterwards."""
    n = _ndev()
    if n < 2:
        pytest.skip("needs >= 2 GPUs in one process")
    zkb.shutdown()
    devs = list(range(min(n, 8)))
    zkb.init(devs)
    zkb.lib().zkb_multi_device_set(10, 10, 12)
    yield devs
    zkb.lib().zkb_multi_device_set(0, 0, 0)
    zkb.shutdown()
    zkb.init(0)


def _single(fn):
    """Runs fn() with the library bound to device 0 only and returns its result (the reference for bit-identity)."""
    devs = zkb.bound_devices()
    zkb.shutdown()
    zkb.init(0)
    try:
        return fn()
    finally:
        zkb.shutdown()
        zkb.init(devs)
        zkb.lib().zkb_multi_device_set(10, 10, 12)


def test_bound_devices(multi):
    assert zkb.bound_devices() == multi


@pytest.mark.parametrize("n", [(1 << 12) + 3, 1 << 14, (1 << 15) + 5])
def test_sharded_commit_vs_oracle(oracle, multi, n):
    s = random_field(n, 70 + n)
    bases = zkb.g1_fixed_base_mul(random_field(n, 71 + n))
    want = oracle.best_multiexp(s, bases)
    assert (zkb.best_multiexp(s, bases) == want).all()            # ad-hoc bases, sharded by point range
    for pre in (0, 1):                                              # resident SRS: plain bases, then the window tables
        zkb.lib().zkb_srs_set_precompute(pre)
        params = zkb.ParamsKZG(int(np.ceil(np.log2(n))), np.concatenate([bases, np.zeros(((1 << int(np.ceil(np.log2(n)))) - n, 8), np.uint64)]))
        assert (params.commit(s) == want).all()
        off = n // 3
        assert (params.commit_range(off, s[: n - off]) == oracle.best_multiexp(s[: n - off], bases[off:])).all()
        params.close()
    zkb.lib().zkb_srs_set_precompute(2)


def test_sharded_commit_edge_cases(oracle, multi):
    n = 1 << 12
    bases = zkb.g1_fixed_base_mul(random_field(n, 5))
    z = zkb.best_multiexp(np.zeros((n, 4), np.uint64), bases)      # every shard returns the identity
    assert not z[8:].any()
    s = random_field(n, 6)
    s[:] = s[0]                                                     # all-equal scalars
    assert (zkb.best_multiexp(s, bases) == oracle.best_multiexp(s, bases)).all()
    s = random_field(n, 7)
    s[: n // 2] = 0                                                 # one shard all zeros
    assert (zkb.best_multiexp(s, bases) == oracle.best_multiexp(s, bases)).all()


def test_batch_commit_split_by_column(oracle, multi):
    k, ncols = 12, 7
    n = 1 << k
    bases = zkb.g1_fixed_base_mul(random_field(n, 11))
    params = zkb.ParamsKZG(k, bases)
    cols = [random_field(n, 100 + i) for i in range(ncols)]
    got = params.commit_batch(cols)
    for i in range(ncols):
        assert (got[i] == oracle.best_multiexp(cols[i], bases)).all()
    params.close()


def test_batch_ntt_split_by_column(oracle, multi):
    k = 12
    d = zkb.EvaluationDomain(4, k)
    cols = [random_field(1 << k, 300 + i) for i in range(5)]
    for got, a in zip(d.lagrange_to_coeff_batch(cols), cols):
        assert (got == oracle.lagrange_to_coeff(a, k)).all()
    for got, a in zip(d.coeff_to_extended_batch(cols), cols):
        assert (got == oracle.coeff_to_extended(a, k, d.extended_k)).all()


@pytest.mark.parametrize("k", [12, 13, 16])
def test_single_transform_sharded_in_process(oracle, multi, k):
    a = random_field(1 << k, 400 + k)
    w = zkb.omega(k)
    f = a.copy()
    zkb.best_fft(f, w, k)
    assert (f == oracle.best_fft(a, w, k)).all()
    d = zkb.EvaluationDomain(4, k)
    assert (d.lagrange_to_coeff(a) == oracle.lagrange_to_coeff(a, k)).all()
    assert (d.coeff_to_lagrange(a) == oracle.coeff_to_lagrange(a, k)).all()
    if k <= 13:
        ext = random_field(1 << d.extended_k, 500 + k)
        got = d.extended_to_coeff(ext)
        assert (got == oracle.extended_to_coeff(ext, k, d.extended_k)[: d.n * d.quotient_poly_degree]).all()


def test_large_sizes_bit_identical_to_single_device(multi):
    """2^22-point commit and one 2^22 transform with the DEFAULT thresholds: multi-device == single-device, bit for bit."""
    k = 22
    n = 1 << k
    s = random_field(n, 900)
    dl = random_field(n, 901)
    a = random_field(n, 902)
    w = zkb.omega(k)

    def run():
        bases = zkb.g1_fixed_base_mul(dl)
        params = zkb.ParamsKZG(k, bases)
        c = params.commit(s)
        params.close()
        f = a.copy()
        zkb.best_fft(f, w, k)
        return c, f

    c1, f1 = _single(run)
    zkb.lib().zkb_multi_device_set(0, 0, 0)
    try:
        c2, f2 = run()
    finally:
        zkb.lib().zkb_multi_device_set(10, 10, 12)
    assert (c1 == c2).all()
    assert (f1 == f2).all()


def test_dev_calls_follow_the_pointer(multi):
    """*_dev entry points act on the device that owns the buffer, so one host thread per device can drive resident work."""
    torch = pytest.importorskip("torch")
    k = 14
    n = 1 << k
    a = random_field(n, 77)
    w = zkb.omega(k)
    want = a.copy()
    zkb.best_fft(want, w, k)
    lib = zkb.lib()
    for dev in multi[:2]:
        with torch.cuda.device(dev):
            d_a = torch.from_numpy(a.view(np.int64).copy()).to(f"cuda:{dev}")
            d_s = torch.empty_like(d_a)
            st = torch.cuda.current_stream()
            rc = lib.zkb_ntt_fr_dev(ctypes.c_void_p(d_a.data_ptr()), ctypes.c_void_p(d_s.data_ptr()), 1,
                                    w.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), k, ctypes.c_void_p(st.cuda_stream))
            assert rc == 0, lib.zkb_last_error()
            st.synchronize()
            assert (d_a.cpu().numpy().view(np.uint64).reshape(-1, 4) == want).all()


def test_thread_bound_to_a_device_keeps_resident_work_there(oracle, multi):
    """zkb_thread_bind_device: a host thread bound to device 1 creates polynomial handles there and runs the resident chain
    (commit against that device's SRS replica, lagrange_to_coeff, eval) with the same results as the home device; it never fans out."""
    import threading
    k = 12
    n = 1 << k
    bases = zkb.g1_fixed_base_mul(random_field(n, 21))
    params = zkb.ParamsKZG(k, bases)
    a = random_field(n, 22)
    x = random_field(1, 23)[0]
    d = zkb.EvaluationDomain(4, k)
    res = {}

    def chain(tag, device):
        try:
            if device is not None:
                assert zkb.lib().zkb_thread_bind_device(device) == 0
            p = zkb.Polynomial(a)
            c = p.commit(params, lagrange=False)
            p.lagrange_to_coeff(d)
            res[tag] = (c, p.eval(x), p.to_host())
            p.free()
        except Exception as exc:  # surfaced by the assertions below
            res[tag] = exc

    chain("home", None)
    t = threading.Thread(target=chain, args=("dev1", multi[1]))
    t.start()
    t.join()
    assert not isinstance(res["dev1"], Exception), res["dev1"]
    assert (res["home"][0] == oracle.best_multiexp(a, bases)).all()
    for u, v in zip(res["home"], res["dev1"]):
        assert (u == v).all()
    params.close()


def test_inprocess_sharded_ntt_on_device_slices(multi):
    """zkb_dist_create_inprocess + one host thread per device calling zkb_dist_ntt_fr_dev on resident slices == the single-device
    transform of the whole vector."""
    torch = pytest.importorskip("torch")
    from concurrent.futures import ThreadPoolExecutor
    lib = zkb.lib()
    world = 1
    while world * 2 <= len(multi):
        world *= 2
    k = 16
    n = 1 << k
    a = random_field(n, 31)
    w = zkb.omega(k)
    wp = w.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
    want = a.copy()
    zkb.lib().zkb_multi_device_set(10, 10, 28)      # keep the reference transform on one device
    zkb.best_fft(want, w, k)
    zkb.lib().zkb_multi_device_set(10, 10, 12)
    assert lib.zkb_dist_create_inprocess(k) == 0, lib.zkb_last_error()
    ln = n // world
    outs = [None] * world

    def rank(r):
        assert lib.zkb_thread_bind_device(multi[r]) == 0
        dev = torch.device("cuda", multi[r])
        d_in = torch.from_numpy(a[r * ln:(r + 1) * ln].view(np.int64).copy()).to(dev)
        d_out = torch.empty_like(d_in)
        st = torch.cuda.Stream(device=dev)
        sp = ctypes.c_void_p(st.cuda_stream)
        assert lib.zkb_dist_ntt_fr_dev(ctypes.c_void_p(d_in.data_ptr()), ctypes.c_void_p(d_out.data_ptr()), wp, k, sp) == 0, lib.zkb_last_error()
        assert lib.zkb_dist_status(sp) == 0, lib.zkb_last_error()
        outs[r] = d_out.cpu().numpy().view(np.uint64).reshape(-1, 4)

    with ThreadPoolExecutor(world) as pool:
        list(pool.map(rank, range(world)))
    lib.zkb_dist_destroy()
    assert (np.concatenate(outs) == want).all()


def test_sharded_ntt_peer_never_arrives_is_an_error_not_a_hang(multi):
    """One rank calls the collective, the other never does: the device-side barrier gives up after the configured bound, the
    remaining passes return without touching half-exchanged data, zkb_dist_status reports the failure ONCE, and after the contexts
    are rebuilt the same transform succeeds."""
    torch = pytest.importorskip("torch")
    import threading
    lib = zkb.lib()
    k = 14
    n = 1 << k
    a = random_field(n, 61)
    w = zkb.omega(k)
    wp = w.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
    world = 1
    while world * 2 <= len(multi):
        world *= 2
    ln = n // world
    lib.zkb_dist_set_timeout_ms(300)
    try:
        assert lib.zkb_dist_create_inprocess(k) == 0, lib.zkb_last_error()
        res = {}

        def lonely():
            assert lib.zkb_thread_bind_device(multi[0]) == 0
            dev = torch.device("cuda", multi[0])
            d_in = torch.from_numpy(a[:ln].view(np.int64).copy()).to(dev)
            d_out = torch.zeros_like(d_in)
            st = torch.cuda.Stream(device=dev)
            sp = ctypes.c_void_p(st.cuda_stream)
            assert lib.zkb_dist_ntt_fr_dev(ctypes.c_void_p(d_in.data_ptr()), ctypes.c_void_p(d_out.data_ptr()), wp, k, sp) == 0
            res["first"] = lib.zkb_dist_status(sp)
            res["msg"] = lib.zkb_last_error().decode()
            res["second"] = lib.zkb_dist_status(sp)       # reported once, then cleared

        t = threading.Thread(target=lonely)
        t.start()
        t.join(timeout=30)
        assert not t.is_alive(), "the lonely rank hung"
        assert res["first"] != 0 and "never arrived" in res["msg"]
        assert res["second"] == 0
        lib.zkb_dist_destroy()
        lib.zkb_dist_set_timeout_ms(0)
        # rebuilt contexts: the host-buffer entry point (in-process sharding) works again
        want = a.copy()
        lib.zkb_multi_device_set(10, 10, 28)
        zkb.best_fft(want, w, k)
        lib.zkb_multi_device_set(10, 10, 12)
        got = a.copy()
        zkb.best_fft(got, w, k)
        assert (got == want).all()
    finally:
        lib.zkb_dist_set_timeout_ms(0)
