"""halo2-shaped host API over libzkb200's C ABI (see package docstring)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _ffi
from ._ffi import check, u64p, u64pp


def lib():
    return _ffi.load()


def device_count() -> int:
    return lib().zkb_device_count()


def init(device=None) -> None:
    """Bind this process to one GPU (an int), to several GPUs of the box (a list — commits are then sharded by SRS point
    range, batches by column, one large NTT over NVLink peer memory, all behind the same calls), or to what ZKB_DEVICES /
    the current CUDA device select (None).  Raises ZkbError if no CUDA device is usable."""
    if device is None:
        check(lib().zkb_init(None, 0))
    else:
        devs = [int(device)] if isinstance(device, (int, np.integer)) else [int(d) for d in device]
        arr = (ctypes.c_int * len(devs))(*devs)
        check(lib().zkb_init(arr, len(devs)))


def bound_devices() -> list:
    """CUDA ordinals bound by init(), home device first."""
    arr = (ctypes.c_int * 8)()
    n = lib().zkb_bound_devices(arr, 8)
    return [int(arr[i]) for i in range(n)]


def shutdown() -> None:
    lib().zkb_shutdown()


def launch_count() -> int:
    return int(lib().zkb_launch_count())


def _fr(a, name="array") -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim == 1:
        a = a.reshape(-1, 4)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError(f"{name}: expected (n, 4) uint64 Montgomery limbs, got {a.shape}")
    return a


def _p(a: np.ndarray):
    return a.ctypes.data_as(u64p)


def omega(k: int) -> np.ndarray:
    """EvaluationDomain::get_omega for n = 2^k: Fr::ROOT_OF_UNITY^(2^(28-k))."""
    out = np.zeros(4, dtype=np.uint64)
    check(lib().zkb_fr_omega(k, _p(out)))
    return out


# ---- halo2_proofs::arithmetic ------------------------------------------------------------------------------------
def best_fft(a: np.ndarray, omega_: np.ndarray, log_n: int) -> None:
    """best_fft(a: &mut [Fr], omega, log_n): in place; asserts a.len() == 1 << log_n like the Rust original."""
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"] and a.flags["WRITEABLE"]):
        raise ValueError("best_fft operates in place on a contiguous writable uint64 array")
    v = a.reshape(-1, 4)
    assert v.shape[0] == 1 << log_n, "assertion failed: a.len() == 1 << log_n"
    w = np.ascontiguousarray(omega_, dtype=np.uint64).reshape(4)
    check(lib().zkb_ntt_fr(_p(v), _p(w), log_n))


def best_multiexp(coeffs: np.ndarray, bases: np.ndarray) -> np.ndarray:
    """best_multiexp(coeffs: &[Fr], bases: &[G1Affine]) -> G1 (12 limbs, z normalised)."""
    s = _fr(coeffs, "coeffs")
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    assert s.shape[0] == b.shape[0], "assertion failed: coeffs.len() == bases.len()"
    out = np.zeros(12, dtype=np.uint64)
    check(lib().zkb_msm_g1(_p(s), _p(b), s.shape[0], _p(out)))
    return out


def g1_sum(points_jac: np.ndarray) -> np.ndarray:
    """Host-side fold of partial MSM results (the `results.iter().fold(identity, |a, b| a + b)` of best_multiexp,
    applied to per-GPU shards)."""
    p = np.ascontiguousarray(points_jac, dtype=np.uint64).reshape(-1, 12)
    out = np.zeros(12, dtype=np.uint64)
    check(lib().zkb_g1_sum(_p(p), p.shape[0], _p(out)))
    return out


def g1_fixed_base_mul(scalars: np.ndarray) -> np.ndarray:
    s = _fr(scalars, "scalars")
    out = np.zeros((s.shape[0], 8), dtype=np.uint64)
    check(lib().zkb_g1_fixed_base_mul(_p(s), s.shape[0], _p(out)))
    return out


def g1_fixed_base_mul_naive(scalars: np.ndarray) -> np.ndarray:
    """The plain double-and-add kernel (independent of the windowed table path)."""
    s = _fr(scalars, "scalars")
    out = np.zeros((s.shape[0], 8), dtype=np.uint64)
    check(lib().zkb_g1_fixed_base_mul_naive(_p(s), s.shape[0], _p(out)))
    return out


def best_fft_g1(points_affine: np.ndarray, omega_: np.ndarray, log_n: int) -> np.ndarray:
    """best_fft::<Fr, G1>(a, omega, log_n) on affine points (returns the transformed points, normalised)."""
    p = np.ascontiguousarray(points_affine, dtype=np.uint64).reshape(-1, 8)
    assert p.shape[0] == 1 << log_n, "assertion failed: a.len() == 1 << log_n"
    out = np.zeros_like(p)
    check(lib().zkb_g1_ntt(_p(p), _p(out), _p(np.ascontiguousarray(omega_, dtype=np.uint64).reshape(4)), log_n))
    return out


def batch_normalize(points_jac: np.ndarray) -> np.ndarray:
    """halo2curves Curve::batch_normalize(&[G1], &mut [G1Affine])."""
    p = np.ascontiguousarray(points_jac, dtype=np.uint64).reshape(-1, 12)
    out = np.zeros((p.shape[0], 8), dtype=np.uint64)
    check(lib().zkb_g1_batch_normalize(_p(p), p.shape[0], _p(out)))
    return out


def eval_polynomial(poly: np.ndarray, point: np.ndarray) -> np.ndarray:
    """arithmetic::eval_polynomial(poly: &[Fr], point: Fr) -> Fr"""
    a = _fr(poly, "poly")
    out = np.zeros(4, dtype=np.uint64)
    check(lib().zkb_fr_eval_polynomial(_p(a), a.shape[0], _p(np.ascontiguousarray(point, dtype=np.uint64).reshape(4)), _p(out)))
    return out


def kate_division(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """arithmetic::kate_division(a, b): the len-1 coefficients of a(X) / (X - b) (remainder discarded, as upstream)."""
    v = _fr(a, "a")
    assert v.shape[0] >= 1
    out = np.zeros((v.shape[0] - 1, 4), dtype=np.uint64)
    check(lib().zkb_fr_kate_division(_p(v), v.shape[0], _p(np.ascontiguousarray(b, dtype=np.uint64).reshape(4)), _p(out)))
    return out


def batch_invert(values: np.ndarray) -> np.ndarray:
    """ff::BatchInvert: element-wise inverse, zeros stay zero."""
    v = np.array(_fr(values, "values"), copy=True)
    check(lib().zkb_fr_batch_invert(_p(v), v.shape[0]))
    return v


class Polynomial:
    """A polynomial (or column of evaluations) resident in HBM: uploaded once, then committed / transformed / evaluated on
    the device (the prover's commit_lagrange -> lagrange_to_coeff -> coeff_to_extended -> eval -> kate_division chain)."""

    def __init__(self, values: np.ndarray | None = None, _handle: int | None = None):
        self._h = ctypes.c_uint64(0)
        if _handle is not None:
            self._h.value = _handle
        else:
            v = _fr(values, "values")
            check(lib().zkb_poly_upload(_p(v), v.shape[0], ctypes.byref(self._h)))

    def __len__(self) -> int:
        n = ctypes.c_size_t(0)
        check(lib().zkb_poly_len(self._h, ctypes.byref(n)))
        return n.value

    def to_host(self) -> np.ndarray:
        out = np.zeros((len(self), 4), dtype=np.uint64)
        check(lib().zkb_poly_download(self._h, _p(out), out.shape[0]))
        return out

    def write(self, offset: int, values: np.ndarray) -> "Polynomial":
        """self[offset : offset + len(values)] = values (e.g. the random blinding rows of a product column)"""
        v = _fr(values, "values")
        check(lib().zkb_poly_write(self._h, offset, _p(v), v.shape[0]))
        return self

    @classmethod
    def zeros(cls, n: int) -> "Polynomial":
        """n zero elements allocated and cleared on the device (nothing crosses PCIe)"""
        h = ctypes.c_uint64(0)
        check(lib().zkb_poly_alloc(n, ctypes.byref(h)))
        return cls(_handle=h.value)

    @classmethod
    def load_file(cls, path: str, offset: int, n: int) -> "Polynomial":
        """n raw Fr (RawBytesUnchecked element encoding) from a file straight into HBM (proving-key residency)"""
        h = ctypes.c_uint64(0)
        check(lib().zkb_poly_load_file(path.encode(), offset, n, ctypes.byref(h)))
        return cls(_handle=h.value)

    def slice(self, offset: int, n: int) -> "Polynomial":
        """A new resident polynomial holding self[offset : offset + n] (the pieces of h(X) after extended_to_coeff)."""
        h = ctypes.c_uint64(0)
        check(lib().zkb_poly_slice(self._h, offset, n, ctypes.byref(h)))
        return Polynomial(_handle=h.value)

    def commit(self, params: "ParamsKZG", lagrange: bool = False) -> np.ndarray:
        out = np.zeros(12, dtype=np.uint64)
        h = params.handle_g_lagrange if lagrange else params.handle_g
        check(lib().zkb_poly_commit(h, self._h, _p(out)))
        return out

    def lagrange_to_coeff(self, domain: "EvaluationDomain") -> "Polynomial":
        check(lib().zkb_poly_lagrange_to_coeff(self._h, domain.k))
        return self

    def coeff_to_lagrange(self, domain: "EvaluationDomain") -> "Polynomial":
        check(lib().zkb_poly_coeff_to_lagrange(self._h, domain.k))
        return self

    def coeff_to_extended(self, domain: "EvaluationDomain") -> "Polynomial":
        h = ctypes.c_uint64(0)
        check(lib().zkb_poly_coeff_to_extended(self._h, domain.k, domain.extended_k, ctypes.byref(h)))
        return Polynomial(_handle=h.value)

    def extended_to_coeff(self, domain: "EvaluationDomain") -> "Polynomial":
        check(lib().zkb_poly_extended_to_coeff(self._h, domain.k, domain.extended_k))
        return self

    def eval(self, point: np.ndarray) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        check(lib().zkb_poly_eval(self._h, _p(np.ascontiguousarray(point, dtype=np.uint64).reshape(4)), _p(out)))
        return out

    def kate_division(self, b: np.ndarray) -> "Polynomial":
        h = ctypes.c_uint64(0)
        check(lib().zkb_poly_kate_division(self._h, _p(np.ascontiguousarray(b, dtype=np.uint64).reshape(4)), ctypes.byref(h)))
        return Polynomial(_handle=h.value)

    def batch_invert(self) -> "Polynomial":
        check(lib().zkb_poly_batch_invert(self._h))
        return self

    def mul(self, other: "Polynomial") -> "Polynomial":
        check(lib().zkb_poly_mul(self._h, other._h))
        return self

    def scale_add(self, k: np.ndarray, other: "Polynomial | None" = None) -> "Polynomial":
        """self <- self * k + other (Horner fold of query polynomials in the multiopen provers)"""
        check(lib().zkb_poly_scale_add(self._h, _p(np.ascontiguousarray(k, dtype=np.uint64).reshape(4)), other._h if other is not None else 0))
        return self

    def add_const(self, c: np.ndarray) -> "Polynomial":
        check(lib().zkb_poly_add_const(self._h, _p(np.ascontiguousarray(c, dtype=np.uint64).reshape(4))))
        return self

    def prefix_product(self) -> "Polynomial":
        """z[0] = 1, z[i] = prod_{j<i} v[j] (the scan that builds the permutation / lookup grand products)"""
        check(lib().zkb_poly_prefix_product(self._h))
        return self

    def free(self) -> None:
        if self._h.value:
            lib().zkb_poly_free(self._h)
            self._h.value = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- halo2_proofs::poly::EvaluationDomain ------------------------------------------------------------------------------
class EvaluationDomain:
    """EvaluationDomain::<Fr>::new(j, k): j = cs.degree(), n = 2^k."""

    def __init__(self, j: int, k: int):
        self.k = k
        self.n = 1 << k
        self.quotient_poly_degree = j - 1
        ek = k
        while (1 << ek) < self.n * self.quotient_poly_degree:
            ek += 1
        self.extended_k = ek

    def extended_len(self) -> int:
        return 1 << self.extended_k

    def get_omega(self) -> np.ndarray:
        return omega(self.k)

    def get_extended_omega(self) -> np.ndarray:
        return omega(self.extended_k)

    def lagrange_to_coeff(self, a: np.ndarray) -> np.ndarray:
        v = np.array(_fr(a), copy=True)
        assert v.shape[0] == self.n, "assertion failed: a.values.len() == 1 << self.k"
        check(lib().zkb_lagrange_to_coeff(_p(v), self.k))
        return v

    def lagrange_to_coeff_batch(self, cols: list[np.ndarray]) -> list[np.ndarray]:
        vs = [np.array(_fr(a), copy=True) for a in cols]
        for v in vs:
            assert v.shape[0] == self.n
        ptrs = (u64p * len(vs))(*[_p(v) for v in vs])
        check(lib().zkb_lagrange_to_coeff_batch(ptrs, len(vs), self.k))
        return vs

    def coeff_to_lagrange(self, a: np.ndarray) -> np.ndarray:
        v = np.array(_fr(a), copy=True)
        assert v.shape[0] == self.n
        check(lib().zkb_coeff_to_lagrange(_p(v), self.k))
        return v

    def t_evaluations_inverse(self) -> np.ndarray:
        """1 / (X^n - 1) on the extended coset zeta <w_ext> (upstream's `t_evaluations`, inverted as divide_by_vanishing_poly uses
        them): X^n takes only 2^(extended_k - k) distinct values there, so the (extended_len, 4) array is that many Montgomery
        values tiled.  Host integers; no GPU needed."""
        return np.tile(self._t_inverse_period(), (self.extended_len() >> (self.extended_k - self.k), 1))

    def _t_inverse_period(self) -> np.ndarray:
        """the 2^(extended_k - k) distinct values of 1 / (X^n - 1) on the extended coset, Montgomery limbs"""
        r = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
        R = (1 << 256) % r
        Rinv = pow(R, -1, r)
        zeta = pow(7, 2 * (r - 1) // 3, r)                                   # Fr::ZETA, the coset generator
        w = sum(int(x) << (64 * i) for i, x in enumerate(self.get_extended_omega())) * Rinv % r
        period = 1 << (self.extended_k - self.k)
        vals = [pow((pow(zeta, self.n, r) * pow(w, i * self.n, r) - 1) % r, -1, r) * R % r for i in range(period)]
        return np.array([[(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)] for v in vals], dtype=np.uint64)

    def divide_by_vanishing_poly(self, h: "Polynomial") -> "Polynomial":
        """EvaluationDomain::divide_by_vanishing_poly on a resident polynomial of extended evaluations: h[i] *= 1 / (X^n - 1) at the
        i-th coset point.  Only the 2^(extended_k - k) distinct inverses travel (in the kernel arguments); no extended-size array
        of them is built, uploaded or read."""
        if getattr(self, "_t_inv", None) is None:
            self._t_inv = np.ascontiguousarray(self._t_inverse_period())
        check(lib().zkb_poly_mul_periodic(h._h, _p(self._t_inv), self._t_inv.shape[0]))
        return h

    def coeff_to_extended(self, a: np.ndarray) -> np.ndarray:
        v = _fr(a)
        assert v.shape[0] == self.n, "assertion failed: a.values.len() == 1 << self.k"
        out = np.zeros((self.extended_len(), 4), dtype=np.uint64)
        check(lib().zkb_coeff_to_extended(_p(v), _p(out), self.k, self.extended_k))
        return out

    def coeff_to_extended_batch(self, cols: list[np.ndarray]) -> list[np.ndarray]:
        vs = [_fr(a) for a in cols]
        for v in vs:
            assert v.shape[0] == self.n
        outs = [np.zeros((self.extended_len(), 4), dtype=np.uint64) for _ in vs]
        pi = (u64p * len(vs))(*[_p(v) for v in vs])
        po = (u64p * len(vs))(*[_p(o) for o in outs])
        check(lib().zkb_coeff_to_extended_batch(pi, po, len(vs), self.k, self.extended_k))
        return outs

    def extended_to_coeff(self, a: np.ndarray) -> np.ndarray:
        v = np.array(_fr(a), copy=True)
        assert v.shape[0] == self.extended_len(), "assertion failed: a.values.len() == self.extended_len()"
        check(lib().zkb_extended_to_coeff(_p(v), self.k, self.extended_k))
        return v[: self.n * self.quotient_poly_degree]


# ---- halo2_proofs::poly::kzg::commitment::ParamsKZG --------------------------------------------------------------------
class ParamsKZG:
    """The MSM-facing part of ParamsKZG: `g` (monomial SRS) and `g_lagrange`, kept resident in HBM."""

    def __init__(self, k: int, g: np.ndarray, g_lagrange: np.ndarray | None = None):
        self.k = k
        self.n = 1 << k
        g = np.ascontiguousarray(g, dtype=np.uint64).reshape(-1, 8)
        assert g.shape[0] == self.n, "g must hold 2^k points"
        self._h_g = ctypes.c_uint64(0)
        check(lib().zkb_srs_register(_p(g), g.shape[0], ctypes.byref(self._h_g)))
        self._h_gl = None
        if g_lagrange is not None:
            gl = np.ascontiguousarray(g_lagrange, dtype=np.uint64).reshape(-1, 8)
            assert gl.shape[0] == self.n
            self._h_gl = ctypes.c_uint64(0)
            check(lib().zkb_srs_register(_p(gl), gl.shape[0], ctypes.byref(self._h_gl)))

    @classmethod
    def setup(cls, k: int, s: np.ndarray) -> "ParamsKZG":
        """ParamsKZG::<Bn256>::setup(k, rng) with the sampled scalar `s` passed in (Montgomery Fr): g and g_lagrange are
        generated on the device and stay resident — nothing crosses PCIe.  (g2, s_g2 stay on the host side.)"""
        self = cls.__new__(cls)
        self.k, self.n = k, 1 << k
        self._h_g, self._h_gl = ctypes.c_uint64(0), ctypes.c_uint64(0)
        sv = np.ascontiguousarray(s, dtype=np.uint64).reshape(4)
        check(lib().zkb_kzg_setup_resident(k, _p(sv), ctypes.byref(self._h_g), ctypes.byref(self._h_gl)))
        return self

    @classmethod
    def read(cls, path: str, check_points: bool = True) -> "ParamsKZG":
        """ParamsKZG::read_custom(reader, SerdeFormat::RawBytes) for the G1 side, straight from the file into HBM
        (the `kzg_bn254_{k}.srs` files of the reference's demo, /root/reference: app/src/.../worker.js:218-225 and
        snark-verifier-sdk's gen_srs).  Layout [UPSTREAM-UNVERIFIED: halo2-axiom is not vendored] as ParamsKZG::write_custom
        writes it: u32 little-endian k, g[2^k], g_lagrange[2^k] as raw 64-byte G1Affine (Montgomery limbs), then g2 and s_g2
        (raw G2Affine, 128 bytes each), which stay on the host and are not read here.  check_points=False is
        SerdeFormat::RawBytesUnchecked."""
        with open(path, "rb") as f:
            head = f.read(4)
        if len(head) != 4:
            raise ValueError("params file too short")
        k = int.from_bytes(head, "little")
        if not 0 < k <= 28:
            raise ValueError(f"params file claims k = {k}")
        self = cls.__new__(cls)
        self.k, self.n = k, 1 << k
        self._h_g, self._h_gl = ctypes.c_uint64(0), ctypes.c_uint64(0)
        check(lib().zkb_srs_load_file(path.encode(), 4, self.n, int(check_points), ctypes.byref(self._h_g)))
        check(lib().zkb_srs_load_file(path.encode(), 4 + self.n * 64, self.n, int(check_points), ctypes.byref(self._h_gl)))
        return self

    def write(self, path: str, g2: bytes = bytes(128), s_g2: bytes = bytes(128)) -> None:
        """The same layout back to a file (G2 elements supplied by the caller, who holds them on the host)."""
        with open(path, "wb") as f:
            f.write(self.k.to_bytes(4, "little"))
            f.write(self.get_g().tobytes())
            f.write(self.get_g_lagrange().tobytes())
            f.write(g2)
            f.write(s_g2)

    @classmethod
    def from_g(cls, k: int, g: np.ndarray) -> "ParamsKZG":
        """Params from the monomial basis only (an SRS file without g_lagrange): g_lagrange = g_to_lagrange(g, k) on the
        device (inverse G1 FFT and 1/n)."""
        self = cls(k, g)
        self._h_gl = ctypes.c_uint64(0)
        check(lib().zkb_srs_g_to_lagrange(self._h_g, k, ctypes.byref(self._h_gl)))
        return self

    def get_g(self) -> np.ndarray:
        out = np.zeros((self.n, 8), dtype=np.uint64)
        check(lib().zkb_srs_download(self._h_g, _p(out), self.n))
        return out

    def get_g_lagrange(self) -> np.ndarray:
        out = np.zeros((self.n, 8), dtype=np.uint64)
        check(lib().zkb_srs_download(self._h_gl, _p(out), self.n))
        return out

    def _commit(self, h, poly: np.ndarray) -> np.ndarray:
        s = _fr(poly, "poly")
        assert s.shape[0] <= self.n, "assertion failed: bases.len() >= scalars.len()"
        out = np.zeros(12, dtype=np.uint64)
        check(lib().zkb_msm_g1_srs(h, _p(s), s.shape[0], _p(out)))
        return out

    def commit(self, poly: np.ndarray) -> np.ndarray:
        """commit(&Polynomial<Fr, Coeff>, Blind) -> G1   (the blind is unused under KZG)"""
        return self._commit(self._h_g, poly)

    def commit_lagrange(self, poly: np.ndarray) -> np.ndarray:
        """commit_lagrange(&Polynomial<Fr, LagrangeCoeff>, Blind) -> G1"""
        if self._h_gl is None:
            raise ValueError("g_lagrange was not supplied")
        return self._commit(self._h_gl, poly)

    def commit_batch(self, polys: list[np.ndarray], lagrange: bool = False) -> np.ndarray:
        h = self._h_gl if lagrange else self._h_g
        vs = [_fr(p) for p in polys]
        n = vs[0].shape[0]
        for v in vs:
            assert v.shape[0] == n
        ptrs = (u64p * len(vs))(*[_p(v) for v in vs])
        out = np.zeros((len(vs), 12), dtype=np.uint64)
        check(lib().zkb_msm_g1_srs_batch(h, ptrs, len(vs), n, _p(out)))
        return out

    def commit_range(self, offset: int, scalars: np.ndarray, lagrange: bool = False) -> np.ndarray:
        """Partial commitment over srs[offset .. offset+len) — one GPU's shard of a point-range-sharded MSM."""
        h = self._h_gl if lagrange else self._h_g
        s = _fr(scalars)
        out = np.zeros(12, dtype=np.uint64)
        check(lib().zkb_msm_g1_srs_range(h, offset, _p(s), s.shape[0], _p(out)))
        return out

    @property
    def handle_g(self) -> int:
        return self._h_g.value

    @property
    def handle_g_lagrange(self) -> int:
        return self._h_gl.value if self._h_gl is not None else 0

    def close(self) -> None:
        for h in (self._h_g, self._h_gl):
            if h is not None and h.value:
                lib().zkb_srs_release(h)
                h.value = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- measurement helpers ---------------------------------------------------------------------------------------------------
class prof:
    @staticmethod
    def enable(on: bool = True) -> None:
        check(lib().zkb_prof_enable(1 if on else 0))

    @staticmethod
    def reset() -> None:
        check(lib().zkb_prof_reset())

    @staticmethod
    def get(name: str) -> tuple[float, int]:
        ms = ctypes.c_double(0)
        n = ctypes.c_uint64(0)
        check(lib().zkb_prof_get(name.encode(), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value
