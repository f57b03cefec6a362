"""Pure-Python big-int twin of the hot path (TEST INFRASTRUCTURE — never imported by the product).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this.

What it restates
----------------
The reference repo (aerius-labs/zksnap-circuits-halo2) holds no field / curve / FFT / MSM code of its
own: the hot path lives in the un-vendored git dependencies ``halo2-axiom`` (``halo2_proofs``) and
``halo2curves-axiom`` (floating branches, no Cargo.lock — see SURVEY.md §8c).  This file therefore
restates their *published algorithms* from the mathematical definitions, anchored on the call sites the
reference does own:

* ``create_proof`` / ``keygen_pk`` callers  — /root/reference/aggregator/src/wrapper.rs:106-109,129-137
* bench drivers                             — /root/reference/voter/benches/voter_circuit.rs:60-62,80
                                              /root/reference/aggregator/benches/state_transition_circuit.rs:64-66,84
                                              /root/reference/aggregator/benches/wrapper_circuit.rs:107,140

PARITY UNPINNED by reference fixtures: the reference has no golden vectors for this path.  The twin is
pinned instead to first-principles known-answer values (SURVEY.md §8c [COMPUTED]) in
``tests/test_oracle.py`` and cross-checked against the independent C restatement ``oracle/zkb_oracle.c``.

Upstream functions followed (names only; sources not on disk):
  halo2_proofs::arithmetic::{best_fft, best_multiexp, multiexp_serial}
  halo2_proofs::poly::EvaluationDomain::{new, lagrange_to_coeff, coeff_to_extended, extended_to_coeff}
  halo2_proofs::poly::kzg::commitment::ParamsKZG::{commit, commit_lagrange}
  halo2curves::bn256::{Fr, Fq, G1Affine, G1}
"""
from __future__ import annotations

import math

# ----------------------------------------------------------------------------------------------
# BN254 constants (halo2curves bn256::{Fq, Fr})
# ----------------------------------------------------------------------------------------------
FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # base field modulus p
FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # scalar field modulus r
R256 = 1 << 256
FR_S = 28  # two-adicity of r-1
FR_GENERATOR = 7
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (FR - 1) >> FR_S, FR)
FR_ROOT_OF_UNITY_INV = pow(FR_ROOT_OF_UNITY, FR - 2, FR)
FR_ZETA = pow(FR_GENERATOR, 2 * (FR - 1) // 3, FR)  # primitive cube root of unity used as coset shift
CURVE_B = 3
G1_GENERATOR = (1, 2)

FR_R = R256 % FR
FR_R2 = FR_R * FR_R % FR
FQ_R = R256 % FQ
FQ_R2 = FQ_R * FQ_R % FQ
FR_INV64 = (-pow(FR, -1, 1 << 64)) % (1 << 64)
FQ_INV64 = (-pow(FQ, -1, 1 << 64)) % (1 << 64)


# ----------------------------------------------------------------------------------------------
# Memory encodings: 4 little-endian u64 limbs holding a*2^256 mod m (Montgomery form), as halo2curves
# keeps `Fr`/`Fq` in memory (SURVEY.md §8 "Data layouts").
# ----------------------------------------------------------------------------------------------
def to_mont(a: int, m: int) -> int:
    return (a % m) * R256 % m


def from_mont(a: int, m: int) -> int:
    return a * pow(R256, -1, m) % m


def limbs(a: int) -> list[int]:
    return [(a >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def unlimbs(l) -> int:
    return int(l[0]) | (int(l[1]) << 64) | (int(l[2]) << 128) | (int(l[3]) << 192)


def fr_encode(a: int) -> list[int]:
    """canonical integer -> 4 Montgomery limbs"""
    return limbs(to_mont(a, FR))


def fr_decode(l) -> int:
    return from_mont(unlimbs(l), FR)


def fq_encode(a: int) -> list[int]:
    return limbs(to_mont(a, FQ))


def fq_decode(l) -> int:
    return from_mont(unlimbs(l), FQ)


# ----------------------------------------------------------------------------------------------
# Fr FFT — halo2_proofs::arithmetic::best_fft (bit-reverse, twiddle table, radix-2 DIT stages).
# In-place, natural order in, natural order out: out[i] = sum_j a[j] * omega^(i*j).
# ----------------------------------------------------------------------------------------------
def omega_for(k: int) -> int:
    """2^k-th primitive root of unity: ROOT_OF_UNITY^(2^(S-k)) (EvaluationDomain::new)."""
    assert 0 <= k <= FR_S
    w = FR_ROOT_OF_UNITY
    for _ in range(k, FR_S):
        w = w * w % FR
    return w


def bitreverse(n: int, l: int) -> int:
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def best_fft(a: list[int], omega: int, log_n: int) -> None:
    n = len(a)
    assert n == 1 << log_n
    for k in range(n):
        rk = bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    tw = [1] * max(n // 2, 1)
    for i in range(1, n // 2):
        tw[i] = tw[i - 1] * omega % FR
    chunk, tchunk = 2, n // 2
    for _ in range(log_n):
        half = chunk // 2
        for start in range(0, n, chunk):
            for i in range(half):
                t = a[start + half + i] * tw[i * tchunk] % FR
                u = a[start + i]
                a[start + i] = (u + t) % FR
                a[start + half + i] = (u - t) % FR
        chunk *= 2
        tchunk //= 2


def dft_naive(a: list[int], omega: int) -> list[int]:
    n = len(a)
    return [sum(a[j] * pow(omega, i * j, FR) for j in range(n)) % FR for i in range(n)]


# ----------------------------------------------------------------------------------------------
# EvaluationDomain — halo2_proofs::poly::domain::EvaluationDomain
# ----------------------------------------------------------------------------------------------
class EvaluationDomain:
    def __init__(self, j: int, k: int):
        self.k = k
        self.n = 1 << k
        self.quotient_poly_degree = j - 1
        ek = k
        while (1 << ek) < self.n * self.quotient_poly_degree:
            ek += 1
        self.extended_k = ek
        self.extended_omega = omega_for(ek)
        self.extended_omega_inv = pow(self.extended_omega, FR - 2, FR)
        w = self.extended_omega
        for _ in range(k, ek):
            w = w * w % FR
        self.omega = w
        self.omega_inv = pow(w, FR - 2, FR)
        self.g_coset = FR_ZETA
        self.g_coset_inv = FR_ZETA * FR_ZETA % FR
        self.ifft_divisor = pow(1 << k, FR - 2, FR)
        self.extended_ifft_divisor = pow(1 << ek, FR - 2, FR)

    def extended_len(self) -> int:
        return 1 << self.extended_k

    @staticmethod
    def _ifft(a, omega_inv, log_n, divisor):
        best_fft(a, omega_inv, log_n)
        for i in range(len(a)):
            a[i] = a[i] * divisor % FR

    def lagrange_to_coeff(self, a: list[int]) -> list[int]:
        assert len(a) == self.n
        a = list(a)
        self._ifft(a, self.omega_inv, self.k, self.ifft_divisor)
        return a

    def coeff_to_lagrange(self, a: list[int]) -> list[int]:
        assert len(a) == self.n
        a = list(a)
        best_fft(a, self.omega, self.k)
        return a

    def _distribute_powers_zeta(self, a, into_coset: bool):
        cp = [self.g_coset, self.g_coset_inv] if into_coset else [self.g_coset_inv, self.g_coset]
        for idx in range(len(a)):
            i = idx % 3
            if i:
                a[idx] = a[idx] * cp[i - 1] % FR

    def coeff_to_extended(self, a: list[int]) -> list[int]:
        assert len(a) == self.n
        a = list(a)
        self._distribute_powers_zeta(a, True)
        a += [0] * (self.extended_len() - self.n)
        best_fft(a, self.extended_omega, self.extended_k)
        return a

    def extended_to_coeff(self, a: list[int]) -> list[int]:
        assert len(a) == self.extended_len()
        a = list(a)
        self._ifft(a, self.extended_omega_inv, self.extended_k, self.extended_ifft_divisor)
        self._distribute_powers_zeta(a, False)
        return a[: self.n * self.quotient_poly_degree]


# ----------------------------------------------------------------------------------------------
# G1: y^2 = x^3 + 3 over Fq.  Affine identity encoded as (0, 0) like halo2curves' G1Affine;
# here identity is None.
# ----------------------------------------------------------------------------------------------
def g1_is_on_curve(P) -> bool:
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - CURVE_B) % FQ == 0


def g1_neg(P):
    return None if P is None else (P[0], (-P[1]) % FQ)


def g1_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % FQ == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, FQ) % FQ
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, FQ) % FQ
    x3 = (lam * lam - x1 - x2) % FQ
    return (x3, (lam * (x1 - x3) - y1) % FQ)


def g1_mul(P, s: int):
    s %= FR
    acc = None
    while s:
        if s & 1:
            acc = g1_add(acc, P)
        P = g1_add(P, P)
        s >>= 1
    return acc


def msm_naive(scalars: list[int], bases: list) -> object:
    acc = None
    for s, b in zip(scalars, bases):
        acc = g1_add(acc, g1_mul(b, s))
    return acc


def multiexp_serial(coeffs: list[int], bases: list, acc):
    """halo2_proofs::arithmetic::multiexp_serial — unsigned c-bit windows, running-sum buckets."""
    n = len(bases)
    c = 1 if n < 4 else 3 if n < 32 else math.ceil(math.log(n))
    segments = 256 // c + 1
    for seg in reversed(range(segments)):
        for _ in range(c):
            acc = g1_add(acc, acc)
        buckets = [None] * ((1 << c) - 1)
        for s, b in zip(coeffs, bases):
            d = (s >> (seg * c)) & ((1 << c) - 1)
            if d:
                buckets[d - 1] = g1_add(buckets[d - 1], b)
        running = None
        for e in reversed(buckets):
            running = g1_add(running, e)
            acc = g1_add(acc, running)
    return acc


def best_multiexp(coeffs: list[int], bases: list, num_threads: int = 8):
    """halo2_proofs::arithmetic::best_multiexp — contiguous chunks per thread, fold partial sums."""
    assert len(coeffs) == len(bases)
    if len(coeffs) > num_threads:
        chunk = len(coeffs) // num_threads
        acc = None
        for i in range(0, len(coeffs), chunk):
            acc = g1_add(acc, multiexp_serial(coeffs[i : i + chunk], bases[i : i + chunk], None))
        return acc
    return multiexp_serial(coeffs, bases, None)


def g1_affine_encode(P) -> list[int]:
    """G1Affine in memory: x,y Montgomery limbs; identity = all zero."""
    if P is None:
        return [0] * 8
    return fq_encode(P[0]) + fq_encode(P[1])


def g1_affine_decode(l):
    x, y = unlimbs(l[0:4]), unlimbs(l[4:8])
    if x == 0 and y == 0:
        return None
    return (from_mont(x, FQ), from_mont(y, FQ))


def g1_jacobian_decode(l):
    """G1 (x,y,z) Jacobian Montgomery limbs -> affine tuple / None."""
    x, y, z = (from_mont(unlimbs(l[4 * i : 4 * i + 4]), FQ) for i in range(3))
    if z == 0:
        return None
    zi = pow(z, -1, FQ)
    return (x * zi * zi % FQ, y * zi * zi * zi % FQ)
