"""Shared test helpers: seeded synthetic field elements / points in halo2curves' memory layout."""
import numpy as np

FR_LIMBS = np.array([0x43E1F593F0000001, 0x2833E84879B97091, 0xB85045B68181585D, 0x30644E72E131A029], dtype=np.uint64)
FQ_LIMBS = np.array([0x3C208C16D87CFD47, 0x97816A916871CA8D, 0xB85045B68181585D, 0x30644E72E131A029], dtype=np.uint64)


def _lt(a: np.ndarray, m: np.ndarray) -> np.ndarray:
    """rowwise a < m for (n,4) LE limb arrays"""
    lt = np.zeros(a.shape[0], dtype=bool)
    eq = np.ones(a.shape[0], dtype=bool)
    for i in (3, 2, 1, 0):
        lt |= eq & (a[:, i] < m[i])
        eq &= a[:, i] == m[i]
    return lt


def random_field(n: int, seed: int, modulus: np.ndarray = FR_LIMBS) -> np.ndarray:
    """n uniform elements < modulus as (n,4) uint64 (rejection sampling on 254 bits).

    A uniform canonical residue read as a Montgomery residue is still uniform, so this is directly a
    valid in-memory `Fr`/`Fq` array."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((n, 4), dtype=np.uint64)
    filled = 0
    while filled < n:
        m = max(16, int((n - filled) * 1.4))
        c = rng.integers(0, 2**64, size=(m, 4), dtype=np.uint64)
        c[:, 3] &= np.uint64((1 << 62) - 1)
        c = c[_lt(c, modulus)]
        take = min(len(c), n - filled)
        out[filled : filled + take] = c[:take]
        filled += take
    return out


def int_to_limbs(x: int) -> np.ndarray:
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def limbs_to_int(l) -> int:
    return sum(int(l[i]) << (64 * i) for i in range(4))


def ints_to_limbs(xs) -> np.ndarray:
    return np.array([[(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for x in xs], dtype=np.uint64).reshape(-1, 4)


def limbs_to_ints(a: np.ndarray) -> list:
    a = np.asarray(a).reshape(-1, 4)
    return [limbs_to_int(r) for r in a]
