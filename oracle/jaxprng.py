"""TEST INFRASTRUCTURE (oracle): JAX's default PRNG, restated in NumPy, so that the reference's jitters
are THE reference's and not a stand-in.

The reference draws its warm-start jitters from a FIXED key (src/caustics/extended_source.py:76-85,146):
    key = random.PRNGKey(0); key1, key2 = random.split(key)
    u1 = random.uniform(key1, shape=z_init.shape, minval=-1e-6, maxval=1e-6)   # z_init: (deg, n)
    u2 = random.uniform(key2, ...)
    ... z + random.uniform(key, shape=z.shape, minval=-1e-9, maxval=1e-9)       # z: (deg, npts)
so its extended-source results are deterministic.  jax is not installable here; the algorithm is JAX's
published default implementation ("threefry2x32", jax/_src/prng.py, jax/_src/random.py of the JAX
releases the reference's XLA custom-call API works with, i.e. before jax_threefry_partitionable became
the default in 0.5.0):

  threefry2x32        Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11), 20 rounds,
                      rotations (13, 15, 26, 6), (17, 29, 16, 24), key schedule constant 0x1BD11BDA
  PRNGKey(seed)       (seed >> 32, seed & 0xffffffff) as two uint32
  split(key, num)     threefry_2x32(key, iota(2 num)) reshaped (num, 2)
  threefry_2x32(key, count)   the counter array is cut in two halves (x0 | x1), one block per pair
  random_bits(64)     counters iota(2 N); element i = y0[i] << 32 | y1[i]  with (x0, x1) = (i, N + i)
  uniform(float64)    (bits >> 12 | bits_of(1.0)) reinterpreted as float64, minus 1.0, times (maxval -
                      minval) plus minval, clamped below by minval
Known answers pinned in tests/test_oracle.py: the Random123 vectors of threefry2x32, JAX's documented
split(PRNGKey(0)) = [[4146024105, 967050713], [2718843009, 1272950319]] and uniform(PRNGKey(0)) =
0.41845703 (float32).
"""
import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M32 = np.uint32(0xFFFFFFFF)


def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))) & _M32


def threefry2x32(k0, k1, x0, x1):
    """one block per element of the uint32 arrays x0, x1 under the key (k0, k1)"""
    with np.errstate(over="ignore"):
        k0, k1 = np.uint32(k0), np.uint32(k1)
        ks = (k0, k1, k0 ^ k1 ^ np.uint32(0x1BD11BDA))
        x0 = np.asarray(x0, dtype=np.uint32) + ks[0]
        x1 = np.asarray(x1, dtype=np.uint32) + ks[1]
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r) ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + np.uint32(g + 1)
    return x0, x1


def prng_key(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def _threefry_2x32(key, count):
    count = np.asarray(count, dtype=np.uint32).ravel()
    odd = count.size % 2
    if odd:
        count = np.concatenate([count, np.zeros(1, np.uint32)])
    h = count.size // 2
    y0, y1 = threefry2x32(key[0], key[1], count[:h], count[h:])
    out = np.concatenate([y0, y1])
    return out[:-1] if odd else out


def split(key, num=2):
    return _threefry_2x32(key, np.arange(2 * num, dtype=np.uint32)).reshape(num, 2)


def random_bits(key, bit_width, shape):
    size = int(np.prod(shape, dtype=np.int64))
    nwords = -(-bit_width * size // 32)
    bits = _threefry_2x32(key, np.arange(nwords, dtype=np.uint32))
    if bit_width == 64:
        hi, lo = bits[:size].astype(np.uint64), bits[size:].astype(np.uint64)
        bits = (hi << np.uint64(32)) | lo
    return bits.reshape(shape)


def uniform(key, shape=(), dtype=np.float64, minval=0.0, maxval=1.0):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        bits = random_bits(key, 64, shape)
        f = ((bits >> np.uint64(12)) | np.float64(1.0).view(np.uint64)).view(np.float64) - 1.0
    elif dtype == np.float32:
        bits = random_bits(key, 32, shape)
        f = ((bits >> np.uint32(9)) | np.float32(1.0).view(np.uint32)).view(np.float32) - np.float32(1.0)
    else:
        raise TypeError(dtype)
    lo, hi = dtype.type(minval), dtype.type(maxval)
    return np.maximum(lo, (f * (hi - lo) + lo).astype(dtype))


def limb_jitters(deg, n):
    """(deg, n) complex table the reference adds to the warm starts of EVERY refinement round
    (extended_source.py:76-85: the keys never change)"""
    k1, k2 = split(prng_key(0))
    return uniform(k1, (deg, n), minval=-1e-6, maxval=1e-6) + 1j * uniform(k2, (deg, n), minval=-1e-6, maxval=1e-6)


def duplicate_jitters(deg, npts):
    """(deg, npts) real table added to exact duplicates (extended_source.py:144-148)"""
    return uniform(prng_key(0), (deg, npts), minval=-1e-9, maxval=1e-9)
