"""Synthetic inputs of SURVEY.md section 8(d): deterministic, generated from xorshift64*.

Used by the tests, bench.py (both arms) and smoke(), so that the GPU path and the CPU oracle see
identical bytes.
"""
import numpy as np

SEED0 = 0x5A5041512D560001


def _xorshift_stream(seed, n):
    """n uint64 values of xorshift64* (vectorised in chunks by running 4096 independent lanes)."""
    lanes = 4096
    s = (np.arange(lanes, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)) | np.uint64(1)
    out = np.empty((n + lanes - 1) // lanes * lanes, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for _ in range(4):  # warm up so that nearby seeds decorrelate
            s ^= s >> np.uint64(12); s ^= s << np.uint64(25); s ^= s >> np.uint64(27)
        for i in range(0, len(out), lanes):
            s ^= s >> np.uint64(12); s ^= s << np.uint64(25); s ^= s >> np.uint64(27)
            out[i:i + lanes] = s * np.uint64(0x2545F4914F6CDD1D)
    return out[:n]


_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_W = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0,
                      1.9, 1.5, 1.0, 0.8, 0.15, 0.15, 0.1, 0.07])


def _vocabulary(seed):
    r = _xorshift_stream(seed ^ 0xABCDEF, 4096 * 16)
    cdf = np.cumsum(_LETTER_W) / _LETTER_W.sum()
    words = []
    k = 0
    for _ in range(4096):
        ln = 2 + int(r[k] % 11); k += 1
        u = (r[k:k + ln] >> np.uint64(11)).astype(np.float64) / float(1 << 53); k += ln
        words.append(_LETTERS[np.searchsorted(cdf, u)].tobytes())
    return words


_vocab_cache = {}


def text(n, seed=SEED0):
    """n bytes of word-sampled text: Zipf(s=1.1) over a 4096-word vocabulary, newline every ~72 chars."""
    if n == 0:
        return b""
    if "v" not in _vocab_cache:  # one fixed vocabulary for every seed
        _vocab_cache["v"] = _vocabulary(SEED0)
    words = _vocab_cache["v"]
    ranks = np.arange(1, 4097, dtype=np.float64)
    w = ranks ** -1.1
    cdf = np.cumsum(w) / w.sum()
    n_words = n // 4 + 16
    u = (_xorshift_stream(seed, n_words) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    idx = np.searchsorted(cdf, u)
    out = bytearray()
    col = 0
    for i in idx:
        wd = words[int(i)]
        out += wd
        col += len(wd) + 1
        if col >= 72:
            out += b"\n"
            col = 0
        else:
            out += b" "
        if len(out) >= n:
            break
    while len(out) < n:
        out += b" "
    return bytes(out[:n])


def random_bytes(n, seed=SEED0 + 1):
    if n == 0:
        return b""
    return _xorshift_stream(seed, (n + 7) // 8).view(np.uint8)[:n].tobytes()


def structured(n, seed=SEED0 + 2):
    """Little-endian u32 counters + 16-byte repeating records with 10 % noise."""
    if n == 0:
        return b""
    rec = np.frombuffer(b"RECORD\x00\x01\x10\x20\x30\x40ABCD", dtype=np.uint8)
    m = (n + 19) // 20
    body = np.zeros((m, 20), dtype=np.uint8)
    body[:, :4] = np.arange(m, dtype="<u4").view(np.uint8).reshape(m, 4)
    body[:, 4:] = rec
    r = _xorshift_stream(seed, m * 3)
    noisy = (r[:m] % np.uint64(10)) == 0
    pos = (r[m:2 * m] % np.uint64(16)).astype(np.int64) + 4
    val = (r[2 * m:] & np.uint64(255)).astype(np.uint8)
    rows = np.nonzero(noisy)[0]
    body[rows, pos[rows]] = val[rows]
    return body.reshape(-1)[:n].tobytes()


def mixed_block(k, n, seed=SEED0 + 3):
    """cfg 3: block k is text if k%4 in {0,1}, random if k%4==2, structured if k%4==3."""
    r = k % 4
    if r < 2:
        return text(n, seed + k)
    if r == 2:
        return random_bytes(n, seed + k)
    return structured(n, seed + k)


def text_blocks(n_blocks, block_bytes, seed=SEED0):
    """A long text stream cut into blocks (cfg 2); generated once and sliced."""
    whole = text(n_blocks * block_bytes, seed)
    return [whole[i * block_bytes:(i + 1) * block_bytes] for i in range(n_blocks)]
