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


def _vocab_arrays():
    """The fixed vocabulary as a padded byte matrix (word + separator slot) and word lengths."""
    if "v" not in _vocab_cache:
        words = _vocabulary(SEED0)
        mat = np.zeros((4096, 13), dtype=np.uint8)
        lens = np.zeros(4096, dtype=np.int64)
        for i, w in enumerate(words):
            mat[i, :len(w)] = np.frombuffer(w, dtype=np.uint8)
            lens[i] = len(w)
        ranks = np.arange(1, 4097, dtype=np.float64)
        wgt = ranks ** -1.1
        _vocab_cache["v"] = (mat, lens, np.cumsum(wgt) / wgt.sum())
    return _vocab_cache["v"]


def _text_chunk(n, seed):
    mat, lens, cdf = _vocab_arrays()
    n_words = n // 3 + 64  # mean word length with separator is well above 3
    u = (_xorshift_stream(seed, n_words) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    idx = np.searchsorted(cdf, u)
    wl = lens[idx] + 1                       # word + separator
    end = np.cumsum(wl)
    start = end - wl
    keep = int(np.searchsorted(end, n, side="left")) + 1
    idx, wl, end, start = idx[:keep], wl[:keep], end[:keep], start[:keep]
    out = np.full(int(end[-1]), 32, dtype=np.uint8)
    for j in range(12):
        sel = wl - 1 > j
        out[start[sel] + j] = mat[idx[sel], j]
    # the separator becomes a newline where the running column crosses a multiple of 72
    line = end // 72
    brk = np.empty(keep, dtype=bool)
    brk[0] = line[0] > 0
    brk[1:] = line[1:] > line[:-1]
    out[end[brk] - 1] = 10
    return out[:n]


CHUNK = 8 << 20  # text is generated in independent 8 MiB chunks so that ranges can be produced alone


def text_range(first, n, seed=SEED0, threads=None):
    """Bytes [first, first+n) of the endless text stream of `seed` (chunk-parallel)."""
    if n == 0:
        return b""
    from concurrent.futures import ThreadPoolExecutor
    import os
    c0, c1 = first // CHUNK, (first + n - 1) // CHUNK
    _vocab_arrays()
    ids = list(range(c0, c1 + 1))
    workers = threads or min(len(ids), os.cpu_count() or 1)
    # a chunk's bytes do not depend on how many of them are generated (the random stream and the word layout
    # are prefix-stable), so the last chunk is only made as long as the range needs
    need = {k: CHUNK for k in ids}
    need[c1] = first + n - c1 * CHUNK
    if workers > 1:
        with ThreadPoolExecutor(workers) as ex:
            parts = list(ex.map(lambda k: _text_chunk(need[k], seed + 0x9E37 * k), ids))
    else:
        parts = [_text_chunk(need[k], seed + 0x9E37 * k) for k in ids]
    whole = np.concatenate(parts) if len(parts) > 1 else parts[0]
    lo = first - c0 * CHUNK
    return whole[lo:lo + n].tobytes()


def text(n, seed=SEED0):
    """n bytes of word-sampled text: Zipf(s=1.1) over a fixed 4096-word lowercase vocabulary
    (word length 2..12, English letter frequencies), separated by spaces, newline every ~72 chars."""
    return text_range(0, n, seed)


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


def text_stream(n, seed=SEED0, first=0):
    """numpy uint8 view of text_range, for bench.py (no per-block slicing copies)."""
    return np.frombuffer(text_range(first, n, seed), dtype=np.uint8)


def mixed_stream(n_blocks, block_bytes, seed=SEED0 + 3):
    """cfg 3 for the bench, generated in one go: block k is text if k%4 in {0,1} (consecutive slices of
    the text stream of `seed`), uniform random bytes if k%4==2 (seed + k), structured binary if k%4==3
    (seed + k).  Returns one uint8 array of n_blocks * block_bytes bytes."""
    out = np.empty(n_blocks * block_bytes, dtype=np.uint8)
    n_text = sum(1 for k in range(n_blocks) if k % 4 < 2)
    txt = np.frombuffer(text_range(0, max(n_text, 1) * block_bytes, seed), dtype=np.uint8)
    t = 0
    for k in range(n_blocks):
        dst = out[k * block_bytes:(k + 1) * block_bytes]
        if k % 4 < 2:
            dst[:] = txt[t * block_bytes:(t + 1) * block_bytes]
            t += 1
        elif k % 4 == 2:
            dst[:] = np.frombuffer(random_bytes(block_bytes, seed + k), dtype=np.uint8)
        else:
            dst[:] = np.frombuffer(structured(block_bytes, seed + k), dtype=np.uint8)
    return out


def file_tree(n_files, seed=SEED0 + 5):
    """cfg 5: n_files files, sizes log-uniform 1 KiB..1 MiB, 30 % exact duplicates of an earlier file,
    text/binary 70/30.  Returns (names, [bytes...]) in sorted path order."""
    r = _xorshift_stream(seed, n_files * 3)
    u = (r[:n_files] >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    sizes = np.exp(np.log(1024) + u * (np.log(1 << 20) - np.log(1024))).astype(np.int64)
    files = []
    pool = text(256 << 20, SEED0 + 55)
    for k in range(n_files):
        if k > 10 and int(r[n_files + k] % np.uint64(10)) < 3:
            files.append(files[int(r[2 * n_files + k] % np.uint64(k))])
        elif int(r[n_files + k] % np.uint64(100)) < 70:
            at = int(r[2 * n_files + k] % np.uint64((256 << 20) - (1 << 20)))
            files.append(pool[at:at + int(sizes[k])])
        else:
            files.append(structured(int(sizes[k]), SEED0 + k))
    names = ["dir%02d/file%05d" % (k % 37, k) for k in range(n_files)]
    return names, files
