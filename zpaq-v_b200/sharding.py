"""Multi-GPU partitioning of a batch of independent ZPAQ blocks, and of
`jidac add` (the one path with a real exchange step: duplicates across ranks need the fragment
digests of every rank).

Blocks share nothing (compressor.v:84-187 re-initialises every piece of model state in
start_block), so the data path needs no collective: each rank (one process per GPU) codes a
contiguous range of blocks and the host concatenates the results in block order.  The only
communication is the gather of finished bytes on the host side (torch.distributed object gather;
gloo on CPU in the tests, any backend in production).
"""


def shard_range(n_blocks, world, rank):
    """Contiguous, near-equal block range [lo, hi) of `rank`."""
    base, extra = divmod(n_blocks, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_bytes(sizes, world):
    """Contiguous ranges balanced by input bytes (block = file, cmd/main.v:288-317): returns
    world+1 boundaries.  Greedy prefix split at multiples of total/world."""
    total = sum(sizes)
    bounds = [0]
    acc = 0
    k = 1
    for i, s in enumerate(sizes):
        acc += s
        while k < world and acc >= total * k / world:
            bounds.append(i + 1)
            k += 1
    while len(bounds) < world:
        bounds.append(len(sizes))
    bounds.append(len(sizes))
    return bounds


def compress_sharded(compress_fn, blocks, names, comments, world, rank, bounds=None):
    """Compress this rank's range with compress_fn(blocks, names, comments) -> list of bytes."""
    if bounds is None:
        lo, hi = shard_range(len(blocks), world, rank)
    else:
        lo, hi = bounds[rank], bounds[rank + 1]
    return lo, compress_fn(blocks[lo:hi], names[lo:hi] if names else None, comments[lo:hi] if comments else None)


def gather_in_block_order(dist, local_first, local_parts, dst=0):
    """Gather every rank's finished blocks on `dst` and return them in block order (others: None)."""
    world = dist.get_world_size()
    bucket = [None] * world if dist.get_rank() == dst else None
    dist.gather_object((local_first, local_parts), bucket, dst=dst)
    if bucket is None:
        return None
    bucket.sort(key=lambda t: t[0])
    out = []
    for _, parts in bucket:
        out.extend(parts)
    return out


# ---- jidac add over several ranks -------------------------------------------------------------
def _jidac_name(date, kind, num):
    """make_jidac_filename, jidac.v:47-49"""
    return "jDC%s%s%s" % (str(date).rjust(14, "0"), kind, str(num).rjust(10, "0"))


def jidac_add_sharded(dist, fragment_fn, compress_fn, names, files, date, level=1, fragment=6, block_bytes=1 << 20,
                      dst=0):
    """`jidac add` of `files` (every rank passes the same lists) over dist.get_world_size() ranks.

    1. every rank cuts and hashes the files of its byte-balanced range:
       fragment_fn(files) -> [{off, len, file, sha1}, ...] (Context.jidac_fragment: rolling hash and
       SHA-1 kernels on the rank's GPU);
    2. EXCHANGE: the (sha1, len) lists are all-gathered -- 24 bytes per fragment -- and every rank
       derives the same global table: ids count first occurrences in file order (jidac.v:153-163), a
       later equal fragment, on whichever rank, takes the id of the first;
    3. every rank packs the fragments it stores (first occurrences inside its range) into d blocks of
       up to block_bytes and codes them: compress_fn(level, blocks, names, comments) -> [bytes, ...];
       a d block never spans ranks, so the block cut depends on the number of ranks, the contents of
       the archive do not;
    4. the d blocks are gathered on `dst`, which writes c block, d blocks, one h block per d block and
       the i block (jidac.v:216-295) -- the index blocks through compress_fn(0, ...).
    Returns the archive on `dst`, None elsewhere.  With one rank the bytes equal zpaqgpu_jidac_add's.
    """
    import struct
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = shard_by_bytes([len(f) for f in files], world)
    lo, hi = bounds[rank], bounds[rank + 1]
    mine = fragment_fn(files[lo:hi])
    local = [(f["file"] + lo, f["off"], f["len"], bytes(f["sha1"])) for f in mine]
    # 2. the exchange step
    table = [None] * world
    dist.all_gather_object(table, [(fi, ln, sh) for fi, _, ln, sh in local])
    ids, seen, frag_id, owner_of = [], {}, 0, []
    for r, part in enumerate(table):
        for fi, ln, sh in part:
            key = (sh, ln)
            if key not in seen:
                frag_id += 1
                seen[key] = frag_id
                owner_of.append(r)
            ids.append((fi, seen[key], key))
    n_stored = frag_id
    # 3. this rank's d blocks
    blob = b"".join(bytes(f) for f in files[lo:hi])
    first_seen = set()
    base = sum(len(t) for t in table[:rank])
    blocks, metas, cur, cur_meta = [], [], [], []
    for k, (fi, off, ln, sh) in enumerate(local):
        fid = ids[base + k][1]
        if owner_of[fid - 1] != rank or fid in first_seen:
            continue
        first_seen.add(fid)
        if cur_meta and (block_bytes == 0 or sum(m[2] for m in cur_meta) + ln > block_bytes):
            blocks.append(b"".join(cur)), metas.append(cur_meta)
            cur, cur_meta = [], []
        cur.append(blob[off:off + ln])
        cur_meta.append((fid, sh, ln))
    if cur_meta:
        blocks.append(b"".join(cur)), metas.append(cur_meta)
    d_names = [_jidac_name(date, "d", m[0][0]) for m in metas]
    d_comments = ["%d jDC\x01" % len(b) for b in blocks]
    coded = compress_fn(level, blocks, d_names, d_comments) if blocks else []
    # 4. gather and index
    bucket = [None] * world if rank == dst else None
    dist.gather_object((rank, coded, metas), bucket, dst=dst)
    if rank != dst:
        return None
    bucket.sort(key=lambda t: t[0])
    d_all, h_plain, h_names = [], [], []
    for _, cd, ms in bucket:
        for blk, meta in zip(cd, ms):
            d_all.append(blk)
            h_plain.append(struct.pack("<I", len(blk)) + b"".join(sh + struct.pack("<I", ln) for _, sh, ln in meta))
            h_names.append(_jidac_name(date, "h", meta[0][0]))
    i_plain = bytearray()
    at = 0
    for fi, nm in enumerate(names):
        i_plain += struct.pack("<q", date) + (nm.encode() if isinstance(nm, str) else nm) + b"\0"
        ptr = []
        while at < len(ids) and ids[at][0] == fi:
            ptr.append(ids[at][1])
            at += 1
        if date != 0:
            i_plain += struct.pack("<II", 0, len(ptr)) + b"".join(struct.pack("<I", p) for p in ptr)
    small = [struct.pack("<q", sum(len(b) for b in d_all))] + h_plain + ([bytes(i_plain)] if i_plain else [])
    small_names = [_jidac_name(date, "c", n_stored + 1)] + h_names + ([_jidac_name(date, "i", 1)] if i_plain else [])
    coded_small = compress_fn(0, small, small_names, ["%d jDC\x01" % len(b) for b in small])
    return coded_small[0] + b"".join(d_all) + b"".join(coded_small[1:])
