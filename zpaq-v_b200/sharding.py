"""Multi-GPU partitioning of a batch of independent ZPAQ blocks.

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
