"""Race hunting without compute-sanitizer (the GPU pool refuses it: profiles/r02_sanitizer_refused.txt).

The kernels that could hide a race are the ones whose warps exchange data: the three-warp encoder (named
barrier + shared rings), the tree decoder (table updates handed down between lanes, shared stage and rings)
and the paged tables (pages mapped by atomics from a pool shared by every block of the wave).  A race shows
up as a result that depends on timing, so the same blocks are coded under different packings -- 1, 5 and 7
blocks per CTA, one wave and several, dense and paged tables, twice each -- and every block of every run must
equal the CPU oracle's bytes.  (-m4 runs paged: its dense tables are 385 MiB per block.)"""
import pytest

import datagen
from test_gpu_fullsize import oracle_blocks_mt

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _reset(gpu_ctx):
    yield
    gpu_ctx.set_table_mode(0)
    gpu_ctx.set_workspace_limit(0)


def _blocks(n):
    """n blocks of 1.5-4 KB: text, text, random, structured, ... cut from three buffers made once (a
    datagen call builds an 8 MiB chunk, so one call per block would cost minutes)."""
    sizes = [1500 + (k * 397) % 2600 for k in range(n)]
    total = sum(sizes)
    src = [datagen.text(total, datagen.SEED0 + 500), datagen.random_bytes(total, datagen.SEED0 + 501),
           datagen.structured(total, datagen.SEED0 + 502)]
    out, at = [], 0
    for k, size in enumerate(sizes):
        out.append(src[(0, 0, 1, 2)[k % 4]][at:at + size])
        at += size
    return out


@pytest.mark.parametrize("level", [1, 2, 4])
def test_same_bytes_under_every_packing(gpu_ctx, level):
    # (the oracle sets up 385 MiB of tables per block at -m4: fewer blocks there)
    top = 1040 if level < 4 else 300
    blocks = _blocks(top)
    want = None
    # 1 block per CTA; 3 per CTA twice; paged; and, where the tables allow, 7 per CTA and dense waves of what 2 GiB hold
    packings = [(37, 0, 0, 1), (300, 0, 0, 2), (300, 2, 0, 1)]
    if level < 4:
        packings += [(1040, 0, 0, 1), (300, 1, 2, 1)]
    for n, mode, limit_gib, reps in packings:
        gpu_ctx.set_table_mode(mode)
        gpu_ctx.set_workspace_limit(limit_gib << 30)
        for rep in range(reps):
            got = gpu_ctx.compress_blocks(level, blocks[:n], comments=["%d bytes" % len(b) for b in blocks[:n]])
            if want is None:
                want = oracle_blocks_mt(level, blocks)   # every block, all host threads (table set-up dominates)
            assert got == want[:n], (level, n, mode, limit_gib, rep)
            plain, segs, status = gpu_ctx.decompress_archive(b"".join(got))
            assert status == 0 and plain == b"".join(blocks[:n]) and all(s["sha1_ok"] == 1 for s in segs), (level, n, mode)
