"""GPU parity at the block sizes of BASELINE.json: 1 MiB blocks at -m1..-m4 (cfg 2 text, cfg 3 mixed text /
random / structured) and 4 MiB blocks at -m5 with paged tables (cfg 4).  EVERY block of a batch is
compared byte for byte with the CPU oracle (run block-parallel on all host threads), then decoded on
the GPU and compared with the input and the stored SHA-1."""
import ctypes as C
import os

import numpy as np
import pytest

import datagen
import oracle_binding as ob

pytestmark = pytest.mark.gpu

MIB = 1 << 20


@pytest.fixture(autouse=True)
def _reset(gpu_ctx):
    gpu_ctx.set_table_mode(0)
    gpu_ctx.set_workspace_limit(0)
    gpu_ctx.set_kernel(0)
    yield
    gpu_ctx.set_table_mode(0)
    gpu_ctx.set_workspace_limit(0)
    gpu_ctx.set_kernel(0)


def oracle_blocks_mt(level, blocks):
    """Every block through the oracle (empty name, comment "<n> bytes"), all host threads."""
    L = ob.lib()
    n = len(blocks)
    src = np.frombuffer(b"".join(blocks), dtype=np.uint8)
    off = (C.c_uint64 * (n + 1))()
    pos = 0
    for i, b in enumerate(blocks):
        off[i] = pos
        pos += len(b)
    off[n] = pos
    cap = pos + pos // 2 + 4096 * n
    out = np.empty(cap, dtype=np.uint8)
    out_off = (C.c_uint64 * (n + 1))()
    need = C.c_uint64(0)
    rc = L.zo_compress_blocks_mt(level, src.ctypes.data, off, n, out.ctypes.data, cap, out_off, C.byref(need),
                                 os.cpu_count() or 1)
    assert rc == 0
    return [bytes(out[out_off[k]:out_off[k + 1]]) for k in range(n)]


def check(gpu_ctx, level, blocks, paged=None):
    comments = ["%d bytes" % len(b) for b in blocks]
    got = gpu_ctx.compress_blocks(level, blocks, comments=comments)
    st = gpu_ctx.stats()
    assert st["kernel"] == 2
    if paged is not None:
        assert st["paged"] == int(paged)
    want = oracle_blocks_mt(level, blocks)
    for k, (g, w) in enumerate(zip(got, want)):
        assert g == w, "level %d block %d differs from the oracle" % (level, k)
    plain, segs, status = gpu_ctx.decompress_archive(b"".join(got))
    assert status == 0 and len(segs) == len(blocks) and all(s["sha1_ok"] == 1 for s in segs)
    assert plain == b"".join(blocks)
    return st


def test_cfg2_m2_1mib_text_all_blocks(gpu_ctx):
    whole = datagen.text(16 * MIB)
    check(gpu_ctx, 2, [whole[i * MIB:(i + 1) * MIB] for i in range(16)])


def test_cfg3_m3_1mib_mixed_all_blocks(gpu_ctx):
    """text, text, random, structured, ... -- the random blocks expand and touch every table line"""
    check(gpu_ctx, 3, [datagen.mixed_block(k, MIB) for k in range(16)])


@pytest.mark.parametrize("level", [1, 4])
def test_m1_m4_1mib_mixed_all_blocks(gpu_ctx, level):
    check(gpu_ctx, level, [datagen.mixed_block(k, MIB) for k in range(8)])


def test_cfg4_m5_4mib_text_paged_all_blocks(gpu_ctx):
    """cfg 4's shape: 2 GiB of dense tables per block, so the batch runs on paged tables"""
    gpu_ctx.set_table_mode(2)
    whole = datagen.text(8 * 4 * MIB, datagen.SEED0 + 4)
    st = check(gpu_ctx, 5, [whole[i * 4 * MIB:(i + 1) * 4 * MIB] for i in range(8)], paged=True)
    assert st["waves"] == 1


def test_m3_1mib_mixed_paged_equals_dense(gpu_ctx):
    """random blocks on paged tables at 1 MiB: every line of the 64 MiB of hash tables gets mapped"""
    blocks = [datagen.mixed_block(k, MIB) for k in range(4)]
    gpu_ctx.set_table_mode(2)
    check(gpu_ctx, 3, blocks, paged=True)
