"""Paged hash tables (ZPAQGPU_TABLES_PAGED): same bytes as the dense layout and the oracle, a block
costs what it touches, and a pool that runs dry falls back to dense waves."""
import pytest

import datagen
import oracle_binding as ob

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _reset(gpu_ctx):
    yield
    gpu_ctx.set_table_mode(0)
    gpu_ctx.set_workspace_limit(0)
    gpu_ctx.set_kernel(0)


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5])
def test_paged_matches_oracle(gpu_ctx, level):
    gpu_ctx.set_table_mode(2)
    blocks = [datagen.mixed_block(k, 60000) for k in range(6)] + [b"", b"Hello World!"]
    comments = ["%d bytes" % len(b) for b in blocks]
    got = gpu_ctx.compress_blocks(level, blocks, comments=comments)
    st = gpu_ctx.stats()
    assert st["paged"] == 1 and st["waves"] == 1 and st["pool_bytes_used"] > 0
    for b, c, g in zip(blocks, comments, got):
        assert g == ob.compress_block(level, b, "", c), level
    plain, segs, status = gpu_ctx.decompress_archive(b"".join(got))
    st = gpu_ctx.stats()
    assert st["paged"] == 1
    assert status == 0 and plain == b"".join(blocks) and all(s["sha1_ok"] == 1 for s in segs)


def test_paged_m5_blocks_resident_together(gpu_ctx):
    """-m5 dense tables are 2 GiB per block; paged, 48 text blocks take a few MB each and run as one wave."""
    gpu_ctx.set_workspace_limit(8 << 30)     # dense would need 12 waves of 4 blocks
    blocks = [datagen.text(40000, datagen.SEED0 + k) for k in range(48)]
    got = gpu_ctx.compress_blocks(5, blocks)
    st = gpu_ctx.stats()
    assert st["paged"] == 1 and st["waves"] == 1
    assert st["pool_bytes_used"] < 48 * (64 << 20)
    for k in (0, 17, 47):
        assert got[k] == ob.compress_block(5, blocks[k], "", "")
    plain, segs, status = gpu_ctx.decompress_archive(b"".join(got))
    assert status == 0 and plain == b"".join(blocks) and gpu_ctx.stats()["waves"] == 1


def test_pool_overflow_falls_back_to_dense(gpu_ctx):
    """Random data touches nearly every line: with a pool of a few MiB the paged attempt overflows and
    the call repeats itself with dense waves -- output still identical to the oracle."""
    gpu_ctx.set_table_mode(2)
    gpu_ctx.set_workspace_limit(13 << 20)    # one dense m2 block (12.07 MiB) or a ~12 MiB page pool
    blocks = [datagen.random_bytes(65536, datagen.SEED0 + k) for k in range(4)]
    got = gpu_ctx.compress_blocks(2, blocks)
    st = gpu_ctx.stats()
    assert st["retries"] >= 1 and st["paged"] == 0 and st["waves"] >= 4
    assert got == [ob.compress_block(2, b, "", "") for b in blocks]
    plain, segs, status = gpu_ctx.decompress_archive(b"".join(got))
    st = gpu_ctx.stats()
    assert status == 0 and plain == b"".join(blocks) and st["retries"] >= 1 and st["paged"] == 0
