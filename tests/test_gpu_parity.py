"""GPU parity: libzpaqgpu (through the C ABI) against the CPU oracle, byte for byte.

Bar (BASELINE.json north_star): compressed streams byte-identical to the reference for the same
method and block; decompression bit-exact with matching SHA1.
"""
import pytest

import datagen
import oracle_binding as ob

pytestmark = pytest.mark.gpu

KERNELS = {"generic": 1, "chain": 2}


def small_inputs():
    rep = b"".join(b"line %03d: the quick brown fox jumps over the lazy dog\n" % (i % 7) for i in range(100))
    return {
        "empty": b"",
        "one": b"A",
        "hello": b"Hello World!",
        "aaaabbbb": b"AAAABBBB",
        "zeros8k": bytes(8192),
        "rand4k": datagen.random_bytes(4096),
        "replines": rep,
        "text20k": datagen.text(20000),
        "ff": b"\xff" * 300,
    }


@pytest.mark.parametrize("kernel", ["generic", "chain"])
@pytest.mark.parametrize("level", [1, 2, 3, 4, 5])
def test_compress_matches_oracle_small(gpu_ctx, level, kernel):
    gpu_ctx.set_kernel(KERNELS[kernel])
    gpu_ctx.set_workspace_limit(6 << 30)
    inputs = small_inputs()
    names = list(inputs)
    blocks = [inputs[k] for k in names]
    comments = ["%d bytes" % len(b) for b in blocks]
    got = gpu_ctx.compress_blocks(level, blocks, names=names, comments=comments)
    for name, data, comment, g in zip(names, blocks, comments, got):
        want = ob.compress_block(level, data, name, comment)
        assert g == want, "level %d %s: %d vs %d bytes" % (level, name, len(g), len(want))
    gpu_ctx.set_kernel(0)


@pytest.mark.parametrize("kernel", ["generic", "chain"])
@pytest.mark.parametrize("level", [1, 2, 3, 4, 5])
def test_decompress_oracle_archive_small(gpu_ctx, level, kernel):
    gpu_ctx.set_kernel(KERNELS[kernel])
    gpu_ctx.set_workspace_limit(6 << 30)
    inputs = small_inputs()
    arc = b"".join(ob.compress_block(level, d, n, "%d bytes" % len(d)) for n, d in inputs.items())
    plain, segs, status = gpu_ctx.decompress_archive(arc)
    assert status == 0
    assert [s["filename"] for s in segs] == list(inputs)
    assert plain == b"".join(inputs.values())
    assert all(s["sha1_ok"] == 1 for s in segs)
    gpu_ctx.set_kernel(0)


def test_store_level0(gpu_ctx):
    inputs = small_inputs()
    inputs["big"] = datagen.text(200000)
    inputs["edge65535"] = datagen.random_bytes(65535)
    inputs["edge65536"] = datagen.random_bytes(65536)
    names = list(inputs)
    blocks = [inputs[k] for k in names]
    got = gpu_ctx.compress_blocks(0, blocks, names=names)
    for name, data, g in zip(names, blocks, got):
        assert g == ob.compress_block(0, data, name, ""), name
    plain, segs, status = gpu_ctx.decompress_archive(b"".join(got))
    assert status == 0 and plain == b"".join(blocks)
    assert all(s["sha1_ok"] == 1 for s in segs)
