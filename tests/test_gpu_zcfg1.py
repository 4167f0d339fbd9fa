"""BASELINE.json configs[0] on the GPU: -m1 round trip of 16 MiB synthetic text as ONE block (the
reference's own CPU-runnable case, `Compressor.start_block(1) ... end_block` over a whole file as
cmd/main.v:288-317 does it).  One block is one chain, so this is the slowest shape the device can be given
(about a minute); it is here for the sizes, not the speed: 2^27 coded bits in one arithmetic-coder stream,
a 16 MiB plaintext segment, offsets past 2^24 everywhere.  The file sorts last so that `-x` runs reach it
after everything else."""
import hashlib

import pytest

import datagen
import oracle_binding as ob

pytestmark = pytest.mark.gpu

MIB = 1 << 20


def test_cfg1_m1_16mib_text_single_block(gpu_ctx):
    gpu_ctx.set_table_mode(0)
    gpu_ctx.set_workspace_limit(0)
    gpu_ctx.set_kernel(0)
    data = datagen.text(16 * MIB)
    name, comment = "cfg1.txt", "%d bytes" % len(data)
    got = gpu_ctx.compress_blocks(1, [data], names=[name], comments=[comment])
    assert gpu_ctx.stats()["kernel"] == 2
    want = ob.compress_block(1, data, name, comment)
    assert len(got) == 1 and len(got[0]) == len(want)
    assert got[0] == want, "16 MiB -m1 block differs from the oracle"
    assert got[0][-22:-1] == b"\xfd" + hashlib.sha1(data).digest()   # segment trailer (compressor.v:380-395)
    plain, segs, status = gpu_ctx.decompress_archive(got[0])
    assert status == 0 and len(segs) == 1 and segs[0]["sha1_ok"] == 1
    assert segs[0]["filename"] == name and segs[0]["comment"] == comment
    assert plain == data
