"""The streaming-shaped calls with a queue (zpaqgpu_block_end_queue / zpaqgpu_flush): the shape of the
reference CLI (cmd/main.v:288-317, one start_block..end_block per file).  Queued blocks are coded in one
launch; bytes and order equal the per-block calls and the oracle.  A plain-C caller feeds 1 000 files
through block_begin .. block_end_queue and must reach at least half of the batch call's MB/s."""
import os
import subprocess

import pytest

import datagen
import oracle_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def feed(ctx, level, segs, header=None):
    assert ctx.block_begin(level=level, header=header) == 0
    for name, comment, data in segs:
        assert ctx.segment_begin(name, comment) == 0
        if data is not None:
            for at in range(0, max(len(data), 1), 5000):
                assert ctx.segment_write(data[at:at + 5000]) == 0
        assert ctx.segment_end() == 0


def test_queue_equals_per_block_calls_and_oracle(gpu_ctx):
    blocks = [
        (2, [("a.txt", "6000 bytes", datagen.text(6000))]),
        (1, [("b.bin", "", datagen.random_bytes(3000))]),
        (2, [("c1", "x", datagen.text(2500, 7)), ("c2", "y", datagen.structured(4000)), ("c3", "", b"")]),  # Q17
        (2, []),                                                       # a block without segments
        (0, [("stored", "", datagen.random_bytes(70000))]),            # store mode, chunked
        (3, [("never-compressed", "", None)]),                         # Q16: no compress() call, no PP byte
        (2, [("a.txt", "6000 bytes", datagen.text(6000))]),
    ]
    single = []
    for level, segs in blocks:
        feed(gpu_ctx, level, segs)
        single.append(gpu_ctx.block_end())
    for level, segs in blocks:
        feed(gpu_ctx, level, segs)
        assert gpu_ctx.block_end_queue() == 0
    assert gpu_ctx.queued()[0] == len(blocks)
    # a block may not be coded on its own while others wait (its bytes would overtake theirs): the call is
    # refused with ZPAQGPU_E_STATE -- swallowed by the binding like the reference swallows wrong-state calls --
    # and the block stays open; queued instead, it is delivered last
    feed(gpu_ctx, 1, [("late", "", b"zz")])
    assert gpu_ctx.block_end() is None
    assert gpu_ctx.block_end_queue() == 0
    whole = gpu_ctx.flush()
    assert gpu_ctx.queued() == (0, 0)
    assert whole == b"".join(single) + ob.compress_block(1, b"zz", "late", "")
    assert single[0] == ob.compress_block(2, datagen.text(6000), "a.txt", "6000 bytes")
    assert single[1] == ob.compress_block(1, datagen.random_bytes(3000), "b.bin", "")
    assert single[4] == ob.compress_block(0, datagen.random_bytes(70000), "stored", "")
    plain, segs, status = gpu_ctx.decompress_archive(whole)
    assert status == 0 and all(s["sha1_ok"] == 1 for s in segs)
    assert [s["filename"] for s in segs] == ["a.txt", "b.bin", "c1", "c2", "c3", "stored", "never-compressed", "a.txt",
                                             "late"]
    assert gpu_ctx.flush() == b""


def test_queue_limits_and_mirror(gpu_ctx):
    import zpaq_v_b200 as z
    gpu_ctx.stream_batch(3, 0)
    try:
        files = [("f%d" % k, datagen.text(3000 + 100 * k, 40 + k)) for k in range(7)]
        w = z.FileWriter()
        c = z.Compressor(gpu_ctx, batch=True)
        c.set_output(w)
        for name, data in files:            # cmd/main.v:298-311
            c.set_input(z.FileReader(data))
            c.start_block(2)
            c.start_segment(name, "%d bytes" % len(data))
            while c.compress(65536):
                pass
            c.end_segment()
            c.end_block()
        assert gpu_ctx.queued()[0] == 1     # two full queues of three were delivered on the way
        c.flush()                           # the one line the CLI adds before cmd/main.v:320
        assert w.bytes() == b"".join(ob.compress_block(2, d, n, "%d bytes" % len(d)) for n, d in files)
    finally:
        gpu_ctx.stream_batch(0, 0)


def test_c_caller_1000_files_queue_vs_batch(tmp_path):
    exe = str(tmp_path / "stream_batch")
    pkg = os.path.join(ROOT, "zpaq-v_b200")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "stream_batch.c"),
                           "-o", exe, "-L" + pkg, "-lzpaqgpu", "-Wl,-rpath," + pkg])
    r = subprocess.run([exe, "1000", "128", "2", "4"], capture_output=True, timeout=900)
    assert r.returncode == 0, r.stderr.decode()
    n, total, ms_batch, ms_queued, eq_ab, k, ms_single, eq_ac = r.stdout.decode().split()
    assert int(eq_ab) == 1 and int(eq_ac) == 1
    mb = int(total) / 1e6
    batch, queued = mb / (float(ms_batch) / 1e3), mb / (float(ms_queued) / 1e3)
    print("batch %.1f MB/s, queued streaming calls %.1f MB/s, one launch per block: %.1f ms per block"
          % (batch, queued, float(ms_single) / int(k)))
    assert queued >= 0.5 * batch
