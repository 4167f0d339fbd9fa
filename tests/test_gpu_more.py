"""GPU parity, part 2: generic components, multi-segment blocks, the locator scan, mixed archives,
larger blocks, waves, the device-pointer API, error paths and the mirrored Compressor/Decompresser
classes -- all through the C ABI, all against the CPU oracle."""
import ctypes as C
import hashlib

import pytest

import datagen
import oracle_binding as ob
from test_oracle_kats import CUSTOM_HEADERS, ci_files

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _reset(gpu_ctx):
    gpu_ctx.set_kernel(0)
    gpu_ctx.set_workspace_limit(0)
    yield
    gpu_ctx.set_kernel(0)
    gpu_ctx.set_workspace_limit(0)


@pytest.mark.parametrize("name", sorted(CUSTOM_HEADERS))
def test_generic_components_match_oracle(gpu_ctx, name):
    """CONS/CM/ICM/MATCH/AVG/MIX2/MIX/ISSE/SSE and a branching HCOMP program."""
    hdr = bytes(CUSTOM_HEADERS[name])
    blocks = [b"", b"x", b"abracadabra " * 50, datagen.text(6000), datagen.random_bytes(1500), bytes(3000),
              datagen.structured(5000)]
    names = ["b%d" % i for i in range(len(blocks))]
    got = gpu_ctx.compress_blocks(0, blocks, names=names, header=hdr)
    for n, b, g in zip(names, blocks, got):
        assert g == ob.compress_block(0, b, n, "", header=hdr), (name, n)
    assert gpu_ctx.stats()["kernel"] == 1
    plain, segs, status = gpu_ctx.decompress_archive(b"".join(got))
    assert status == 0 and plain == b"".join(blocks) and all(s["sha1_ok"] == 1 for s in segs)


@pytest.mark.parametrize("name", ["icm_match_mix2_sse", "forward_refs", "twenty", "vm_branches"])
def test_generic_warp_kernel_equals_one_lane_kernel(name, monkeypatch):
    """The warp program (component per lane, levels, MIX as a warp dot product, warp-uniform ZPAQL) against
    the one-lane kernel it replaces and the oracle: multi-segment blocks included (tables persist, Q17)."""
    import zpaq_v_b200 as z
    hdr = bytes(CUSTOM_HEADERS[name])
    blocks = [datagen.text(20000, 91), datagen.structured(9000), datagen.random_bytes(3000), b"", b"a" * 5000]
    want = [ob.compress_block(0, b, "n%d" % k, "c", header=hdr) for k, b in enumerate(blocks)]
    got = {}
    for variant in ("warp", "lane0"):
        if variant == "lane0":
            monkeypatch.setenv("ZPAQGPU_GENERIC", "lane0")
        else:
            monkeypatch.delenv("ZPAQGPU_GENERIC", raising=False)
        ctx = z.Context()
        try:
            got[variant] = ctx.compress_blocks(0, blocks, names=["n%d" % k for k in range(len(blocks))],
                                               comments=["c"] * len(blocks), header=hdr)
            assert ctx.stats()["kernel"] == 1
            plain, segs, status = ctx.decompress_archive(b"".join(want))
            assert status == 0 and plain == b"".join(blocks) and all(s["sha1_ok"] == 1 for s in segs)
            # two segments in one block: component tables, M, H and the VM registers carry over
            assert ctx.block_begin(header=hdr) == 0
            for nm, data in (("s1", blocks[0][:7000]), ("s2", blocks[1][:5000])):
                assert ctx.segment_begin(nm, "") == 0 and ctx.segment_write(data) == 0 and ctx.segment_end() == 0
            two = ctx.block_end()
            c = ob.Compressor()
            c.start_block_header(hdr)
            for nm, data in (("s1", blocks[0][:7000]), ("s2", blocks[1][:5000])):
                c.set_input(data)
                c.start_segment(nm, "")
                while c.compress(65536):
                    pass
                c.end_segment()
            c.end_block()
            assert two == c.output()
        finally:
            ctx.close()
    assert got["warp"] == want and got["lane0"] == want


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("level", [0, 1, 2, 4])
def test_multisegment_block_streaming_api(gpu_ctx, level, kernel):
    """Q16/Q17 through zpaqgpu_block_begin/segment_*/block_end: tables persist across segments, the
    PP byte is coded only if compress() was called."""
    if level == 0 and kernel == 2:
        pytest.skip("store mode has no model")
    gpu_ctx.set_kernel(kernel)
    gpu_ctx.set_workspace_limit(6 << 30)
    parts = [datagen.text(5000), b"", datagen.text(7000, datagen.SEED0 + 9), datagen.random_bytes(700)]
    called = [True, True, True, False]   # the last segment never sees compress()
    c = ob.Compressor()
    c.start_block(level)
    assert gpu_ctx.block_begin(level=level) == 0
    for i, (d, cl) in enumerate(zip(parts, called)):
        c.set_input(d if cl else b"")
        c.start_segment("s%d" % i, "c%d" % i)
        assert gpu_ctx.segment_begin("s%d" % i, "c%d" % i) == 0
        if cl:
            while c.compress(1000):
                pass
            for k in range(0, max(len(d), 1), 1000):
                assert gpu_ctx.segment_write(d[k:k + 1000]) == 0
        c.end_segment()
        assert gpu_ctx.segment_end() == 0
    c.end_block()
    got = gpu_ctx.block_end()
    assert got == c.output()
    plain, segs, status = gpu_ctx.decompress_archive(got)
    want = b"".join(d for d, cl in zip(parts, called) if cl)
    assert status == 0 and plain == want and [s["filename"] for s in segs] == ["s0", "s1", "s2", "s3"]
    assert all(s["sha1_ok"] == 1 for s in segs) and len({s["block_index"] for s in segs}) == 1


def test_streaming_state_errors(gpu_ctx):
    """The reference silently ignores calls in the wrong state; the C ABI reports E_STATE."""
    assert gpu_ctx.segment_begin("a", "b") == -7
    assert gpu_ctx.segment_write(b"x") == -7
    assert gpu_ctx.block_begin(level=1) == 0
    assert gpu_ctx.block_begin(level=2) == -7
    assert gpu_ctx.block_end() == ob.compress_block(1, b"", "", "")[:16 + 4 + 26] + b"\xff"


def test_find_blocks_scan(gpu_ctx):
    blocks = [ob.compress_block(1, datagen.text(3000), "a", ""), ob.compress_block(0, b"stored", "b", ""),
              ob.compress_block(2, datagen.random_bytes(900), "c", "")]
    junk = datagen.random_bytes(777)
    arc = junk + blocks[0] + b"\x00" * 5 + blocks[1] + blocks[2] + junk[:100]
    starts = gpu_ctx.find_blocks(arc)
    want = []
    pos = 0
    loc = blocks[0][:16]
    while True:
        pos = arc.find(loc, pos)
        if pos < 0:
            break
        want.append(pos + 16)
        pos += 1
    assert starts == want and len(starts) == 3
    assert gpu_ctx.find_blocks(b"") == [] and gpu_ctx.find_blocks(loc[:15]) == []
    assert gpu_ctx.find_blocks(loc) == [16]
    # an archive stored inside an archive: the inner locators lie inside the outer block and must
    # not be taken for blocks (the reference's scan resumes only after the outer block ends)
    inner = blocks[0] + blocks[2]
    outer = ob.compress_block(0, inner, "nested.zpaq", "") + blocks[1]
    plain, segs, status = gpu_ctx.decompress_archive(outer)
    want_plain, n_segs, bad = ob.decompress_archive(outer)
    assert status == 0 and plain == want_plain == inner + b"stored" and len(segs) == n_segs == 2


def test_mixed_level_archive(gpu_ctx):
    files = ci_files()
    arc = b""
    for i, (n, d) in enumerate(files.items()):
        arc += ob.compress_block(i % 6, d, n, "%d bytes" % len(d))
    arc += ob.compress_block(0, b"tail", "t", "", header=bytes(CUSTOM_HEADERS["icm_cm_mix"]))
    plain, segs, status = gpu_ctx.decompress_archive(arc)
    assert status == 0 and plain == b"".join(files.values()) + b"tail"
    assert [s["filename"] for s in segs] == list(files) + ["t"]
    assert [s["comment"] for s in segs][:5] == ["%d bytes" % len(d) for d in files.values()]


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5])
def test_larger_blocks_match_oracle(gpu_ctx, level):
    n = 262144 if level < 4 else 131072
    blocks = [datagen.mixed_block(k, n) for k in range(4)]
    comments = ["%d bytes" % n] * 4
    got = gpu_ctx.compress_blocks(level, blocks, comments=comments)
    assert gpu_ctx.stats()["kernel"] == 2
    for k, (b, g) in enumerate(zip(blocks, got)):
        assert g == ob.compress_block(level, b, "", comments[k]), (level, k)
    plain, segs, status = gpu_ctx.decompress_archive(b"".join(got))
    assert status == 0 and plain == b"".join(blocks) and all(s["sha1_ok"] == 1 for s in segs)


def test_waves_when_tables_do_not_fit(gpu_ctx):
    """8 blocks with a budget for 3 resident tables: the batch is coded in 3 waves, same bytes."""
    blocks = [datagen.text(20000, datagen.SEED0 + k) for k in range(8)]
    want = [ob.compress_block(2, b, "", "") for b in blocks]
    gpu_ctx.set_workspace_limit(3 * 12656896 + 4096)
    gpu_ctx.set_table_mode(1)  # dense tables: in auto mode a batch that does not fit is paged instead
    try:
        got = gpu_ctx.compress_blocks(2, blocks)
        assert got == want and gpu_ctx.stats()["waves"] == 3
        plain, segs, status = gpu_ctx.decompress_archive(b"".join(got))
        assert status == 0 and plain == b"".join(blocks) and gpu_ctx.stats()["waves"] == 3
    finally:
        gpu_ctx.set_table_mode(0)
        gpu_ctx.set_workspace_limit(0)


def test_full_size_roundtrip_properties(gpu_ctx):
    """BASELINE configs[1] shape at reduced count: 64 x 1 MiB at -m2.  Size-independent checks:
    exact round trip, every SHA1 verified on the device, sizes add up, and the oracle decoder
    accepts a GPU block (the reference's Decompresser must be able to read what we write)."""
    n, bb = 64, 1 << 20
    whole = datagen.text(n * bb)
    blocks = [whole[i * bb:(i + 1) * bb] for i in range(n)]
    got = gpu_ctx.compress_blocks(2, blocks, comments=["%d bytes" % bb] * n)
    assert all(g[:16] == got[0][:16] and g[-1] == 0xFF for g in got)
    arc = b"".join(got)
    plain, segs, status = gpu_ctx.decompress_archive(arc)
    assert status == 0 and len(segs) == n and all(s["sha1_ok"] == 1 for s in segs)
    assert hashlib.sha1(plain).digest() == hashlib.sha1(whole).digest()
    assert sum(s["out_len"] for s in segs) == n * bb
    for s, g in zip(segs, got):
        assert g[-21:-1] == hashlib.sha1(blocks[s["block_index"]]).digest()
    assert ob.decompress_archive(got[17])[0] == blocks[17]
    assert got[5] == ob.compress_block(2, blocks[5], "", "%d bytes" % bb)


def test_truncated_and_corrupt_archives(gpu_ctx):
    blk = ob.compress_block(2, datagen.text(4000), "f", "4000 bytes")
    plain, segs, status = gpu_ctx.decompress_archive(blk[:-30])       # trailer cut off
    # no 0xFF after the segment: the block never ends (ZPAQGPU_E_FORMAT); the plaintext is complete, as
    # it is for the oracle, because the EOF decision does not depend on the four flush bytes
    assert status == -5 and plain == datagen.text(4000) == ob.decompress_archive(blk[:-30])[0]
    bad = bytearray(blk)
    bad[-10] ^= 0x55                                                  # stored SHA1 damaged
    plain, segs, status = gpu_ctx.decompress_archive(bytes(bad))
    assert plain == datagen.text(4000) and segs[0]["sha1_ok"] == 0    # Q10: reported, output unchanged
    hdr_bad = bytearray(blk)
    hdr_bad[16] = 7                                                   # level byte not 1/2: find_block fails
    plain, segs, status = gpu_ctx.decompress_archive(bytes(hdr_bad))
    assert plain == b"" and segs == []


def test_device_pointer_api(gpu_ctx):
    import numpy as np
    import torch
    from zpaq_v_b200 import binding as zb
    L = zb.lib()
    n, bb = 6, 50000
    data = datagen.text(n * bb)
    d_in = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
    cap = n * (bb + bb // 4 + 4096)
    d_arc = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    in_off = (C.c_uint64 * (n + 1))(*[i * bb for i in range(n + 1)])
    tot = C.c_uint64(0)
    gpu_ctx._check(L.zpaqgpu_compress_blocks_dev(gpu_ctx._h, 3, d_in.data_ptr(), None, in_off, n, d_arc.data_ptr(),
                                                 cap, d_off.data_ptr(), C.byref(tot)))
    off = d_off.cpu().tolist()
    arc = bytes(d_arc[:tot.value].cpu().numpy())
    for k in range(n):
        assert arc[off[k]:off[k + 1]] == ob.compress_block(3, data[k * bb:(k + 1) * bb], "", "%d bytes" % bb)
    d_plain = torch.zeros(n * bb, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(n, dtype=torch.int64, device="cuda")
    bad = C.c_int(-1)
    arc_off = (C.c_uint64 * (n + 1))(*off)
    gpu_ctx._check(L.zpaqgpu_decompress_blocks_dev(gpu_ctx._h, d_arc.data_ptr(), arc_off, n, d_plain.data_ptr(), in_off,
                                                   d_len.data_ptr(), C.byref(bad)))
    assert bad.value == 0 and d_len.cpu().tolist() == [bb] * n
    assert bytes(d_plain.cpu().numpy()) == data


def test_output_too_small_reports_need(gpu_ctx):
    from zpaq_v_b200 import binding as zb
    L = zb.lib()
    data = datagen.random_bytes(5000)
    off = (C.c_uint64 * 2)(0, 5000)
    out_off = (C.c_uint64 * 2)()
    need = C.c_uint64(0)
    out = C.create_string_buffer(100)
    rc = L.zpaqgpu_compress_blocks(gpu_ctx._h, 1, data, off, 1, None, None, out, 100, out_off, C.byref(need))
    assert rc == zb.E_NOSPACE and need.value == len(ob.compress_block(1, data, "", ""))


# ---- the mirrored classes, written the way zpaq_test.v:364-384 / cmd/main.v:288-401 drive them ----
def test_mirror_basic_compression(gpu_ctx):
    import zpaq_v_b200 as z
    inp = z.FileReader(bytes([0x41, 0x41, 0x41, 0x41, 0x42, 0x42, 0x42, 0x42]))
    out = z.FileWriter()
    comp = z.Compressor(gpu_ctx)
    comp.set_input(inp)
    comp.set_output(out)
    comp.start_block(1)
    comp.start_segment("test", "")
    while comp.compress(8):
        pass
    comp.end_segment()
    comp.end_block()
    compressed = out.bytes()
    assert len(compressed) > 0
    assert compressed.hex().endswith("fd7cd188ef3a9ea7fa0ee9c62c168709695460f5c0ff")
    assert compressed == ob.compress_block(1, b"AAAABBBB", "test", "")


@pytest.mark.parametrize("level", range(6))
def test_mirror_add_extract_roundtrip(gpu_ctx, level):
    """cmd/main.v run_add / run_extract shape over the mirrored classes."""
    import zpaq_v_b200 as z
    files = ci_files()
    out = z.FileWriter()
    for name, data in files.items():
        comp = z.Compressor(gpu_ctx)
        comp.set_input(z.FileReader(data))
        comp.set_output(out)
        comp.start_block(level)
        comp.start_segment(name.split("/")[-1], "%d bytes" % len(data))
        while comp.compress(65536):
            pass
        comp.end_segment()
        comp.end_block()
    arc = out.bytes()
    assert arc == b"".join(ob.compress_block(level, d, n.split("/")[-1], "%d bytes" % len(d)) for n, d in files.items())
    d = z.Decompresser(gpu_ctx)
    d.set_input(z.FileReader(arc))
    got = {}
    while d.find_block():
        while d.find_filename():
            w = z.FileWriter()
            d.set_output(w)
            while d.decompress(65536):
                pass
            d.read_segment_end()
            got[d.get_filename()] = (w.bytes(), d.get_comment(), d.last_sha1_ok())
    assert list(got) == [n.split("/")[-1] for n in files]
    for n, data in files.items():
        assert got[n.split("/")[-1]] == (data, "%d bytes" % len(data), 1)
