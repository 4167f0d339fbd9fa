"""CPU tests: the oracle against every known-answer value the reference's own tests hold for this
path and against the independent survey probe vectors in BASELINE.md
section 4.  No GPU needed."""
import hashlib
import struct

import pytest

import datagen
import oracle_binding as ob


# ---- zpaq_test.v:5-27 ----
def test_sha1_kats():
    assert ob.sha1(b"").hex() == "da39a3ee5e6b4b0d3255bfef95601890afd80709"
    assert ob.sha1(b"abc").hex() == "a9993e364706816aba3e25717850c26c9cd0d89d"
    for n in (1, 55, 56, 57, 63, 64, 65, 119, 120, 1000):
        d = datagen.random_bytes(n)
        assert ob.sha1(d) == hashlib.sha1(d).digest()


# ---- zpaq_test.v:55-107 ----
def test_statetable_kats():
    ns = ob.state_table()
    n0n1 = lambda s: (ns[s * 4 + 2], ns[s * 4 + 3])
    assert n0n1(0) == (0, 0) and n0n1(1) == (1, 0) and n0n1(2) == (0, 1)
    L = ob.lib()
    assert L.zo_st_next(0, 0) == 1 and L.zo_st_next(0, 1) == 2
    assert L.zo_st_cminit(0) == 1 << 22
    assert L.zo_st_cminit(1) == (1 << 22) // 2
    assert L.zo_st_cminit(2) == (3 << 22) // 2
    assert L.zo_st_next(256, 0) == 0 and L.zo_st_next(-1, 1) == 0   # statetable.v:76-78
    assert len(ns) == 1024 and ns[1020:] == bytes(4) and max(ns[0::4]) == 253


# ---- zpaq_test.v:266-271 ----
def test_oplen_kats():
    L = ob.lib()
    assert [L.zo_oplen(o) for o in (0, 7, 56, 255)] == [1, 2, 1, 3]
    assert L.zo_oplen(63) == 2 and L.zo_oplen(39) == 2 and L.zo_oplen(8) == 1


# ---- zpaq_test.v:281-292 (loose) + BASELINE.md section 4 (exact probe values) ----
def test_squash_stretch():
    L = ob.lib()
    assert 15000 <= L.zo_squash(0) <= 18000
    assert 50 <= L.zo_stretch(L.zo_squash(100)) <= 150
    sq, st = ob.squash_table(), ob.stretch_table()
    assert hashlib.sha1(struct.pack("<4096i", *sq)).hexdigest() == "018a9d4ea034deda595b4daed1a7ec11abd01c5f"
    assert hashlib.sha1(struct.pack("<32768i", *st)).hexdigest() == "fdbbf249ad2c2651ebccb46005f9240afd788897"
    want = {0: 16384, 1: 16511, -1: 16256, 64: 23955, -64: 8812, 256: 32178, -256: 589, 512: 32756, -512: 11,
            700: 32766, 1017: 32767, 1018: 1, 2047: 1, -1017: 1, -1018: 32767, -2048: 32767}
    for d, v in want.items():
        assert L.zo_squash(d) == v, d
    want = {1: -375, 100: -342, 1000: -221, 4096: -124, 8192: -70, 16384: 0, 24576: 70, 32000: 238, 32766: 375,
            32767: 2047}
    for p, v in want.items():
        assert L.zo_stretch(p) == v, p
    assert len(set(st)) == 753
    assert L.zo_stretch(L.zo_squash(100)) == 100


# ---- zpaq_test.v:387-402, levels.v ----
def test_level_headers():
    assert ob.level_header(0) == bytes(7)
    assert [len(ob.level_header(l)) for l in range(6)] == [7, 26, 30, 42, 57, 69]
    assert ob.level_header(9) == ob.level_header(1)          # levels.v:34 fallback
    for l in range(1, 6):
        assert ob.level_header(l)[4] > 0


# ---- zpaq_test.v:302-314: coder initial state => empty input codes to EOF + flush only ----
def test_empty_stream_is_eof_and_flush():
    out = ob.raw_encode(ob.level_header(1), b"", with_pp=False)
    # encode(1, p=0) from low=1/high=0xFFFFFFFF emits the four bytes of low, flush emits high
    assert out == bytes([0, 0, 0, 1]) + b"\xff\xff\xff\xff"


# ---- zpaq_test.v:430-527 ----
def test_codec_roundtrip_level1_hello():
    hdr = ob.level_header(1)
    code = ob.raw_encode(hdr, b"Hello World!")
    assert code.hex() == "b706907e03c4d499e77943b7ce37a8ffffffff"
    assert ob.raw_decode(hdr, code) == b"Hello World!"


def test_probe_payload_vectors():
    exp = {1: "ffa7cac11976109b43f54b0d0d6e425cffffffff", 2: "ff9f36bf077ab2f2e553cee6d805a461ffffffff",
           3: "ff99184e5b809c9928e2a4055b5d652fffffffff"}
    exp[4] = exp[5] = exp[3]
    for l, h in exp.items():
        assert ob.raw_encode(ob.level_header(l), b"Hello World!", with_pp=True).hex() == h
    assert ob.raw_encode(ob.level_header(1), b"", with_pp=True).hex() == "fede5de4ffffffff"


# ---- zpaq_test.v:364-384 shape, exact bytes from BASELINE.md section 4 ----
def test_basic_compression_block_bytes():
    want = ("376b5374a03183d38cb228b0d37a5051" "01" "01" "1a00" "0102000002031008130000"
            "60041c3b0a3b70190a3b0a3b703800" "01" "7465737400" "00" "00" "ffb093fea1530cc92f3a28ffffffff"
            "00000000" "fd" "7cd188ef3a9ea7fa0ee9c62c168709695460f5c0" "ff")
    c = ob.Compressor()
    c.set_input(b"AAAABBBB")
    c.start_block(1)
    c.start_segment("test", "")
    while c.compress(8):
        pass
    c.end_segment()
    c.end_block()
    assert c.output().hex() == want
    assert ob.compress_block(1, b"AAAABBBB", "test", "").hex() == want


def ci_files():
    """The five file shapes of .github/workflows/compress-decompress.yml:41-67."""
    return {
        "test1.txt": b"Hello, this is a test file for ZPAQ compression.\n",
        "test2.txt": b"".join(b"This is line %d of repetitive content for compression testing.\n" % i for i in range(1, 101)),
        "random.bin": datagen.random_bytes(5120),
        "empty.txt": b"",
        "subdir/nested.txt": b"Nested file content\n",
    }


@pytest.mark.parametrize("level", range(6))
def test_roundtrip_ci_shapes(level):
    files = ci_files()
    arc = b"".join(ob.compress_block(level, d, n.split("/")[-1], "%d bytes" % len(d)) for n, d in files.items())
    plain, segs, bad = ob.decompress_archive(arc)
    assert plain == b"".join(files.values()) and segs == len(files) and bad == 0


def test_object_api_walk_and_multisegment():
    """Q16/Q17: several segments in one block share the model tables; the PP byte is coded on the
    first compress() call of each segment."""
    c = ob.Compressor()
    c.start_block(2)
    parts = [b"first segment " * 40, b"", b"second segment " * 40]
    for i, d in enumerate(parts):
        c.set_input(d)
        c.start_segment("s%d" % i, "c%d" % i)
        while c.compress(100):
            pass
        c.end_segment()
    c.end_block()
    arc = c.output()
    d = ob.Decompresser()
    d.set_input(arc)
    assert d.find_block()
    got = []
    while d.find_filename():
        got.append((d.get_filename(), d.get_comment()))
        while d.decompress(64):
            pass
        d.read_segment_end()
        assert d.last_sha1_ok() == 1
    assert got == [("s0", "c0"), ("s1", "c1"), ("s2", "c2")]
    assert d.output() == b"".join(parts)
    assert not d.find_block()
    # a block whose segment never saw compress() has no PP byte and still parses (Q16)
    c = ob.Compressor()
    c.start_block(1)
    c.start_segment("x", "")
    c.end_segment()
    c.end_block()
    plain, segs, bad = ob.decompress_archive(c.output())
    assert plain == b"" and segs == 1 and bad == 0


def test_store_mode_chunking():
    """compressor.v:297-354: PP byte + data in chunks closed at 65536 bytes."""
    for n in (0, 1, 65534, 65535, 65536, 65537, 200000):
        data = datagen.random_bytes(n)
        arc = ob.compress_block(0, data, "f", "")
        body = arc[arc.index(b"\x01f\x00\x00\x00") + 5:]
        first = int.from_bytes(body[:4], "big")
        assert first == min(n + 1, 65536)
        assert ob.decompress_archive(arc)[0] == data


def test_silent_state_errors():
    """compressor.v:80-82, :213-215, :260-262: calls in the wrong state are ignored."""
    c = ob.Compressor()
    c.start_segment("a", "b")      # no block yet
    assert c.output() == b""
    c.set_input(b"abc")
    assert c.compress(3) is False
    c.start_block(1)
    n = len(c.output())
    c.start_block(2)               # ignored: already in a block
    assert len(c.output()) == n


CUSTOM_HEADERS = {
    # hh hm ph pm n comps.. 0 hcomp.. 0
    "cm": [2, 2, 0, 0, 1, 2, 16, 4, 0, 96, 4, 28, 59, 112, 56, 0],
    "cons_cm_avg": [2, 2, 0, 0, 3, 1, 160, 2, 12, 8, 5, 0, 1, 100, 0, 96, 28, 59, 25, 112, 56, 0],
    "icm_cm_mix": [3, 8, 0, 0, 3, 3, 12, 2, 14, 16, 7, 10, 0, 2, 24, 255, 0,
                   104, 17, 28, 65, 112, 25, 59, 112, 25, 68, 112, 56, 0],
    "icm_match_mix2_sse": [3, 10, 0, 0, 5, 3, 14, 4, 12, 14, 6, 10, 0, 1, 20, 255, 8, 12, 2, 9, 10, 3, 32, 100, 0,
                           104, 17, 28, 59, 112, 25, 60, 25, 59, 112, 25, 65, 112, 25, 112, 56, 0],
    "vm_branches": [2, 6, 0, 0, 2, 2, 14, 20, 8, 12, 0, 0,
                    # a>N jt: exercise compare, conditional jump (Q5 offsets), arithmetic, r[] file
                    104, 17, 239, 96, 39, 2, 135, 7, 55, 3, 7, 3, 151, 5, 28, 112, 25, 60, 56, 0],
    # wiring the reference accepts because it only checks j < n (predictor.v:575-631): inputs that come LATER in
    # the header (their prediction is the previous bit's), a MIX that contains itself, a MIX2 fed by itself, two
    # MIX components.  0 ICM, 1 AVG(2,0), 2 CM, 3 ISSE(j=5), 4 MIX(0..5), 5 ICM, 6 MIX2(4,6), 7 MIX(3..6)
    "forward_refs": [3, 8, 0, 0, 8, 3, 12, 5, 2, 0, 100, 2, 14, 8, 8, 12, 5, 7, 8, 0, 6, 24, 255, 3, 10,
                     6, 8, 4, 6, 16, 0, 7, 0, 3, 4, 12, 0, 0,
                     96, 4, 28, 59, 112, 25, 10, 59, 112, 25, 10, 59, 112, 25, 59, 112, 25, 59, 112, 25, 60, 25, 112, 25,
                     112, 56, 0],
    # twenty components of every type, one per lane of the warp kernel: 8 ICM, an ISSE chain and a side ISSE,
    # MATCH, 2 CM, AVG, SSE, a 17-input MIX, MIX2, final SSE
    "twenty": ([5, 10, 0, 0, 20] + [3, 10, 3, 11, 3, 12, 3, 10, 3, 11, 3, 12, 3, 10, 3, 11]
               + [8, 11, 0, 8, 11, 8, 8, 10, 9, 8, 10, 3] + [4, 12, 14] + [2, 12, 20, 2, 14, 255] + [5, 10, 11, 128]
               + [9, 8, 15, 16, 64] + [7, 8, 0, 17, 20, 255] + [6, 10, 16, 17, 24, 255] + [9, 10, 18, 32, 255] + [0]
               + [96, 4, 28] + [59, 112, 25, 10] * 6 + [59, 112, 25] * 8 + [60, 25] * 3 + [59, 112, 25] * 3 + [56, 0]),
}


@pytest.mark.parametrize("name", sorted(CUSTOM_HEADERS))
def test_custom_header_roundtrip(name):
    """All nine component types and a branching HCOMP program through the oracle's own decoder
    (headers the reference decoder accepts, decompressor.v:278-342)."""
    hdr = bytes(CUSTOM_HEADERS[name])
    for data in (b"", b"abracadabra " * 50, datagen.text(6000), datagen.random_bytes(1500), bytes(3000)):
        arc = ob.compress_block(0, data, "f", "c", header=hdr)
        plain, segs, bad = ob.decompress_archive(arc)
        assert plain == data and segs == 1 and bad == 0, name
