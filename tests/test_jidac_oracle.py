"""The jidac oracle (oracle/jidac_oracle.c) against what jidac.v pins: block names, comments,
store-mode framing, c/d/h/i contents and order (jidac.v:47-49, :67-118, :181-296), and the
size-independent properties of the fragmentation rule (parity unpinned, see the oracle header)."""
import hashlib
import struct

import datagen
import oracle_binding as ob

LOCATOR = bytes([0x37, 0x6b, 0x53, 0x74, 0xa0, 0x31, 0x83, 0xd3, 0x8c, 0xb2, 0x28, 0xb0, 0xd3])
DATE = 20260101120000


def store_block(name, data):
    """create_jidac_block (jidac.v:67-91) written out by hand from the container grammar
    (SURVEY 8-F): store mode, one segment, comment "<usize> jDC\\x01"."""
    comment = b"%d jDC\x01" % len(data)
    body = b"\x00" + data  # PP byte then data, chunks close at 65536 bytes (compressor.v:297-354)
    chunks = b"".join(struct.pack(">I", len(body[i:i + 65536])) + body[i:i + 65536]
                      for i in range(0, len(body), 65536))
    return (LOCATOR + b"zPQ\x02\x01\x07\x00" + bytes(7) + b"\x01" + name + b"\x00" + comment + b"\x00\x00" +
            chunks + bytes(4) + b"\xfd" + hashlib.sha1(data).digest() + b"\xff")


def name(kind, num):
    return b"jDC%014d%s%010d" % (DATE, kind, num)


def test_reference_layout_by_hand():
    files = {"a.txt": b"hello world", "empty": b"", "big.bin": datagen.random_bytes(70000, 3)}
    got = ob.jidac_add(list(files), list(files.values()), DATE)
    d, h, ic = [], [], b""
    for k, (nm, data) in enumerate(files.items(), 1):
        blk = store_block(name(b"d", k), data)
        d.append(blk)
        h.append(store_block(name(b"h", k), struct.pack("<I", len(blk)) + hashlib.sha1(data).digest() +
                             struct.pack("<I", len(data))))
        ic += struct.pack("<q", DATE) + nm.encode() + b"\x00" + struct.pack("<III", 0, 1, k)
    want = (store_block(name(b"c", len(files) + 1), struct.pack("<q", sum(map(len, d)))) + b"".join(d) +
            b"".join(h) + store_block(name(b"i", 1), ic))
    assert got == want


def test_no_files_is_a_lone_c_block():
    assert ob.jidac_add([], [], DATE) == store_block(name(b"c", 1), bytes(8))


def test_archive_decodes_with_the_oracle_decompresser():
    files = {"f%d" % i: datagen.text(3000 + 977 * i, i) for i in range(5)}
    arc = ob.jidac_add(list(files), list(files.values()), DATE, level=1, fragment=2, dedup=True, block_bytes=4096)
    d = ob.Decompresser()
    d.set_input(arc)
    names = []
    while d.find_block():
        while d.find_filename():
            names.append(d.get_filename())
            d.decompress(-1)
            d.read_segment_end()
            assert d.last_sha1_ok() == 1
    kinds = "".join(n[17] for n in names)
    assert kinds[0] == "c" and kinds[-1] == "i" and kinds.count("d") == kinds.count("h") >= 2
    assert kinds == "c" + "d" * kinds.count("d") + "h" * kinds.count("h") + "i"


def test_fragment_rule_properties():
    data = datagen.text(600000, 7) + datagen.random_bytes(300000, 8) + bytes(100000)
    for fragment in (0, 2, 4, 6):
        ends = ob.fragment_ends(data, fragment)
        assert ends[-1] == len(data) and ends == sorted(set(ends))
        sizes = [b - a for a, b in zip([0] + ends, ends)]
        assert all(64 << fragment <= s <= 8128 << fragment for s in sizes[:-1])
        assert sizes[-1] <= 8128 << fragment
        # content-defined: the cut points after an insertion at the front realign
        shifted = ob.fragment_ends(b"XYZ" + data, fragment)
        common = {e + 3 for e in ends} & set(shifted)
        assert len(common) >= len(ends) // 2
    # every fragment starts from the same state, so constant input is cut with a constant period
    z = ob.fragment_ends(bytes(3 << 20), 6)
    assert len({b - a for a, b in zip([0] + z[:-1], z[:-1])}) == 1
    # fragment > 22: no hash cuts, only the maximum size
    assert ob.fragment_ends(bytes(100), 23) == [100]
    assert ob.fragment_ends(b"", 6) == [] and ob.fragment_ends(b"", -1) == [0]
    assert ob.fragment_ends(data, 23) == [len(data)]
    # incompressible bytes are never predicted: cuts only where the hash falls below the threshold
    r = ob.fragment_ends(datagen.random_bytes(1 << 20, 5), 0)
    assert 300 < len(r) < 3000      # average fragment 1 KiB + the 64-byte minimum


def test_dedup_ids_follow_first_occurrence():
    a, b = datagen.text(50000, 1), datagen.random_bytes(20000, 2)
    files = [a, b, a, b"", a + b, b""]
    frs, stored = ob.jidac_fragment(files, -1, True)
    assert [f["id"] for f in frs] == [1, 2, 1, 3, 4, 3] and stored == 4
    assert [f["stored"] for f in frs] == [1, 1, 0, 1, 1, 0]
    frs, stored = ob.jidac_fragment(files, 2, True)
    assert stored == len({(f["sha1"], f["len"]) for f in frs})
    seen = {}
    for f in frs:
        key = (f["sha1"], f["len"])
        assert f["stored"] == (key not in seen)
        seen.setdefault(key, f["id"])
        assert f["id"] == seen[key]
        assert f["sha1"] == hashlib.sha1(b"".join(files)[f["off"]:f["off"] + f["len"]]).digest()
    assert sorted(set(seen.values())) == list(range(1, stored + 1))
