"""GPU jidac front end (zpaqgpu_jidac_fragment / zpaqgpu_jidac_add through the C ABI) against the
CPU oracle: fragment boundaries, SHA-1s, dedup ids and complete journaling archives byte for byte;
extraction restores every file."""
import hashlib

import pytest

import datagen
import oracle_binding as ob

pytestmark = pytest.mark.gpu
DATE = 20260101120000


def tree(n=40, seed=11):
    """A small file tree in the shape of BASELINE.json configs[4]: log-uniform sizes, text and
    binary, 30 % exact duplicates, some empty files."""
    import numpy as np
    rng = np.random.RandomState(seed)
    files = {}
    for i in range(n):
        if i > 3 and rng.rand() < 0.3:
            src = list(files.values())[rng.randint(len(files))]
            files["dup%03d" % i] = src
            continue
        size = int(2 ** rng.uniform(6, 17))
        kind = rng.rand()
        if i % 13 == 5:
            data = b""
        elif kind < 0.6:
            data = datagen.text(size, 100 + i)
        elif kind < 0.8:
            data = datagen.random_bytes(size, 100 + i)
        else:
            data = datagen.structured(size, 100 + i)
        files["file%03d" % i] = data
    return files


@pytest.mark.parametrize("fragment", [-1, 0, 2, 6])
@pytest.mark.parametrize("dedup", [False, True])
def test_fragment_table_matches_oracle(gpu_ctx, fragment, dedup):
    files = list(tree().values()) + [datagen.text(700000, 5), bytes(200000), datagen.random_bytes(300000, 6)]
    got, stored = gpu_ctx.jidac_fragment(files, fragment, dedup)
    want, want_stored = ob.jidac_fragment(files, fragment, dedup)
    assert stored == want_stored
    assert got == want
    data = b"".join(files)
    for f in got[:50]:
        assert f["sha1"] == hashlib.sha1(data[f["off"]:f["off"] + f["len"]]).digest()


def test_reference_create_archive_bytes(gpu_ctx):
    """opts {fragment -1, no dedup, one d block per file, store}: JidacArchive.create_archive."""
    files = tree(25)
    got = gpu_ctx.jidac_add(list(files), list(files.values()), DATE)
    assert got == ob.jidac_add(list(files), list(files.values()), DATE)
    assert gpu_ctx.jidac_add([], [], DATE) == ob.jidac_add([], [], DATE)


@pytest.mark.parametrize("level,fragment,dedup,block_bytes", [
    (0, 2, True, 0), (1, 2, True, 65536), (1, 6, True, 1 << 20), (2, 0, False, 30000), (1, -1, True, 1 << 18),
    (3, 4, True, 1 << 17),
])
def test_add_matches_oracle_and_extracts(gpu_ctx, level, fragment, dedup, block_bytes):
    import zpaq_v_b200 as z
    gpu_ctx.set_workspace_limit(8 << 30)
    files = tree(30, seed=level * 7 + fragment + 3)
    names = list(files)
    got = gpu_ctx.jidac_add(names, [files[k] for k in names], DATE, level=level, fragment=fragment, dedup=dedup,
                            block_bytes=block_bytes)
    want = ob.jidac_add(names, [files[k] for k in names], DATE, level=level, fragment=fragment, dedup=dedup,
                        block_bytes=block_bytes)
    assert got == want
    st = gpu_ctx.jidac_stats()
    assert st["launches"] > 0 and st["n_files"] == len(files) and st["archive_bytes"] == len(got)
    if dedup:
        assert st["stored_bytes"] < st["input_bytes"]
    back = z.jidac.extract(got, gpu_ctx)
    assert back == files
    assert list(back) == names                      # index order
    assert z.jidac.extract_on_host(got, gpu_ctx) == files
    gpu_ctx.set_workspace_limit(0)


def test_extract_details(gpu_ctx):
    """Per-file records of zpaqgpu_jidac_extract; a damaged fragment is reported, a damaged table refused."""
    import zpaq_v_b200 as z
    files = {"a.txt": datagen.text(30000, 3), "empty": b"", "b.bin": datagen.random_bytes(9000, 4),
             "a-again": datagen.text(30000, 3)}
    arc = gpu_ctx.jidac_add(list(files), list(files.values()), DATE, level=0, fragment=0, dedup=True, block_bytes=0)
    recs = gpu_ctx.jidac_extract(arc)
    assert [r["name"] for r in recs] == list(files)
    assert all(r["data"] == files[r["name"]] and r["sha1_ok"] == 1 and r["date"] == DATE for r in recs)
    assert recs[1]["n_fragments"] == 0 and recs[0]["n_fragments"] == recs[3]["n_fragments"] > 1
    assert gpu_ctx.jidac_extract(b"") == []
    assert gpu_ctx.jidac_extract(gpu_ctx.jidac_add([], [], DATE)) == []
    # the reference's own create_archive bytes (from the oracle) extract as well
    ref = ob.jidac_add(list(files), list(files.values()), DATE)
    assert {r["name"]: r["data"] for r in gpu_ctx.jidac_extract(ref)} == files
    # flip one stored byte of the first d block (store mode: plaintext sits in the archive): the block
    # checksum no longer matches and the call refuses the archive
    at = arc.index(files["a.txt"][:64])
    bad = bytearray(arc)
    bad[at + 10] ^= 1
    with pytest.raises(z.ZpaqGpuError):
        gpu_ctx.jidac_extract(bytes(bad))


def test_mirror_class(gpu_ctx):
    import zpaq_v_b200 as z
    files = {"a.txt": b"hello world", "b.bin": datagen.random_bytes(5000, 1), "c": b""}
    w = z.FileWriter()
    a = z.JidacArchive.new(gpu_ctx, date=DATE)
    a.create_archive(files, 3)       # no output set: silently returns (jidac.v:182-184)
    assert w.bytes() == b""
    a.set_output(w)
    a.create_archive(files, 3)       # method is ignored by the reference: always store
    assert w.bytes() == ob.jidac_add(list(files), list(files.values()), DATE)
    w2 = z.FileWriter()
    a.set_output(w2)
    a.add(files, level=1, fragment=0, dedup=True, block_bytes=4096)
    assert z.jidac.extract(w2.bytes(), gpu_ctx) == files


def test_large_single_file_and_many_small(gpu_ctx):
    """One file far above the maximum fragment size, and more files than one CTA handles."""
    big = datagen.text(3 << 20, 9)
    pool = datagen.text(4 << 20, 1000)   # one call: every datagen.text call builds an 8 MiB chunk
    files = [big] + [pool[13001 * i:13001 * i + 100 + 37 * i] for i in range(300)] + [big]
    got, stored = gpu_ctx.jidac_fragment(files, 6, True)
    want, want_stored = ob.jidac_fragment(files, 6, True)
    assert got == want and stored == want_stored
    # the second copy of the big file deduplicates completely
    assert sum(f["len"] for f in got if f["stored"]) == sum(map(len, files)) - len(big)
