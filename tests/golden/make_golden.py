#!/usr/bin/env python
"""Regenerates tests/golden/blocks.json from the CPU oracle.

Provenance: the reference is V source and cannot be run in the build image (no V toolchain), and
its own tests hold no golden compressed bytes.  These vectors are therefore ORACLE outputs
(oracle/zpaq_oracle.c, a restatement of the V files checked against every exact KAT the reference
has and against the independent probe vectors of BASELINE.md section 4).  On a machine with V, the
same inputs can be pushed through zpaq.Compressor to confirm them:

    tools/confirm_with_v.sh /path/to/zpaq-v     # see INTEGRATION.md; needs `v` on PATH

Run:  python tests/golden/make_golden.py                       regenerate the json files from the oracle
      python tests/golden/make_golden.py --write-inputs DIR    the nine inputs as files (for tools/ref_golden.v)
      python tests/golden/make_golden.py --compare REF.txt     output of tools/ref_golden.v against the json files
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import datagen  # noqa: E402
import oracle_binding as ob  # noqa: E402


def inputs():
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
        "struct8k": datagen.structured(8192),
    }


JIDAC_DATE = 20260101120000
JIDAC_OPTS = [dict(level=0, fragment=-1, dedup=False, block_bytes=0),
              dict(level=0, fragment=0, dedup=True, block_bytes=0),
              dict(level=1, fragment=2, dedup=True, block_bytes=16384),
              dict(level=2, fragment=6, dedup=True, block_bytes=1 << 20),
              dict(level=1, fragment=23, dedup=False, block_bytes=4096)]


def jidac_tree():
    ins = inputs()
    t = datagen.text(150000, datagen.SEED0 + 77)
    files = dict(ins)
    files.update({"big.txt": t, "big-copy.txt": t, "mix": t[:50000] + ins["rand4k"] + t[50000:90000]})
    return list(files), list(files.values())


def main():
    out = {"provenance": "oracle/zpaq_oracle.c (parity unpinned against a V build; see make_golden.py)",
           "inputs": {}, "blocks": []}
    ins = inputs()
    for name, data in ins.items():
        out["inputs"][name] = {"len": len(data), "sha1": hashlib.sha1(data).hexdigest()}
    for level in range(6):
        for name, data in ins.items():
            arc = ob.compress_block(level, data, name, "%d bytes" % len(data))
            rec = {"level": level, "input": name, "len": len(arc), "sha1": hashlib.sha1(arc).hexdigest()}
            if len(arc) <= 160:
                rec["hex"] = arc.hex()
            out["blocks"].append(rec)
    with open(os.path.join(HERE, "blocks.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["blocks"]), "vectors")
    # journaling archives (oracle/jidac_oracle.c): the reference's create_archive layout with the complete
    # bytes of a tiny tree, and hashes + fragment tables for the fragmentation / dedup option sets
    jd = {"provenance": "oracle/jidac_oracle.c: archive layout pinned by jidac.v:47-296; fragmentation rule "
                        "(fragment >= 0) is upstream zpaq's, parity unpinned",
          "date": JIDAC_DATE, "archives": []}
    names, files = jidac_tree()
    tiny = {"a": b"hello world", "empty": b""}
    jd["tiny_reference_hex"] = ob.jidac_add(list(tiny), list(tiny.values()), JIDAC_DATE).hex()
    for opt in JIDAC_OPTS:
        arc = ob.jidac_add(names, files, JIDAC_DATE, **opt)
        frs, stored = ob.jidac_fragment(files, opt["fragment"], opt["dedup"])
        jd["archives"].append({"opts": opt, "len": len(arc), "sha1": hashlib.sha1(arc).hexdigest(),
                               "n_fragments": len(frs), "n_stored": stored,
                               "cuts_sha1": hashlib.sha1(b"".join(b"%d,%d,%d;" % (f["off"], f["len"], f["id"])
                                                                   for f in frs)).hexdigest()})
    with open(os.path.join(HERE, "jidac.json"), "w") as f:
        json.dump(jd, f, indent=1)
    print("wrote", len(jd["archives"]), "jidac vectors")


def write_inputs(dst):
    os.makedirs(dst, exist_ok=True)
    for name, data in inputs().items():
        with open(os.path.join(dst, name), "wb") as f:
            f.write(data)
    print("wrote", len(inputs()), "inputs to", dst)


def compare(ref_txt):
    """Lines of tools/ref_golden.v (the V reference) against blocks.json / jidac.json (the oracle)."""
    with open(os.path.join(HERE, "blocks.json")) as f:
        want = {(b["level"], b["input"]): (b["len"], b["sha1"]) for b in json.load(f)["blocks"]}
    with open(os.path.join(HERE, "jidac.json")) as f:
        tiny = json.load(f)["tiny_reference_hex"]
    seen, bad = set(), []
    for line in open(ref_txt):
        parts = line.split()
        if not parts:
            continue
        if parts[0] == "block":
            key = (int(parts[1]), parts[2])
            seen.add(key)
            if key not in want:
                bad.append("unexpected %r" % (key,))
            elif want[key] != (int(parts[3]), parts[4]):
                bad.append("level %d input %s: reference %s bytes sha1 %s, oracle %d bytes sha1 %s"
                           % (key[0], key[1], parts[3], parts[4], want[key][0], want[key][1]))
        elif parts[0] == "jidac_tiny":
            seen.add("jidac")
            if parts[1] != tiny:
                bad.append("jidac tiny archive differs: reference %s..., oracle %s..." % (parts[1][:64], tiny[:64]))
    for key in want:
        if key not in seen:
            bad.append("missing from the reference output: %r" % (key,))
    if "jidac" not in seen:
        bad.append("missing from the reference output: jidac_tiny")
    if bad:
        print("\n".join(bad))
        print("PARITY WITH THE V REFERENCE: %d of %d vectors differ" % (len(bad), len(want) + 1))
        sys.exit(1)
    print("PARITY WITH THE V REFERENCE: all %d block vectors and the journaling archive are byte-identical"
          % len(want))


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--write-inputs":
        write_inputs(sys.argv[2])
    elif len(sys.argv) == 3 and sys.argv[1] == "--compare":
        compare(sys.argv[2])
    else:
        main()
