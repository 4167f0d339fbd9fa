#!/usr/bin/env python
"""Regenerates tests/golden/blocks.json from the CPU oracle.

Provenance: the reference is V source and cannot be run in the build image (no V toolchain), and
its own tests hold no golden compressed bytes.  These vectors are therefore ORACLE outputs
(oracle/zpaq_oracle.c, a restatement of the V files checked against every exact KAT the reference
has and against the independent probe vectors of BASELINE.md section 4).  On a machine with V, the
same inputs can be pushed through zpaq.Compressor to confirm them:

    v run tests/golden/confirm_with_v.v

Run:  python tests/golden/make_golden.py
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


if __name__ == "__main__":
    main()
