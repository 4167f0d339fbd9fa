"""Constant data of the oracle and of libzpaqgpu against the reference's literals, parsed from the
V sources.  /root/reference only exists in the build container: these tests skip elsewhere (the
GPU box never reads it)."""
import os
import re

import pytest

import oracle_binding as ob

REF = "/root/reference/zpaq"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not mounted")


def _ints(text):
    return [int(x) for x in re.findall(r"-?\d+", text)]


def _level_bytes(src, fn):
    body = src[src.index("fn %s()" % fn):]
    body = re.sub(r"//[^\n]*", "", body)
    body = body[body.index("hcomp: ["):]
    body = body[:body.index("]")]
    body = body.replace("u8(", "(")
    return bytes(_ints(body))


def test_state_table_matches_reference():
    src = open(os.path.join(REF, "statetable.v")).read()
    lit = src[src.index("const state_table_data = ["):]
    lit = lit[:lit.index("]!")]
    vals = _ints(lit.replace("u8(1)", "1"))
    assert len(vals) == 1024
    assert bytes(vals) == ob.state_table()


def test_dt_table_matches_reference():
    src = open(os.path.join(REF, "predictor.v")).read()
    lit = src[src.index("const dt_table = ["):]
    lit = lit[:lit.index("]!")]
    vals = _ints(lit.replace("int(87380)", "87380"))
    assert len(vals) == 1024
    assert vals == list(ob.lib().zo_dt_table()[:1024])


def test_level_headers_match_reference():
    src = open(os.path.join(REF, "levels.v")).read()
    names = ["level_0_store", "level_1_fast", "level_2_normal", "level_3_high", "level_4_max", "level_5_max"]
    for level, fn in enumerate(names):
        assert _level_bytes(src, fn) == ob.level_header(level), fn


def test_compsize_and_locator_match_reference():
    src = open(os.path.join(REF, "types.v")).read()
    lit = src[src.index("pub const compsize = ["):]
    lit = re.sub(r"//[^\n]*", "", lit[:lit.index("]")])
    assert _ints(lit) == [ob.lib().zo_compsize(i) for i in range(10)]
    src = open(os.path.join(REF, "compressor.v")).read()
    lit = src[src.index("const zpaq_block_locator = ["):]
    lit = lit[:lit.index("]")]
    loc = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", lit))
    assert ob.compress_block(1, b"", "", "")[:13] == loc


def test_library_constants_match_reference():
    """The product's own copies (zpaq-v_b200/csrc/model.cpp) are built independently of the oracle."""
    import zpaq_v_b200 as z
    src = open(os.path.join(REF, "levels.v")).read()
    names = ["level_0_store", "level_1_fast", "level_2_normal", "level_3_high", "level_4_max", "level_5_max"]
    for level, fn in enumerate(names):
        assert _level_bytes(src, fn) == z.level_header(level), fn
    sq, st, ns = z.tables()
    assert ns == ob.state_table()
