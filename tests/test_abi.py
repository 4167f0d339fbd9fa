"""CPU tests of the C-ABI library: it loads, exports every symbol include/zpaqgpu.h declares, its
host-side constants agree with the oracle, and it fails loudly (no fallback) without a GPU."""
import ctypes as C
import os
import re

import pytest

import oracle_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "zpaqgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zpaqgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from zpaq_v_b200 import binding
    L = binding.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert sorted(binding.EXPORTS) == names


def test_host_constants_match_oracle():
    import zpaq_v_b200 as z
    for level in range(6):
        assert z.level_header(level) == ob.level_header(level)
    sq, st, ns = z.tables()
    assert sq == ob.squash_table() and st == ob.stretch_table() and ns == ob.state_table()


def test_model_geometry():
    import zpaq_v_b200 as z
    want = {1: (10, 11, 25, 26), 2: (13, 14, 28, 29), 3: (19, 20, 40, 41), 4: (28, 29, 55, 56), 5: (34, 35, 67, 68)}
    for level, (cend, hbegin, hend, hsize) in want.items():
        info = z.describe_model(z.level_header(level))
        assert (info["cend"], info["hbegin"], info["hend"], info["hsize"]) == (cend, hbegin, hend, hsize)
        assert info["is_chain"] == 1
    assert [z.describe_model(z.level_header(l))["n_isse"] for l in range(1, 6)] == [1, 2, 4, 5, 7]
    assert [z.describe_model(z.level_header(l))["has_mix2"] for l in range(1, 6)] == [0, 0, 0, 1, 1]
    assert z.describe_model(z.level_header(1))["ctx_mode"] == 1
    assert z.describe_model(z.level_header(3))["ctx_mode"] == 2
    # dense hash-table bytes per block
    mib = 1 << 20
    assert [z.describe_model(z.level_header(l))["hash_table_bytes"] // mib for l in range(1, 6)] == [36, 12, 80, 384, 2048]
    # a header with a CM component is not chain-shaped and goes to the generic kernel
    info = z.describe_model(bytes([2, 2, 0, 0, 1, 2, 16, 4, 0, 96, 4, 28, 59, 112, 56, 0]))
    assert info["is_chain"] == 0 and info["n"] == 1
    # the header written into the block equals what the oracle writes
    blk = ob.compress_block(3, b"", "", "")
    assert blk[18] | (blk[19] << 8) == 41


def test_no_cpu_fallback():
    """Without a CUDA device the library refuses to work instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import zpaq_v_b200 as z
    with pytest.raises(z.ZpaqGpuError) as e:
        z.Context()
    assert e.value.code == z.binding.E_NODEVICE
    # and the mirrored classes cannot be used either
    with pytest.raises(z.ZpaqGpuError):
        z.Compressor()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under zpaq-v_b200/ or include/ may reference it."""
    bad = []
    for base in ("zpaq-v_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh", ".v")):
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    if re.search(r"zpaq_oracle|oracle_binding|oracle/", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof / offsetof of every public struct as gcc sees include/zpaqgpu.h against the ctypes mirrors."""
    import ctypes as C
    import subprocess
    from zpaq_v_b200 import binding as zb
    structs = {"zpaqgpu_model_info": zb.ModelInfo, "zpaqgpu_segment": zb.Segment, "zpaqgpu_stats": zb.Stats,
               "zpaqgpu_jidac_opts": zb.JidacOpts, "zpaqgpu_fragment": zb.Fragment, "zpaqgpu_jidac_file": zb.JidacFile,
               "zpaqgpu_jidac_stats": zb.JidacStats, "zpaqgpu_multi_stats": zb.MultiStats}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "zpaqgpu.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for field, _ in cls._fields_:
            lines.append('printf(" %%zu", offsetof(%s, %s));' % (cname, field))
        lines.append('printf("\\n");')
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe])
    out = subprocess.check_output([exe]).decode().strip().splitlines()
    assert len(out) == len(structs)
    for line, (cname, cls) in zip(out, structs.items()):
        parts = line.split()
        assert parts[0] == cname
        nums = [int(x) for x in parts[1:]]
        assert nums[0] == C.sizeof(cls), cname
        assert nums[1:] == [getattr(cls, f).offset for f, _ in cls._fields_], cname
    # the oracle's fragment record is copied into the same layout by the tests
    import oracle_binding as ob
    assert C.sizeof(ob._Frag) == C.sizeof(zb.Fragment)
