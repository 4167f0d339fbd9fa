"""The C++ host mirror (zpaq-v_b200/host/zpaq_gpu.hpp) compiled against libzpaqgpu and run on the GPU:
Compressor + Decompresser round trip, bytes equal to the oracle."""
import os
import subprocess

import pytest

import datagen
import oracle_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_cpp_mirror_roundtrip(tmp_path):
    exe = str(tmp_path / "roundtrip")
    pkg = os.path.join(ROOT, "zpaq-v_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(pkg, "host"),
                           os.path.join(pkg, "host", "roundtrip_main.cpp"), "-o", exe, "-L" + pkg, "-lzpaqgpu",
                           "-Wl,-rpath," + pkg])
    for level, data in ((1, b"AAAABBBB"), (2, datagen.text(30000)), (0, datagen.random_bytes(70000))):
        r = subprocess.run([exe, str(level)], input=data, capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr.decode()
        hexarc, tail, hexjidac, batch = r.stdout.decode().strip().split("\n")
        assert tail == "1 1 1" and batch == "batch 1"   # batch: two queued blocks delivered by flush() in order
        assert bytes.fromhex(hexarc) == ob.compress_block(level, data, "test", "")
        assert bytes.fromhex(hexjidac) == ob.jidac_add(["a", "b"], [data, data[::-1]], 20260101120000)
