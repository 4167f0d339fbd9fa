"""Untrusted header bytes into the host-side parser and layout builder (zpaq-v_b200/csrc/model.cpp), built
with gcc's UB sanitizer in trap mode and libstdc++ assertions: the headers of all levels, every truncation, byte
mutations at every position and 200 000 random component tables (tests/c/model_fuzz.cpp).  Accepting or
rejecting is both fine; crashing, reading out of range or an inconsistent layout is not."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_model_parser_survives_untrusted_headers(tmp_path):
    exe = str(tmp_path / "model_fuzz")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-D_GLIBCXX_ASSERTIONS", "-fsanitize=undefined",
                           "-fsanitize-undefined-trap-on-error", "-ffp-contract=off",
                           "-I" + os.path.join(ROOT, "zpaq-v_b200", "csrc"),
                           os.path.join(ROOT, "tests", "c", "model_fuzz.cpp"),
                           os.path.join(ROOT, "zpaq-v_b200", "csrc", "model.cpp"), "-o", exe, "-lpthread"])
    out = subprocess.run([exe], capture_output=True, timeout=300)
    assert out.returncode == 0, (out.stdout + out.stderr).decode()[-2000:]
    assert out.stdout.decode().strip().endswith("ok")
