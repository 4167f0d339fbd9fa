"""The checker checked: the oracle's C sources built with gcc's undefined-behaviour sanitizer in trap mode
(-fsanitize=undefined,bounds-strict -fsanitize-undefined-trap-on-error: shifts, fixed-array bounds, alignment,
... end the process; no runtime library needed) must pass the reference KATs, the golden vectors and the jidac
layout tests exactly like the normal build.  -fwrapv stays: V ints wrap, so signed overflow is
defined behaviour here, not a finding."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_suite_under_ubsan_trap_mode(tmp_path):
    so = str(tmp_path / "libzpaq_oracle_ubsan.so")
    src = [os.path.join(ROOT, "oracle", f) for f in ("zpaq_oracle.c", "jidac_oracle.c")]
    subprocess.check_call(["gcc", "-O1", "-g", "-fwrapv", "-ffp-contract=off", "-fPIC", "-std=c99",
                           "-D_POSIX_C_SOURCE=200809L", "-fsanitize=undefined,bounds-strict",
                           "-fsanitize-undefined-trap-on-error", "-fstack-protector-all"] + src +
                          ["-o", so, "-shared", "-lpthread"])
    env = dict(os.environ, ZPAQ_ORACLE_SO=so)
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "not gpu", "-p", "no:cacheprovider",
                          os.path.join(ROOT, "tests", "test_oracle_kats.py"), os.path.join(ROOT, "tests", "test_golden.py"),
                          os.path.join(ROOT, "tests", "test_jidac_oracle.py")], env=env, capture_output=True,
                         timeout=900, cwd=ROOT)
    tail = out.stdout.decode()[-1500:]
    assert out.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail
