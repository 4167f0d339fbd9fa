"""The pure host-side decisions of libzpaqgpu (zpaq-v_b200/csrc/hostlogic.h: the split of a batch over devices,
where an archive may be cut, wave and CTA sizes) compiled with g++ and checked on the CPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hostlogic(tmp_path):
    exe = str(tmp_path / "hostlogic_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "zpaq-v_b200", "csrc"),
                           os.path.join(ROOT, "tests", "c", "hostlogic_test.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, timeout=120)
    assert out.returncode == 0, out.stderr.decode()
    assert out.stdout.decode().strip().endswith("ok")
