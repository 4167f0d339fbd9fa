"""The lane logic of the SHA-1 kernel (zpaq-v_b200/csrc/sha1_lane.h: staging plan at any byte alignment, message
words out of the staged window, padding blocks) under a host emulation of the warp
(tests/c/sha1_lane_test.cpp, g++ only), against hashlib.  The emulation asserts that no byte outside a job's
range is ever read."""
import hashlib
import os
import random
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("sha1lane") / "sha1_lane_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "zpaq-v_b200", "csrc"),
                           os.path.join(ROOT, "tests", "c", "sha1_lane_test.cpp"), "-o", path])
    return path


def run(exe, data, jobs, base_shift):
    blob = struct.pack("<QQQ", len(jobs), len(data), base_shift)
    blob += b"".join(struct.pack("<QQ", o, n) for o, n in jobs) + data
    out = subprocess.run([exe], input=blob, capture_output=True, timeout=300)
    assert out.returncode == 0, out.stderr.decode()
    got = out.stdout.decode().split()
    assert len(got) == len(jobs)
    for k, (o, n) in enumerate(jobs):
        assert got[k] == hashlib.sha1(data[o:o + n]).hexdigest(), "job %d (off %d, len %d, shift %d)" % (k, o, n, base_shift)


@pytest.mark.parametrize("base_shift", [0, 1, 7, 15, 16, 33])
def test_every_length_and_alignment_around_the_padding_edges(exe, base_shift):
    rnd = random.Random(1234 + base_shift)
    data = bytes(rnd.getrandbits(8) for _ in range(6000))
    lens = list(range(0, 200)) + [255, 256, 257, 311, 312, 319, 320, 321, 511, 512, 513, 1000, 1023, 1024, 1025]
    jobs = [(off, n) for n in lens for off in (0, 1, 3, 4, 13, 16, 17)]
    jobs += [(len(data) - n, n) for n in (0, 1, 15, 16, 17, 55, 56, 63, 64, 65, 300)]   # up against the end of the data
    run(exe, data, jobs, base_shift)


def test_ragged_warps_and_long_ranges(exe):
    rnd = random.Random(99)
    data = bytes(rnd.getrandbits(8) for _ in range(300000))
    jobs = []
    for _ in range(70):   # three warps, the last one ragged; lengths from empty to a thousand rounds
        n = rnd.choice([0, 1, 64, rnd.randrange(0, 300), rnd.randrange(0, 5000), rnd.randrange(0, 260000)])
        jobs.append((rnd.randrange(0, len(data) - n + 1), n))
    run(exe, data, jobs, 5)
    run(exe, data, [(0, len(data))], 0)
    run(exe, b"", [(0, 0)], 0)
    run(exe, b"", [], 0)
