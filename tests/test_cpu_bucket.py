"""CPU test of the bucketed path's partition logic: tests/host/bucket_lane_check.cpp drives the same
__host__ __device__ functions the CUDA kernel uses (pycuda-euler_b200/csrc/bucket.cuh: m-mer scores, sliding
minimum, window validity, piece cutting, 16-byte record packing) lane by lane over warp tiles and compares the
decoded records with a brute-force statement of the delivery rule (every valid l-mer window reaches the bucket
of its prefix vertex and of its suffix vertex exactly once, with the right ownership bits), for l from 2 to 32,
1 / 3 / 8 ranks, reads with N's, lowercase, ragged lengths and repeats."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_lane_logic_against_brute_force():
    src = os.path.join(ROOT, "tests", "host", "bucket_lane_check.cpp")
    inc = os.path.join(ROOT, "pycuda-euler_b200", "csrc")
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "bucket_lane_check")
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-I", inc, src, "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "0 failed" in out.stdout
