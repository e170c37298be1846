"""CPU tests of the CUDA kernels themselves: the kernels of the bucketed hot path (pycuda-euler_b200/csrc/bucket_part.cu,
bucket_build.cu) are compiled with -DEULER_SIMT_EMU against tests/host/simt_emu.h (every CUDA thread an OS thread, warp
collectives as rendezvous that abort when the lanes of a warp disagree about which collective they are at) and run on
small inputs against brute force:
* tests/host/simt_build_check.cpp -- the per-bucket build: first and second pass, the look-back, the cross-bucket
  post-pass, the TABLE / OUTPUT overflow protocol;
* tests/host/simt_part_check.cpp -- the partition pass in its direct and its multi-GPU stream form (1, 2, 3, 8 and 16
  emulated ranks scattering into each other's stream areas), count push, owner-side regroup, then the build per owner:
  the union of the owners' graphs must be the graph of all reads.
A deadlock shows as the timeout below.

The emulator exists because of a bug it reproduces in seconds: a version of the first pass read a block-shared flag
inside its warp-uniform record loop, one rank of an 8-GPU run stalled for good, and under the emulator the same source
stops with "lanes of a warp met at different collectives"."""
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pycuda-euler_b200", "csrc")
HOST = os.path.join(ROOT, "tests", "host")


def _cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def test_emulator_constants_match_the_library():
    """simt_emu.h restates a few constants of common.cuh / kernels.h: they must not drift."""
    emu = open(os.path.join(HOST, "simt_emu.h")).read()
    lib = open(os.path.join(CSRC, "kernels.h")).read() + open(os.path.join(CSRC, "common.cuh")).read()
    names = re.findall(r"#define (BKT_[A-Z_]+|EULER_EMPTY_KEY|EULER_NO_ID) ", emu)
    assert len(names) >= 9
    for name in names:
        a = re.search(r"#define %s\s+(\S+)" % name, emu).group(1)
        b = re.search(r"#define %s\s+(\S+)" % name, lib).group(1)
        assert int(a.rstrip("ulUL"), 0) == int(b.rstrip("ulUL"), 0), name


def test_build_kernels_under_the_simt_emulator():
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "simt_build_check")
        subprocess.check_call([_cxx(), "-O1", "-std=c++17", "-pthread", "-I", CSRC, "-I", HOST,
                               os.path.join(HOST, "simt_build_check.cpp"), "-o", exe])
        out = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "13 cases, 0 failed" in out.stdout
    assert len(re.findall(r"redo=[1-9]", out.stdout)) >= 3      # the second pass really ran


def test_partition_kernels_under_the_simt_emulator():
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "simt_part_check")
        subprocess.check_call([_cxx(), "-O1", "-std=c++17", "-pthread", "-I", CSRC, "-I", HOST,
                               os.path.join(HOST, "simt_part_check.cpp"), "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "9 cases, 0 failed" in out.stdout
