"""CPU-only: the host-side transposition of the packers (hostpack.cpp) - scattered strings vs one blob."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scattered_strings_give_the_same_planes_as_the_blob(tmp_path):
    exe = str(tmp_path / "hostpack_check")
    csrc = os.path.join(ROOT, "epialleler_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + csrc, "-o", exe, os.path.join(ROOT, "tests", "hostpack_check.cpp"),
                    os.path.join(csrc, "hostpack.cpp")], check=True)
    for env in ({}, {"EA_NO_AVX512": "1"}):   # the AVX-512 + GFNI path where the CPU has it, and the AVX2 / scalar one
        out = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, **env))
        assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
