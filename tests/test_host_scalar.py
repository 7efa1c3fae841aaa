"""Host-side scalar arithmetic of the protocol mirror (csrc/host/sc_host.hpp), checked without a GPU: the binary
inversion used for the inner-product challenges (reference src/inner_product_proof.rs:122-123, `u.invert()`)
against the exponentiation it replaced."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scalar_invert_against_exponentiation():
    src = os.path.join(ROOT, "tests", "hostcheck", "sc_invert_check.cpp")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "sc_invert_check")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, src])
        out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "mismatches 0", out.stdout + out.stderr
