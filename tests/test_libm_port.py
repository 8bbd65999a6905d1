"""csrc/nm_libm.h (the host C library's sin / cos / atan2 restated with their fused operations, for the simulation
kernel's pose chain) against the C library itself, bit for bit, on the host: tests/libm_port_check.c compiled with
-ffp-contract=off and run on 2e6 arguments per regime (headings, small arguments, large arguments up to 1e8, points next
to multiples of pi/2; displacement vectors on a disc, 52 binades either way, next to the axes, on the diagonals, signed
zeros / subnormals / huge values).  Every operation in the header is an IEEE basic operation or an FMA, which round the
same on the device, so this is also the device's parity proof; tests/test_sim_scale_gpu.py confirms it on 4.1e6
agent-steps of poses.  (A 2e8-per-regime run -- 2e9 evaluations -- is clean too: profiles/r02_libm_port_check.txt.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_port_equals_host_libm(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "libm_port_check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", os.path.join(ROOT, "tests", "libm_port_check.c"),
                    "-lm", "-o", exe], check=True)
    res = subprocess.run([exe, "2000000"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    lines = [ln.split() for ln in res.stdout.strip().splitlines()]
    assert len(lines) == 10 and all(ln[-1] == "0" for ln in lines), res.stdout


def test_tables_header_is_what_the_extractor_writes():
    """nm_libm_tables.h is generated from the image's libm.so.6 (tools/libm_tables.py re-checks the sin / cos table against
    exact arithmetic); when that library is present and is the one that was used, regenerating must give the same file."""
    import hashlib
    lib = "/usr/lib/x86_64-linux-gnu/libm.so.6"
    hdr = os.path.join(ROOT, "navigation_model_b200", "csrc", "nm_libm_tables.h")
    text = open(hdr).read()
    if not os.path.exists(lib) or hashlib.sha256(open(lib, "rb").read()).hexdigest() not in text:
        pytest.skip("another C library build on this host")
    before = text
    subprocess.run(["python", os.path.join(ROOT, "tools", "libm_tables.py")], check=True, capture_output=True)
    assert open(hdr).read() == before
