"""The kernels' exponentials.  The simulation kernel's (ExpSim::exp_nonpos, csrc/nm_exp.cuh: 64-entry table, two-step Cody-Waite reduction,
q = r + r^2 P3(r), fma(s, q, s)) restated on the host with the header's own table and constants (tests/expsim_check.c) and
measured against the long-double exponential: every operation is an IEEE basic operation or an FMA, which round identically
on the device, so the host figures are the device's.  Bar: within 1.5 ulp everywhere, within 1 ulp on 99.5 % of the
arguments, exp(0) == 1 exactly (DESIGN.md section 6.1 quotes 1.27 ulp / 99.86 % on 2e7 arguments)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _initialiser(text, name):
    m = re.search(r"__constant__ double " + name + r"\[\d+\] = \{(.*?)\};", text, re.S)
    assert m, name
    return re.sub(r"//[^\n]*", "", m.group(1))


def _build_check(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    text = open(os.path.join(ROOT, "navigation_model_b200", "csrc", "nm_exp.cuh")).read()
    with open(tmp_path / "expsim_tables.h", "w") as fh:
        fh.write("static const double tab64[64] = {" + _initialiser(text, "c_exp2tab64") + "};\n")
        fh.write("static const double k[7] = {" + _initialiser(text, "c_expsim_k") + "};\n")
        fh.write("static const double kll[4] = {" + _initialiser(text, "c_expll_k") + "};\n")
    exe = str(tmp_path / "expsim_check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-I", str(tmp_path),
                    os.path.join(ROOT, "tests", "expsim_check.c"), "-lm", "-o", exe], check=True)
    return exe


def test_expll_truncation_bound(tmp_path):
    """ExpLL (likelihood kernels: fit, Q-table, model-based): one fused reduction step, 64-entry table, degree-3
    polynomial -- relative error below 4e-11 per exponential on [-40, 0] (the bar on a session's log-likelihood is 1e-9),
    exp(0) exact, hugely negative arguments harmless."""
    exe = _build_check(tmp_path)
    res = subprocess.run([exe, "2000000", "ll"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    worst, ok = res.stdout.split()
    assert float(worst) < 4e-11, res.stdout
    assert ok == "1", res.stdout


def test_expsim_accuracy(tmp_path):
    exe = _build_check(tmp_path)
    res = subprocess.run([exe, "4000000"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    worst, within, exact, one = res.stdout.split()
    assert float(worst) < 1.5, res.stdout
    assert float(within) > 0.995, res.stdout
    assert float(exact) > 0.75, res.stdout
    assert one == "1", res.stdout


def test_table_is_two_to_the_j_over_64():
    """c_exp2tab64[j] is the double nearest to 2^(j/64) (checked in exact rational arithmetic)."""
    from fractions import Fraction
    text = open(os.path.join(ROOT, "navigation_model_b200", "csrc", "nm_exp.cuh")).read()
    vals = [float(v) for v in _initialiser(text, "c_exp2tab64").replace("\n", " ").split(",") if v.strip()]
    assert len(vals) == 64 and vals[0] == 1.0
    import math
    for j, v in enumerate(vals):
        # |v^64 - 2^j| must not shrink when v moves to a neighbouring double: v is the nearest
        def miss(x):
            return abs(Fraction(x) ** 64 - Fraction(2) ** j)
        assert miss(v) <= miss(math.nextafter(v, 0.0)) and miss(v) <= miss(math.nextafter(v, 4.0)), j
