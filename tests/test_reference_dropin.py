"""Drop-in check: the reference's OWN unit-test files run against this package aliased as
``navigation_model``.  Needs /root/reference (build container only); skipped on the GPU box."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/test"
SUITES = ["test_rl.py", "test_mouse.py", "test_maze.py", "test_components.py", "test_exploration.py", "test_results.py",
          "test_fitness.py", "test_statistics.py", "test_session.py", "test_metrics.py"]

PLUGIN = """
import sys
sys.path.insert(0, {root!r})
import navigation_model_b200
navigation_model_b200.install_as("navigation_model")
"""


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference tests not available")
def test_reference_suites_pass_against_this_package(tmp_path):
    (tmp_path / "nm_dropin_plugin.py").write_text(PLUGIN.format(root=ROOT))
    env = dict(os.environ, PYTHONPATH=str(tmp_path), PYTHONDONTWRITEBYTECODE="1")
    res = subprocess.run([sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "-p", "nm_dropin_plugin", "-q",
                          *SUITES], cwd=REF_TESTS, env=env, capture_output=True, text=True, timeout=600)
    tail = res.stdout[-1500:] + res.stderr[-500:]
    assert res.returncode == 0, tail
    assert "73 passed" in res.stdout, tail
