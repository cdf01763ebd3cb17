"""A Python-2 style driver (same shape as the reference's case scripts) runs unchanged through run_case.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_case_script_runs_unchanged(tmp_path):
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "understanding-hmc_b200", "run_case.py"),
                                   os.path.join(ROOT, "tests", "data", "mini-case-script.py")], cwd=str(tmp_path),
                                  stderr=subprocess.STDOUT, timeout=300).decode()
    assert "#---- mini case ----#" in out
    assert "After warm up:" in out and "Completed." in out
    assert "Total number of samples: %d" % (80 * 12) in out
    assert "Effective number per param:" in out and "Ratio" in out
