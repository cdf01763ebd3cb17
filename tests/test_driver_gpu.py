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
    # the numbers the driver prints, not only the strings: n_eff per parameter and its ratio to the stored samples
    import re
    import numpy as np
    def array_after(label):
        txt = out[out.index(label) + len(label):]
        txt = txt[:txt.index("]") + 1]
        return np.array([float(v) for v in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", txt)])
    n_eff, ratio = array_after("Effective number per param:"), array_after("Ratio")
    assert n_eff.shape == ratio.shape == (60,)
    assert np.all(np.isfinite(n_eff)) and np.all(n_eff > 20) and np.all(n_eff <= 2.5 * 80 * 12)
    np.testing.assert_allclose(ratio, n_eff / (80 * 12), rtol=1e-6)
