import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    gdir = os.path.join(ROOT, "tests", "golden")

    def load(name):
        return np.load(os.path.join(gdir, name + ".npz"))
    return load


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Print the achieved error of every recorded parity comparison next to its tolerance (tests/parity_report.py)."""
    try:
        import parity_report
    except Exception:
        return
    if not parity_report.LEDGER:
        return
    tr = terminalreporter
    tr.write_sep("-", "achieved parity errors (max-norm relative)")
    worst = {}
    for e in parity_report.LEDGER:
        k = e["name"]
        if k not in worst or e["err"] > worst[k]["err"]:
            worst[k] = e
    for k in sorted(worst):
        e = worst[k]
        tr.write_line("%-58s err %.2e  tol %.1e%s" % (k, e["err"], e["tol"], ("  [" + e["why"] + "]") if e["why"] else ""))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        parity_report.dump(os.path.join(out, "parity_report.json"))
