import os, sys
os.environ["GPK_DBG_DIAG_CLK"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pygps_b200 import _lib
e = _lib.Engine(0)
rng = np.random.default_rng(1)
X = rng.standard_normal((128, 3))
d = ((X[:, None] - X[None]) ** 2).sum(-1)
A = np.exp(-0.5 * d / 4.0) / 0.01 + np.eye(128)
for i in range(3):
    e.dbg_diag(A)
