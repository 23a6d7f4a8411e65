#!/usr/bin/env python
"""Per-panel timeline of the dependent chain of one evaluation (GPK_CHAIN_DUMP): python scripts/chain_dump.py [N]"""
import math, os, sys
os.environ["GPK_CHAIN_DUMP"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pygps_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
e = _lib.Engine(0)
rng = np.random.default_rng(0)
X = rng.standard_normal((N, 8)); y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
e.set_data(X)
for i in range(3):
    e.exact_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), y.reshape(-1), False)
e.set_profile(True)
e.exact_eval(_lib.COV_RBF, 3, [math.log(2.0), 0.0], math.log(0.1), y.reshape(-1), False)
print(e.stats())
