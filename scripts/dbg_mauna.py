import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pygps_b200 as pg
from pygps_b200 import _lib
from oracle import gp_oracle as go
from test_oracle import MAUNA
from test_gpu_programs import build
g = np.load("tests/golden/cov_programs.npz")
X, Y = g["mauna_x"], g["mauna_y"]
k = build(MAUNA)
Kd = k.getCovMatrix(x=X, mode="train")
Kr = go.cov_matrix(MAUNA, x=X, mode="train")
print("K max rel err", np.abs(Kd - Kr).max() / np.abs(Kr).max(), "asym", np.abs(Kd - Kd.T).max())
A = Kr / 0.01 + np.eye(X.shape[0])
print("min eig", np.linalg.eigvalsh(A)[:3], "cond", np.linalg.cond(A))
eng = _lib.Engine(0)
try:
    R, ld = eng.potrf(A)
    print("potrf ok", ld, np.log(np.diag(np.linalg.cholesky(A))).sum())
except Exception as e:
    print("potrf failed", e)
for n in (128, 200, 256, 300, 400, 545):
    m = pg.GPR(); m.setData(X[:n], Y[:n]); m.setPrior(kernel=build(MAUNA))
    try:
        print(n, m.getPosterior(der=False)[0], go.exact_evaluate(("const", float(np.mean(Y[:n]))), MAUNA, np.log(0.1), X[:n], Y[:n], 2)[1])
    except Exception as e:
        print(n, "failed", e)
# native RBF with the same scale / conditioning
for n in (300, 545):
    m = pg.GPR(); m.setData(X[:n], Y[:n]); m.setPrior(kernel=pg.cov.RBF(np.log(67.), np.log(66.)))
    try:
        print("rbf", n, m.getPosterior(der=False)[0], go.exact_evaluate(("const", float(np.mean(Y[:n]))), ("rbf", [np.log(67.), np.log(66.)]), np.log(0.1), X[:n], Y[:n], 2)[1])
    except Exception as e:
        print("rbf", n, "failed", e)
