import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygps_b200 import _lib
e = _lib.Engine(0)
for w in (4, 8, 16):
    print("dmma only  w=%d" % w, e.bench_dmma(2, w, 4000))
    print("dfma only  w=%d" % w, e.bench_dmma(4, w, 4000))
    tf, ms = e.bench_dmma(5, w, 4000)
    print("mixed      w=%d" % w, (tf, ms), "(dmma-only with w/2 warps, same iters: %.3f ms)" % e.bench_dmma(2, w // 2, 4000)[1])
