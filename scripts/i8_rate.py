import sys
sys.path.insert(0, ".")
from pygps_b200._lib import Engine
eng = Engine()
for N in (64, 128, 256):
    cfgs = [(2048 if N <= 128 else 4096, 128, 0, "rowgroups adjacent (SBO=128)")]
    if N <= 128:
        cfgs += [(128, 2048, 256, "kchunks adjacent (LBO=128,SBO=2048), slices 256 B apart"), (128, 256, 0, "LBO=128,SBO=256")]
    for (lbo, sbo, astep, name) in cfgs:
        for ctas in (1, 148):
            for same in (0, 1):
                c, tops = eng.bench_i8_rate(N, 40000, lbo, sbo, astep=astep, same_acc=same, ctas=ctas)
                print(f"N={N:3d} {name:58s} ctas={ctas:3d} same_acc={same}: {c:7.1f} clk/MMA  (floor {128*N/256:.0f})  {tops:8.1f} TOP/s", flush=True)
