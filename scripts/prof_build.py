import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import lisa_b200.rt as rt
from oracle.make_golden_gpu import soup
T = int(float(sys.argv[1]))
v, n = soup(T)
m = np.zeros(T, np.int32)
mats = [dict(emit=False, alpha=1.0, diffuse=(0.7, 0.7, 0.7), roughness=1.0)]
R = rt.Renderer(v, n, m, mats, 64, 64, (0.5, 0.6, 3.2), (0.5, 0.45, 0.5), 35.0, 1, 7)
print(R.stats()["bvh_build_ms"])
