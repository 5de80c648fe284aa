import sys, os
sys.path.insert(0, os.getcwd())
from nbodygo_b200 import capi, clouds
n = int(sys.argv[1])
b = clouds.config("C4", n=n)
sim = capi.Sim(n + 64); sim.upload(b)
for _ in range(6): sim.step(1e-9, 1.0)
sim.close()
