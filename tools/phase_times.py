import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from nbodygo_b200 import capi, clouds
for name, n in (("C4", 1000), ("C4", 4000), ("C2", 10000), ("C4", 16000)):
    b = clouds.config(name, n=n)
    sim = capi.Sim(n + 64); sim.upload(b)
    for _ in range(5): sim.step(1e-9, 1.0)
    acc = np.zeros(6); K = 20
    for _ in range(K):
        r = sim.step(1e-9, 1.0, capi.STEP_DEFAULT | capi.STEP_PHASE_TIMINGS)
        acc += [r.ms_prep, r.ms_force, r.ms_exchange, r.ms_resolve, r.ms_integrate, r.ms_total]
    acc /= K
    tot = 0.0
    for _ in range(K):
        tot += sim.step(1e-9, 1.0).ms_total
    print(f"{name} n={n}: prep {acc[0]*1e3:.1f} force {acc[1]*1e3:.1f} exch {acc[2]*1e3:.1f} resolve {acc[3]*1e3:.1f} integ {acc[4]*1e3:.1f} total(phase-timed) {acc[5]*1e3:.1f} us | graph replay total {tot/K*1e3:.1f} us")
    sim.close()
