import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pycs_b200
from pycs_b200 import cs_datastruct, advection_ic, advection_vars, advection_timestep
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = cs_datastruct.cubed_sphere(N)
for mf in (1, 3):
  for nsteps in (1, 2, 3, 10, 100):
    tup = (3, 1, 1, 3, 1, mf)
    sims = []
    for fused in (True, False):
        sim = advection_ic.adv_simulation_par(g, 0.025*16/N, 5, 2, 1, 1, *tup)
        advection_vars.init_vars_adv(g, sim)
        advection_timestep.run_steps(g, sim, 0, nsteps, fused=fused)
        sims.append(np.asarray(sim.Q))
    d = np.abs(sims[0]-sims[1])[4:-4,4:-4,:]
    print("mf", mf, "steps", nsteps, "max err", d.max(), "at", np.unravel_index(d.argmax(), d.shape))
    dm = d.max(axis=2)
    np.set_printoptions(linewidth=250, precision=1)
    if d.max() > 1e-12 and N <= 32:
        print((dm > 1e-12).astype(int))
        break
