"""Generate tests/golden/*.npz from the REAL reference (run in the dev container).

TEST INFRASTRUCTURE ONLY.  Needs /root/reference and oracle/_ref (make -C oracle ref).
  * ph4_plummer1k_*.npz : reference fixture src/amuse_ph4/src/plummer1k.in pushed through the
    reference's own idata::setup() sweep (idata.cc:66-83,147-237) at eps2 = 1e-4 and 0.
  * ph4_predict_force_512.npz : jdata::predict_all (jdata.cc:710-750) + force loop on a
    block-timestep state, i-list = 64 scattered particles predicted on the host.
  * amuse_plummer_256_seed1.npz : amuse.ic.plummer.new_plummer_model(256, random=RandomState(1)),
    imported from /root/reference/src, pins amuse_b200/plummer.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from amuse_b200 import plummer as P  # noqa: E402

REF = os.environ.get("REF", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

ids, m, x, v = P.read_ph4_snapshot(os.path.join(REF, "src/amuse_ph4/src/plummer1k.in"))
for tag, eps2 in (("eps1e-4", 1e-4), ("eps0", 0.0)):
    r = O.ref_full_sweep(m, x, v, eps2, ids)
    np.savez_compressed(os.path.join(OUT, "ph4_plummer1k_%s.npz" % tag), ids=ids, mass=m, pos=x, vel=v,
                        eps2=eps2, acc=r["acc"], jerk=r["jerk"], pot=r["pot"], nn=r["nn"], dnn=r["dnn"])

# block-step state: forces at t=0 give acc/jerk, then t_j scattered in the past
n = 512
mm, xx, vv = P.new_plummer_model(n, seed=2)
f0 = O.ref_full_sweep(mm, xx, vv, 1e-4)
t = 0.125
tj, dtj = P.block_step_state(n, t, seed=3)
rnd = np.random.RandomState(5)
ilist = np.sort(rnd.choice(n, 64, replace=False))
dti = t - tj[ilist]
# i-particles predicted on the host the way idata::predict does (idata.cc:347-365)
ipos = xx[ilist] + dti[:, None] * (vv[ilist] + 0.5 * dti[:, None] * (f0["acc"][ilist] + dti[:, None] * f0["jerk"][ilist] / 3))
ivel = vv[ilist] + dti[:, None] * (f0["acc"][ilist] + 0.5 * dti[:, None] * f0["jerk"][ilist])
r = O.ref_predict_force(mm, tj, xx, vv, f0["acc"], f0["jerk"], t, 1e-4, ipos, ivel)
np.savez_compressed(os.path.join(OUT, "ph4_predict_force_512.npz"), mass=mm, pos=xx, vel=vv, acc0=f0["acc"],
                    jerk0=f0["jerk"], tj=tj, t=t, eps2=1e-4, ilist=ilist.astype(np.int32), ipos=ipos, ivel=ivel,
                    pred_pos=r["pred_pos"], pred_vel=r["pred_vel"], acc=r["acc"], jerk=r["jerk"], pot=r["pot"],
                    nn=r["nn"], dnn=r["dnn"])

# AMUSE's own Plummer generator
sys.path.insert(0, os.path.join(REF, "src"))
try:
    from amuse.ic.plummer import new_plummer_model
    from amuse.units import nbody_system
    p = new_plummer_model(256, random=np.random.RandomState(1))
    np.savez_compressed(os.path.join(OUT, "amuse_plummer_256_seed1.npz"),
                        mass=p.mass.value_in(nbody_system.mass),
                        pos=p.position.value_in(nbody_system.length),
                        vel=p.velocity.value_in(nbody_system.speed))
    print("amuse plummer fixture written")
except Exception as e:  # pragma: no cover
    print("could not import the reference plummer generator:", e)
print(sorted(os.listdir(OUT)))
