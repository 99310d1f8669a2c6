"""Driver script of tests/test_gpu_amuse.py: the reference's own GPU-vs-CPU test
(src/amuse_ph4/tests/test_ph4.py:848-872, test22_gpu) and the textbook performance pattern
(examples/textbook/plot_Nbody_performance.py: evolve a Plummer sphere, report wall time and dE), run through the
UNMODIFIED AMUSE framework (oracle/_ref/amuse/py) with ph4_sapporo_worker linked to the B200 library.
Prints 'AMUSE-PH4 ...' result lines."""
import json
import sys
import time

import numpy

from amuse.units import nbody_system
from amuse.ic.plummer import new_plummer_model
from amuse_ph4.interface import ph4

workers = sys.argv[1]
opts = dict(channel_type="sockets", worker_code_directory=workers, redirection="none")


def potentials(mode, particles, x, zero):
    instance = ph4(mode=mode, **opts)
    instance.initialize_code()
    instance.parameters.epsilon_squared = 0.00000 | nbody_system.length**2
    instance.particles.add_particles(particles)
    pot = instance.get_potential_at_point(zero, x, zero, zero)
    instance.stop()
    return pot.value_in(nbody_system.length**2 * nbody_system.time**-2)


# --- test22_gpu -----------------------------------------------------------------------------------
numpy.random.seed(22)
particles = new_plummer_model(200)
particles.scale_to_standard()
x = numpy.arange(-1, 1, 0.1) | nbody_system.length
zero = numpy.zeros(len(x)) | nbody_system.length
gpu = potentials("gpu", particles, x, zero)
cpu = potentials("cpu", particles, x, zero)
rel = float(numpy.max(numpy.abs(gpu - cpu) / numpy.abs(cpu)))
print("AMUSE-PH4 test22_gpu " + json.dumps({"max_rel_diff_potential": rel, "points": len(gpu)}))

# --- evolve (textbook performance pattern) ----------------------------------------------------------
res = {}
for mode, n, t_end in (("gpu", 1024, 0.25), ("cpu", 1024, 0.25), ("gpu", 16384, 0.03125)):
    numpy.random.seed(7)
    stars = new_plummer_model(n)
    stars.scale_to_standard()
    g = ph4(mode=mode, **opts)
    g.initialize_code()
    g.parameters.epsilon_squared = 0.0 | nbody_system.length**2
    g.particles.add_particles(stars)
    g.commit_particles()
    e0 = (g.potential_energy + g.kinetic_energy).value_in(nbody_system.energy)
    t0 = time.time()
    g.evolve_model(t_end | nbody_system.time)
    wall = time.time() - t0
    e1 = (g.potential_energy + g.kinetic_energy).value_in(nbody_system.energy)
    t = g.model_time.value_in(nbody_system.time)
    g.stop()
    res["%s N=%d" % (mode, n)] = {"E0": e0, "dE_over_E": abs((e1 - e0) / e0), "t": t, "wall_s": wall,
                                   "wall_s_per_unit": wall / t}
print("AMUSE-PH4 evolve " + json.dumps(res))
