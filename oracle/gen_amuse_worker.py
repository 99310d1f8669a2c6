"""TEST INFRASTRUCTURE.  Generates ph4_worker.cc / ph4_worker.h with the reference's OWN code generator
(amuse.rfi.tools.create_c, what `amusifier --type=c interface.py ph4Interface` runs,
src/amuse_ph4/Makefile:30-34) from the unmodified interface specification src/amuse_ph4/interface.py.
The generated source carries `#ifndef NOMPI` guards (create_c.py:27,108,...,602): built with -DNOMPI the worker talks
to the Python side over AMUSE's sockets channel (rfi/core.py:1078-1085).  Usage: gen_amuse_worker.py <outdir>
with PYTHONPATH = tests/amuse_stub : <amuse python package>."""
import os
import sys

outdir = sys.argv[1]
from amuse.rfi.tools import create_c          # noqa: E402
import amuse_ph4.interface as iface           # noqa: E402

src = create_c.GenerateACSourcecodeStringFromASpecificationClass()
src.specification_class = iface.ph4Interface
src.needs_mpi = False
open(os.path.join(outdir, "ph4_worker.cc"), "w").write(src.result)
hdr = create_c.GenerateACHeaderStringFromASpecificationClass()
hdr.specification_class = iface.ph4Interface
hdr.needs_mpi = False
open(os.path.join(outdir, "ph4_worker.h"), "w").write(hdr.result)
print("generated ph4_worker.cc (%d bytes) and ph4_worker.h" % len(src.result))
