"""Worker of tests/test_gpu_multidev.py: ONE process drives k devices through the g6 C ABI (G6_B200_DEVICES=k:
j-addresses dealt out in chunks of 256, partial forces gathered at device 0 over peer memory) and compares
with the FP64 oracle and with the same calls on one device.  Prints 'MULTI-DEVICE OK ...' on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from amuse_b200 import g6lib, plummer as P  # noqa: E402
from helpers import check_forces, check_nn  # noqa: E402
from oracle import oracle as O  # noqa: E402


def run(k, ids, m, x, v, eps2, blocks):
    os.environ["G6_B200_DEVICES"] = str(k)
    g = g6lib.G6(0)
    assert g.L.g6x_device_count_open() == k, "opened %d devices, wanted %d" % (g.L.g6x_device_count_open(), k)
    # one particle at a time for a part (g6_set_j_particle_ routes to the owner), batched for the rest
    n = len(m)
    z = np.zeros(3)
    for a in range(0, 700):
        g.set_j_particle(a, int(ids[a]), 0.0, 0.0, m[a], z, z, z, v[a], x[a])
    g.set_j_particles(ids[700:], m[700:], x[700:], v[700:], address0=700)
    g.nj = n
    g.set_ti(0.0)
    res = {}
    for name, sel in blocks.items():
        res[name] = g.calc(ids[sel], x[sel], v[sel], eps2)
    # neighbour lists of the last block (merged over the devices)
    sel = blocks["small"]
    ref = O.force(x[sel], v[sel], m, x, v, eps2, iid=ids[sel], jid=ids)
    h2 = np.minimum(8 * ref["dnn"] ** 2, 1.0)
    out = g.calc(ids[sel], x[sel], v[sel], eps2, h2=h2)
    ovf = g.read_neighbour_list()
    lists = [g.get_neighbour_list(i)[2] for i in range(len(sel))]
    # a j-update (block-step pattern) must reach the owner and be seen by the next call
    xs = x.copy()
    xs[sel] += 1e-3
    for a in sel:
        g.set_j_particle(int(a), int(ids[a]), 0.0, 0.0, m[a], z, z, z, v[a], xs[a])
    res["after update"] = g.calc(ids[sel], xs[sel], v[sel], eps2)
    g.close()
    return res, lists, ovf, xs


def main():
    kmax = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    m, x, v = P.new_plummer_model(n, seed=4)
    ids = np.arange(1, n + 1, dtype=np.int32)
    rng = np.random.RandomState(3)
    blocks = {"all": np.arange(n), "big": np.sort(rng.choice(n, 5000, replace=False)),
              "mid": np.sort(rng.choice(n, 700, replace=False)), "small": np.sort(rng.choice(n, 37, replace=False))}
    eps2 = 0.0
    one, lists1, ovf1, xs = run(1, ids, m, x, v, eps2, blocks)
    msgs = []
    for k in sorted(set([2, kmax])):
        many, lists, ovf, _ = run(k, ids, m, x, v, eps2, blocks)
        for name, sel in list(blocks.items()) + [("after update", blocks["small"])]:
            xi = xs if name == "after update" else x
            xj = xs if name == "after update" else x
            ref = O.force(xi[sel], v[sel], m, xj, v, eps2, iid=ids[sel], jid=ids)
            ea, ej, ep = check_forces(many[name], ref, what="%d devices, block %s" % (k, name))
            check_nn(many[name]["nn"], ref["nn"], ids, xi[sel], xj)
            assert np.array_equal(many[name]["nn"], one[name]["nn"]), "nn differs from the one-device run (%s)" % name
            d = np.linalg.norm(many[name]["acc"] - one[name]["acc"], axis=1) / np.linalg.norm(one[name]["acc"], axis=1)
            assert d.max() < 1.5e-6, "acc differs from the one-device run by %.2e (%s)" % (d.max(), name)
            msgs.append("%s acc %.1e jerk %.1e pot %.1e" % (name, ea, ej, ep))
        assert ovf == ovf1 and all(np.array_equal(a, b) for a, b in zip(lists, lists1)), "neighbour lists differ"
        print("MULTI-DEVICE OK: %d devices in one process, N=%d: %s; neighbour lists of %d particles equal to the "
              "one-device lists (mean length %.1f)" % (k, n, "; ".join(msgs), len(lists), np.mean([len(t) for t in lists])))


if __name__ == "__main__":
    main()
