"""Golden vectors for the workload steps around the decoders (generate_data.py:53-261), recorded by running the
UNMODIFIED Python reference, seeded.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_workload.py          # writes golden_workload.npz

Per case: the reference's generate_random_error output, the uniforms it consumed (re-drawn from the same seed in
the same order), eq_true = define_equivalence_class(), and the lattice after apply_random_logical() with the numba
stream's uniforms for that call.  tests/test_workload.py feeds the uniforms to the CUDA kernels and compares.
"""
import os
import random as pyrandom
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    from oracle import refshim
    from oracle import oracle as O
    ref = refshim.load("shipped")
    rng = np.random.default_rng(20203)
    out = {}
    n = 0
    for geom in ("toric", "planar", "rotated", "xzzx"):
        for L in (3, 5, 7, 9):
            for rep in range(3):
                seed = int(rng.integers(1, 2**31))
                code = ref.codes[geom](L)
                # the defect map is not part of the vectors, and RotSurCode.syndrome() opens a plot
                code.syndrome = code.syndrom = lambda: None
                if geom == "toric":
                    p = float(rng.uniform(0.05, 0.3))
                    ref.seed_all(np_=seed)
                    code.generate_random_error(p)
                    np.random.seed(seed)
                    u, pa = [], []
                    for _ in range(2):                      # toric_model.py:16-22 draw order
                        u.append(np.random.uniform(0, 1, size=(L, L)))
                        pa.append(np.random.randint(3, size=(L, L)) + 1)
                    u, pa = np.array(u), np.array(pa, dtype=np.uint8)
                    params = np.array([p, 0, 0, 0])
                else:
                    pxyz = rng.uniform(0.01, 0.12, 3)
                    ref.seed_all(py=seed)
                    code.generate_random_error(float(pxyz[0]), float(pxyz[1]), float(pxyz[2]))
                    pyrandom.seed(seed)
                    nsites = 2 * L * L if geom == "planar" else L * L
                    u = np.array([pyrandom.random() for _ in range(nsites)])
                    pa = np.zeros(0, np.uint8)
                    params = np.array([0, pxyz[0], pxyz[1], pxyz[2]])
                q = code.qubit_matrix.astype(np.uint8).copy()
                eq_true = int(code.define_equivalence_class())
                nb_seed = int(rng.integers(1, 2**31))
                ref.seed_all(nb=nb_seed)
                q2, _ = code.apply_random_logical()
                st = O.Stream.mt(nb_seed)                   # numba's MT19937 stream, reproduced by the oracle
                u_log = np.array([st.next() for _ in range(6)])
                k = f"c{n}_"
                out[k + "geom"] = np.array(geom)
                out[k + "L"] = np.array(L)
                out[k + "params"] = params
                out[k + "u"] = u.reshape(-1)
                out[k + "pauli"] = pa.reshape(-1)
                out[k + "q"] = q.reshape(-1)
                out[k + "eq_true"] = np.array(eq_true)
                out[k + "u_log"] = u_log
                out[k + "q_hidden"] = np.asarray(q2, dtype=np.uint8).reshape(-1)
                out[k + "eq_hidden"] = np.array(int(ref.models[geom]._define_equivalence_class(np.asarray(q2, dtype=np.uint8))))
                n += 1
    # ---- PTDC with the conv_mult early stop (decoders.py:138-233), one ladder per class, droplets = 1 (in process)
    sys.path.insert(0, HERE)
    import make_golden as MG
    m = 0
    for g2 in ("toric", "planar"):
        for conv in (2.0, 1.5):
            q2 = MG.rand_lattice(rng, g2, 5, 0.1)
            if g2 == "toric":
                arr = np.array([ref.toric_model._to_class(e, q2) for e in range(16)])
            else:
                arr = np.array([c.qubit_matrix for c in MG.class_inits(ref, g2, q2)])
            seeds = [int(x) for x in rng.integers(1, 2**31, 2)]
            ref.seed_all(py=seeds[0], nb=seeds[1])
            res = ref.decoders.PTDC([MG.make_code(ref, g2, a) for a in arr], 0.1, 0.25, droplets=1, Nc=4, steps=2400, conv_mult=conv)
            k = f"p{m}_"
            out[k + "geom"] = np.array(g2)
            out[k + "inits"] = arr.astype(np.uint8)
            out[k + "conv_mult"] = np.array(conv)
            out[k + "seeds"] = np.array(seeds)
            out[k + "out"] = np.asarray(res)
            m += 1
    out["n_ptdc_conv"] = np.array(m)
    out["n_cases"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "golden_workload.npz"), **out)
    print("wrote", n, "cases")


if __name__ == "__main__":
    main()
