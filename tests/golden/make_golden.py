"""Generate golden vectors by running the UNMODIFIED Python reference, seeded.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # writes golden_shipped.npz, golden_toric.npz

Each variant runs in its own process (oracle/refshim.py explains why).  Every
case records the inputs, the three RNG seeds (CPython / numba / numpy -- the
reference itself never seeds) and the reference's outputs.
tests/test_oracle_golden.py replays all of them through oracle/qec_oracle.c.
"""
import copy
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def rand_lattice(rng, geom, L, p):
    shape = (2, L, L) if geom in ("toric", "planar") else (L, L)
    q = (rng.random(shape) < p) * rng.integers(1, 4, shape)
    q = q.astype(np.uint8)
    if geom == "planar":
        q[1, -1, :] = 0
        q[1, :, -1] = 0
    return q


def make_code(ref, geom, q):
    code = ref.codes[geom](q.shape[-1])
    code.qubit_matrix = q.copy()
    return code


def class_inits(ref, geom, q):
    """Per-class init codes via apply_logical(class ^ eq) (decoders.py:556-560 route)."""
    out = []
    for eq in range(ref.codes[geom].nbr_eq_classes):
        code = make_code(ref, geom, q)
        op = code.define_equivalence_class() ^ eq
        code.qubit_matrix = code.apply_logical(op)[0]
        assert code.define_equivalence_class() == eq
        out.append(code)
    return out


def gen(variant):
    from oracle import refshim
    ref = refshim.load(variant)
    rng = np.random.default_rng(20201 if variant == "shipped" else 20202)
    cases = []

    def add(kind, **kw):
        kw["kind"] = kind
        cases.append(kw)

    geoms = ["toric", "planar", "rotated", "xzzx"]
    if variant == "shipped":
        # ---- A. primitives -------------------------------------------------
        for geom in geoms:
            for L in (3, 5, 7):
                mod = ref.models[geom]
                q = rand_lattice(rng, geom, L, 0.4)
                # all stabilizers
                stabs = []
                if geom == "toric":
                    stabs = [(r, c, op) for op in (1, 3) for r in range(L) for c in range(L)]
                elif geom == "planar":
                    stabs = [(r, c, 1) for r in range(L - 1) for c in range(L)] + \
                            [(r, c, 3) for r in range(L) for c in range(L - 1)]
                else:
                    stabs = [(r, c, 1) for r in range(L - 1) for c in range(L - 1)] + \
                            [(k, s, 3) for k in range((L - 1) // 2) for s in range(4)]
                outs, ds = [], []
                for (r, c, op) in stabs:
                    m, d = mod._apply_stabilizer(q, r, c, op)
                    outs.append(m)
                    ds.append(d)
                add("apply_stabilizer", geom=geom, L=L, q=q, stabs=np.array(stabs), out=np.array(outs),
                    dE=np.array(ds))
                # logicals
                logs, louts, lds = [], [], []
                for op in range(4):
                    for layer in ((0, 1) if geom == "toric" else (0,)):
                        for xp in range(L):
                            zp = (xp * 2 + 1) % L
                            if geom == "toric":
                                m, d = mod._apply_logical(q, op, layer, xp, zp)
                            else:
                                m, d = mod._apply_logical(q, op, xp, zp)
                            logs.append((op, layer, xp, zp))
                            louts.append(m)
                            lds.append(d)
                add("apply_logical", geom=geom, L=L, q=q, args=np.array(logs), out=np.array(louts),
                    dE=np.array(lds))
                # classes
                qs = np.array([rand_lattice(rng, geom, L, 0.3) for _ in range(24)])
                add("class", geom=geom, L=L, q=qs,
                    out=np.array([mod._define_equivalence_class(x) for x in qs]))
                # seeded random stabilizers / logicals applied cumulatively
                for fn_name in ("_apply_random_stabilizer", "_apply_random_logical"):
                    seed = int(rng.integers(1, 2**31))
                    ref.seed_all(nb=seed)
                    cur, ds = q.copy(), []
                    for _ in range(60):
                        cur, d = getattr(mod, fn_name)(cur)
                        ds.append(d)
                    add("random_" + fn_name.split("_")[-1], geom=geom, L=L, q=q, nb_seed=seed, n=60, out=cur,
                        dE=np.array(ds))
            if geom == "toric":
                for L in (3, 5):
                    q = rand_lattice(rng, geom, L, 0.3)
                    add("to_class", geom=geom, L=L, q=q,
                        out=np.array([ref.toric_model._to_class(e, q) for e in range(16)]))
            if geom in ("toric", "planar"):
                for L in (4, 5):
                    q = rand_lattice(rng, geom, L, 0.3)
                    seed = int(rng.integers(1, 2**31))
                    ref.seed_all(np_=seed)
                    add("rain", geom=geom, L=L, q=q, np_seed=seed,
                        out=ref.models[geom]._apply_stabilizers_uniform(q, 0.5))

        # ---- B. slow-path chains, ladders -----------------------------------
        for geom in geoms:
            for L in (5, 7):
                q = rand_lattice(rng, geom, L, 0.15)
                for (p, p_logical) in ((0.2, 0.0), (0.3, 0.5), (0.75, 0.5)):
                    seeds = [int(x) for x in rng.integers(1, 2**31, 2)]
                    ref.seed_all(py=seeds[0], nb=seeds[1])
                    ch = ref.mcmc.Chain(p, make_code(ref, geom, q))
                    ch.p_logical = p_logical
                    snaps = []
                    for _ in range(30):
                        ch.update_chain(10)
                        snaps.append(ch.code.qubit_matrix.copy())
                    add("update_chain", geom=geom, L=L, q=q, p=p, p_logical=p_logical, py_seed=seeds[0],
                        nb_seed=seeds[1], iters=10, blocks=30, out=np.array(snaps))
                # weighted chains (frozen pb)
                for (kind, a, b) in (("alpha", 0.3, 2.0), ("alpha", 0.2, 1.5), ("biased", 0.3, 10.0), ("biased", 0.45, 1.0)):
                    for p_logical in (0.0, 0.5):
                        seeds = [int(x) for x in rng.integers(1, 2**31, 2)]
                        ref.seed_all(py=seeds[0], nb=seeds[1])
                        if kind == "alpha":
                            ch = ref.mcmc_alpha.Chain_alpha(a, b, make_code(ref, geom, q))
                        else:
                            ch = ref.mcmc_biased.Chain_biased(a, b, make_code(ref, geom, q))
                        ch.p_logical = p_logical
                        snaps, neff = [], []
                        for _ in range(30):
                            ch.update_chain(10)
                            snaps.append(ch.code.qubit_matrix.copy())
                            neff.append(float(getattr(ch, "n_eff", 0.0)))
                        add("update_chain_weighted", wkind=kind, geom=geom, L=L, q=q, a=a, b=b,
                            p_logical=p_logical, py_seed=seeds[0], nb_seed=seeds[1], iters=10, blocks=30,
                            out=np.array(snaps), n_eff=np.array(neff))
                # ladders
                for lk in ("dep", "alpha", "biased"):
                    seeds = [int(x) for x in rng.integers(1, 2**31, 2)]
                    ref.seed_all(py=seeds[0], nb=seeds[1])
                    Nc = 5
                    if lk == "dep":
                        bottom, b = 0.15, 0.0
                        lad = ref.mcmc.Ladder(bottom, make_code(ref, geom, q), Nc, 0.5)
                    elif lk == "alpha":
                        bottom, b = 0.2, 2.0
                        lad = ref.mcmc_alpha.Ladder_alpha(bottom, make_code(ref, geom, q), b, Nc, 0.5)
                    else:
                        bottom, b = 0.2, 10.0
                        lad = ref.mcmc_biased.Ladder_biased(bottom, make_code(ref, geom, q), b, Nc, 0.5)
                    states, flags, tops = [], [], []
                    for _ in range(40):
                        lad.step(10)
                        states.append(np.array([c.code.qubit_matrix for c in lad.chains]))
                        flags.append([c.flag for c in lad.chains])
                        tops.append(lad.tops0)
                    add("ladder", lkind=lk, geom=geom, L=L, q=q, bottom=bottom, b=b, Nc=Nc, p_logical=0.5,
                        py_seed=seeds[0], nb_seed=seeds[1], iters=10, nsteps=40, states=np.array(states),
                        flags=np.array(flags), tops0=np.array(tops))

        # ---- C. PT drivers ---------------------------------------------------
        for geom in geoms:
            L = 5
            q = rand_lattice(rng, geom, L, 0.1)
            seeds = [int(x) for x in rng.integers(1, 2**31, 2)]
            ref.seed_all(py=seeds[0], nb=seeds[1])
            out = ref.decoders.PTEQ(make_code(ref, geom, q), 0.1, steps=100000)
            add("pteq", lkind="dep", geom=geom, L=L, q=q, p=0.1, b=0.0, py_seed=seeds[0], nb_seed=seeds[1],
                steps=100000, out=out)
            seeds = [int(x) for x in rng.integers(1, 2**31, 2)]
            ref.seed_all(py=seeds[0], nb=seeds[1])
            out = ref.decoders_biasednoise.PTEQ_biased(make_code(ref, geom, q), 0.1, eta=10.0, steps=100000)
            add("pteq", lkind="biased", geom=geom, L=L, q=q, p=0.1, b=10.0, py_seed=seeds[0], nb_seed=seeds[1],
                steps=100000, out=out)
            seeds = [int(x) for x in rng.integers(1, 2**31, 2)]
            ref.seed_all(py=seeds[0], nb=seeds[1])
            out = ref.decoders_biasednoise.PTEQ_alpha(make_code(ref, geom, q), 0.15, alpha=2.0, steps=100000)
            add("pteq", lkind="alpha", geom=geom, L=L, q=q, p=0.15, b=2.0, py_seed=seeds[0], nb_seed=seeds[1],
                steps=100000, out=out)
            if geom != "toric":  # raises on Toric_code as shipped
                seeds = [int(x) for x in rng.integers(1, 2**31, 2)]
                ref.seed_all(py=seeds[0], nb=seeds[1])
                out = ref.decoders.STDC_Nall_n_alpha(make_code(ref, geom, q), pz_tilde_sampling=0.3, alpha=2.0,
                                                     pz_tilde=0.1, steps=400)
                add("stdc_alpha", geom=geom, L=L, q=q, pz_tilde_sampling=0.3, alpha=2.0, pz_tilde=0.1, steps=400,
                    py_seed=seeds[0], nb_seed=seeds[1], out=out)

    # ---- D. fast path + single-temperature drivers (both variants) ----------
    # shipped: proposal geometry is planar whatever the code;
    # toric:   proposal geometry is toric.
    chain_geom = "planar" if variant == "shipped" else "toric"
    for geom in (["planar", "toric"] if variant == "shipped" else ["toric"]):
        for L in (5, 7):
            q = rand_lattice(rng, geom, L, 0.15)
            seed = int(rng.integers(1, 2**31))
            ref.seed_all(nb=seed)
            ch = ref.mcmc.Chain(0.25, make_code(ref, geom, q))
            snaps = []
            for _ in range(400):
                ch.update_chain_fast(5)
                snaps.append(ch.code.qubit_matrix.copy())
            add("chain_fast", geom=geom, chain_geom=chain_geom, L=L, q=q, p=0.25, nb_seed=seed, iters=5, blocks=400,
                out=np.array(snaps))
            # drivers, droplets=1 (in-process, so the seeded streams are the ones used)
            if geom == "toric":
                init = lambda: make_code(ref, geom, q)
            else:
                init = lambda: class_inits(ref, geom, q)
            inits_arr = np.array([c.qubit_matrix for c in class_inits(ref, geom, q)]) if geom != "toric" else \
                np.array([ref.toric_model._to_class(e, q) for e in range(16)])
            for conv_mult in (0, 2.0):
                seeds = [int(x) for x in rng.integers(1, 2**31, 2)]
                ref.seed_all(nb=seeds[0], np_=seeds[1])
                out = ref.decoders.STDC(init(), 0.1, 0.25, droplets=1, steps=300, conv_mult=conv_mult)
                add("stdc", geom=geom, chain_geom=chain_geom, L=L, q=q, inits=inits_arr, p_error=0.1, p_sampling=0.25,
                    steps=300, conv_mult=conv_mult, randomize=int(geom == "toric"), nb_seed=seeds[0],
                    np_seed=seeds[1], out=out)
                seeds = [int(x) for x in rng.integers(1, 2**31, 2)]
                ref.seed_all(nb=seeds[0], np_=seeds[1])
                out = ref.decoders.STRC(init(), 0.1, 0.25, droplets=1, steps=300, conv_mult=conv_mult)
                add("strc", geom=geom, chain_geom=chain_geom, L=L, q=q, inits=inits_arr, p_error=0.1, p_sampling=0.25,
                    steps=300, conv_mult=conv_mult, randomize=int(geom == "toric"), nb_seed=seeds[0],
                    np_seed=seeds[1], out=out)
            seed = int(rng.integers(1, 2**31))
            ref.seed_all(nb=seed)
            out = ref.decoders.single_temp(init(), 0.2, 200)
            add("single_temp", geom=geom, chain_geom=chain_geom, L=L, q=q, inits=inits_arr, p=0.2, max_iters=200,
                nb_seed=seed, out=out)

    # ---- E. general noise, PTDC / PTRC, PTEQ_alpha_with_shortest (shipped variant only) ----
    if variant == "shipped":
        geom, L = "planar", 5
        q = rand_lattice(rng, geom, L, 0.12)
        inits_arr = np.array([c.qubit_matrix for c in class_inits(ref, geom, q)])
        p_xyz = np.array([0.03, 0.02, 0.06])
        # Chain_xyz trajectory (fast path, planar proposals)
        seed = int(rng.integers(1, 2**31))
        ref.seed_all(nb=seed)
        ps = np.array([0.08, 0.06, 0.10])
        ch = ref.mcmc.Chain_xyz(ps, make_code(ref, geom, q))
        snaps = []
        for _ in range(200):
            ch.update_chain_fast(5)
            snaps.append(ch.code.qubit_matrix.copy())
        add("chain_fast_xyz", geom=geom, chain_geom="planar", L=L, q=q, p_sampling=ps, nb_seed=seed, iters=5, blocks=200,
            out=np.array(snaps))
        for ps in (None, np.array([0.08, 0.06, 0.10])):
            seed = int(rng.integers(1, 2**31))
            ref.seed_all(nb=seed)
            out, out_s = ref.decoders.STDC_general_noise_shortest(class_inits(ref, geom, q), p_xyz, ps, droplets=1, steps=300)
            add("stdc_general_noise", geom=geom, chain_geom="planar", L=L, q=q, inits=inits_arr, p_xyz=p_xyz,
                p_sampling=(np.zeros(0) if ps is None else ps), steps=300, nb_seed=seed, out=out, out_shortest=out_s)
        # PTDC / PTRC: one ladder per class, droplets = 1 (in process)
        for g2 in ("toric", "planar"):
            q2 = rand_lattice(rng, g2, 5, 0.1)
            if g2 == "toric":
                arr = np.array([ref.toric_model._to_class(e, q2) for e in range(16)])
            else:
                arr = np.array([c.qubit_matrix for c in class_inits(ref, g2, q2)])

            def lst():
                return [make_code(ref, g2, a) for a in arr]
            seeds = [int(x) for x in rng.integers(1, 2**31, 2)]
            ref.seed_all(py=seeds[0], nb=seeds[1])
            out = ref.decoders.PTDC(lst(), 0.1, 0.25, droplets=1, Nc=4, steps=1200)
            add("ptdc", geom=g2, L=5, q=q2, inits=arr, p_error=0.1, p_sampling=0.25, Nc=4, steps=1200, py_seed=seeds[0],
                nb_seed=seeds[1], out=out)
            seeds = [int(x) for x in rng.integers(1, 2**31, 2)]
            ref.seed_all(py=seeds[0], nb=seeds[1])
            out = ref.decoders.PTRC(lst(), 0.1, 0.25, droplets=1, Nc=4, steps=1200)
            add("ptrc", geom=g2, L=5, q=q2, inits=arr, p_error=0.1, p_sampling=0.25, Nc=4, steps=1200, py_seed=seeds[0],
                nb_seed=seeds[1], out=out)
        for g2 in ("planar", "rotated", "xzzx"):
            q2 = rand_lattice(rng, g2, 5, 0.1)
            seeds = [int(x) for x in rng.integers(1, 2**31, 2)]
            ref.seed_all(py=seeds[0], nb=seeds[1])
            o1, o2, o3 = ref.decoders_biasednoise.PTEQ_alpha_with_shortest(make_code(ref, g2, q2), 0.15, alpha=2.0, steps=100000)
            add("pteq_alpha_shortest", geom=g2, L=5, q=q2, p=0.15, b=2.0, py_seed=seeds[0], nb_seed=seeds[1], steps=100000,
                out=o1, out_unique=o2, out_shortest_n=o3)

    arrays, manifest = {}, []
    for i, c in enumerate(cases):
        meta = {}
        for k, v in c.items():
            if isinstance(v, np.ndarray):
                arrays[f"{i}.{k}"] = v
            else:
                meta[k] = v
        manifest.append(meta)
    arrays["manifest"] = np.array(json.dumps(manifest))
    path = os.path.join(HERE, f"golden_{variant}.npz")
    np.savez_compressed(path, **arrays)
    print(variant, len(cases), "cases ->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        gen(sys.argv[1])
    else:
        for v in ("shipped", "toric"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), v])
