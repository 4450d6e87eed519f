#!/usr/bin/env python
"""oracle/gen_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.json by RUNNING THE REFERENCE ITSELF (oracle/_ref/libntt_ref.so, i.e. the
sources under /root/reference compiled by oracle/Makefile).  Run in the build container:

    make -C oracle && python oracle/gen_golden.py

What is pinned (all from reference outputs, none from our own restatement):
  * the 19 fixture parameter sets of tests/test_cases.h:145-208 (read through the compiled header);
  * FNV-1a-64 hashes of the reference tables (calc_w / calc_w_con, pre_compute.h:38-77);
  * hashes of fwd_ntt_ref_harvey / fwd_ntt_ref_harvey_lazy / inv_ntt_ref_harvey outputs on
    - the full-range splitmix64 input of SURVEY.md Appendix C (seed 0x5EED0000+idx),
    - a lazy-range input in [0,4q) for forward and [0,2q) for inverse (the reference's input contracts),
    - edge vectors: all-zero, all q-1, delta at 0, delta at N-1;
  * complete input/output/table vectors for case 0 (N=256) so a mismatch can be localised;
  * the synthetic throughput parameter sets of SURVEY.md Appendix D (49-bit q, N = 2^13, 2^14, 2^16).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyoracle import Oracle, Reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def hx(v):
    return "%016x" % v


def edge_inputs(N, q):
    z = np.zeros(N, dtype=np.uint64)
    d0 = z.copy(); d0[0] = 1
    dl = z.copy(); dl[N - 1] = 1
    return {"zero": z, "qm1": np.full(N, q - 1, dtype=np.uint64), "delta0": d0, "deltaN": dl}


def main():
    ref, orc = Reference(), Oracle()
    assert ref.available, "build oracle/_ref first (make -C oracle)"
    fnv, uni = orc.fnv, orc.uniform   # hashing / input generation only; every transform below is ref.*
    cases = []
    for c in ref.cases():
        N, q = 1 << c["m"], c["q"]
        w, wc = ref.tables(N, q, c["w"])
        wi, wic = ref.tables(N, q, c["w_inv"])
        a = uni(N, q, 0x5EED0000 + c["idx"])
        f = ref.fwd(a, q, w, wc)
        fl = ref.fwd_lazy(a, q, w, wc)
        i = ref.inv(f, q, c["n_inv"], wi, wic)
        assert (i == a).all()
        a4 = uni(N, 4 * q, 0x4A2F0000 + c["idx"])      # forward input contract: [0,4q)
        a2 = uni(N, 2 * q, 0x2A2F0000 + c["idx"])      # inverse input contract: [0,2q)
        entry = dict(c)
        entry.update(
            n_inv_con=ref.ninv_con(c["n_inv"], q),
            w_fnv=hx(fnv(w)), w_con_fnv=hx(fnv(wc)), w_inv_fnv=hx(fnv(wi)), w_inv_con_fnv=hx(fnv(wic)),
            in_fnv=hx(fnv(a)), fwd_fnv=hx(fnv(f)), fwd_lazy_fnv=hx(fnv(fl)),
            in4q_fnv=hx(fnv(a4)), fwd4q_fnv=hx(fnv(ref.fwd(a4, q, w, wc))),
            in2q_fnv=hx(fnv(a2)), inv2q_fnv=hx(fnv(ref.inv(a2, q, c["n_inv"], wi, wic))),
            edges={k: dict(fwd=hx(fnv(ref.fwd(v, q, w, wc))), inv=hx(fnv(ref.inv(v, q, c["n_inv"], wi, wic))))
                   for k, v in edge_inputs(N, q).items()},
        )
        cases.append(entry)
        if c["idx"] == 0:
            full = dict(m=c["m"], q=q, psi=c["w"], psi_inv=c["w_inv"], n_inv=c["n_inv"],
                        w=[int(x) for x in w], w_con=[int(x) for x in wc],
                        w_inv=[int(x) for x in wi], w_inv_con=[int(x) for x in wic],
                        a=[int(x) for x in a], fwd=[int(x) for x in f], fwd_lazy=[int(x) for x in fl])
            with open(os.path.join(OUT, "case0_full.json"), "w") as fh:
                json.dump(full, fh)
    with open(os.path.join(OUT, "cases.json"), "w") as fh:
        json.dump(dict(source="reference build oracle/_ref/libntt_ref.so; generator oracle/gen_golden.py",
                       input="splitmix64(seed) % bound, seeds 0x5EED0000+idx ([0,q)), 0x4A2F0000+idx ([0,4q)), "
                             "0x2A2F0000+idx ([0,2q)); hash FNV-1a-64 over little-endian u64",
                       cases=cases), fh, indent=1)

    # synthetic throughput parameter sets; psi = smallest primitive 2N-th root
    q49 = 0x1FFFFFC800001
    synth = []
    for m, batch, seed in ((13, 4, 3), (14, 4, 1), (16, 2, 2)):
        N = 1 << m
        psi = orc.min_root(N, q49)
        psi_inv, n_inv = orc.invmod(psi, q49), orc.invmod(N, q49)
        w, wc = ref.tables(N, q49, psi)
        wi, wic = ref.tables(N, q49, psi_inv)
        a = uni(batch * N, q49, seed).reshape(batch, N)
        f = ref.fwd(a, q49, w, wc)
        i = ref.inv(f, q49, n_inv, wi, wic)
        assert (i == a).all()
        synth.append(dict(m=m, q=q49, psi=psi, psi_inv=psi_inv, n_inv=n_inv, batch=batch, seed=seed,
                          in_fnv=hx(fnv(a)), fwd_fnv=hx(fnv(f)), w_fnv=hx(fnv(w)), w_con_fnv=hx(fnv(wc))))
    with open(os.path.join(OUT, "synthetic.json"), "w") as fh:
        json.dump(dict(source="reference build oracle/_ref/libntt_ref.so", sets=synth), fh, indent=1)
    print("wrote", OUT, len(cases), "cases,", len(synth), "synthetic sets")


if __name__ == "__main__":
    main()
