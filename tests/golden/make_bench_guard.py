"""Expected state of the small smc run bench.py replays on N ranks before it times anything (the multi-GPU guard).

    python tests/golden/make_bench_guard.py        -> tests/golden/bench_guard.json

The run: README normal model with 100 draws, F64 simulator (bit-exact between the CPU oracle and the device), 2^14 particles,
6 iterations, default smc keywords, the benchmark seed.  The digest is computed here with the ORACLE (test infrastructure,
CPU); bench.py only compares the device result of every rank with the committed digest -- it never loads the oracle for it."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

SEED = 0x4B49535341424300
N, ITERS, NDRAWS = 1 << 14, 6, 100


def digest(th, X, lpi, alive):
    h = hashlib.sha256()
    for a in (th, X, lpi):
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    h.update(np.ascontiguousarray(alive, dtype=np.uint8).tobytes())
    return h.hexdigest()


def main():
    O.build()
    pri = O.make_priors([("uniform", 1, 3), ("truncnormal", 0, 0.1, 0, 100)])
    mod = O.make_model(O.NORMAL_MEANSTD, NDRAWS, (2.0, 0.04), (50.0,))
    s = O.Smc(SEED, pri, mod, O.smc_config(nparticles=N, max_iterations=ITERS), nthreads=os.cpu_count() or 1)
    s.init()
    for _ in range(ITERS):
        s.iterate()
    th, X, lpi, alive = s.state()
    sc = s.scalars()
    out = {"seed": SEED, "nparticles": N, "iterations": ITERS, "n_draws": NDRAWS, "precision": "f64", "sha256": digest(th, X, lpi, alive),
           "eps": sc["eps"], "cost_evals": sc["cost_evals"], "accepted": sc["accepted"],
           "generator": "tests/golden/make_bench_guard.py (CPU oracle)"}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_guard.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main()
