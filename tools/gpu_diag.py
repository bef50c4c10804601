"""First-contact GPU diagnostic: prints errors vs the oracle and rough timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openmm_sdm_plugin_b200 import _lib, system as S
from openmm_sdm_plugin_b200.context import SDMContext
from oracle import oracle as O

mode = int(sys.argv[1]) if len(sys.argv) > 1 else _lib.PAIR_ALLPAIRS
RR = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cases = (("cfg1", S.cfg1()), ("cfg2", S.cfg2())) if len(sys.argv) <= 3 else (("cfg2", S.cfg2()),)
for name, case in cases:
    ref = O.sdm_eval(case.system, S.AlchemicalState(**vars(case.alch)), case.displacement, case.positions, nthreads=O.max_threads())
    R = RR
    ctx = SDMContext(case.system, case.displacement, n_replicas=R, pair_mode=mode)
    for r in range(R):
        ctx.set_positions(r, case.positions); ctx.set_alchemical(r, case.alch)
    ctx.set_timing(True)
    ctx.eval(); ctx.synchronize()
    sc = ctx.scalars(R - 1)
    f = ctx.forces(R - 1); f1 = ctx.forces(R - 1, 1); df = ctx.forces(R - 1, 3)
    rr = lambda a, b: np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum())
    print(name, "status", sc["status"], "pairs", sc["n_pairs1"], ref["n_pairs1"], "moved", sc["n_moved1"], sc["n_moved2"])
    print("  E1", sc["E1"], ref["E1"], "rel", abs(sc["E1"] - ref["E1"]) / abs(ref["E1"]))
    print("  u", sc["u"], ref["u"], "abs", abs(sc["u"] - ref["u"]), "sp", sc["sp"], ref["sp"])
    print("  F1 rms rel", rr(f1, ref["f1"]), "dF max abs", np.abs(df - (ref["f2"] - ref["f1"])).max(), "F rms rel", rr(f, ref["forces"]))
    for k in range(5):
        ctx.eval()
    ctx.synchronize()
    t = time.perf_counter()
    for k in range(10):
        ctx.eval()
    ctx.synchronize()
    dt = (time.perf_counter() - t) / 10
    print("  R=%d eval wall %.3f ms  (pair, total device ms) %s" % (R, dt * 1e3, ctx.last_timing()))
    if mode == 2:
        print("  list:", {k: ctx.info(k) for k in ("n_slots", "n_sci", "n_entries", "n_masks", "n_units", "n_cells", "cell_span", "n_list_builds")})
    ctx.close()
