"""Debug helper (GPU box): CUDA engine vs host-emulation engine in lock-step, first divergence report."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_harness as ph
import b2az

def run(G, games, visits, level, seed, lanes, rng_mode, maxgen=400, pool=0):
    kw = ph.level_params(level)
    if pool: kw['pool_nodes'] = pool
    a = ph.make_engine(os.environ.get('B2AZ_DBG_LIB') or None, G, games, visits, 0, rng_mode, seed, lanes=lanes, **kw)
    b = ph.make_engine(ph.HOSTEMU_LIB, G, games, visits, 0, rng_mode, seed, **kw)
    for gen in range(maxgen):
        a.step(1); b.step(1)
        ia, ca = a.leaf_batch_host(); ib, cb = b.leaf_batch_host()
        if len(ia) == 0 and len(ib) == 0: break
        oa, ob = np.argsort(ia, kind="stable"), np.argsort(ib, kind="stable")
        bad = None
        if not np.array_equal(ia[oa], ib[ob]): bad = "ids"
        elif not np.array_equal(ca[oa], cb[ob]): bad = "canon"
        for g in range(G):
            for seat in (0, 1):
                pa, pb = a.peek(g, seat), b.peek(g, seat)
                for k in ("state", "counts", "q", "policy", "root_value", "depth", "root_n"):
                    if not np.array_equal(np.asarray(pa[k]).view(np.uint32) if np.asarray(pa[k]).dtype == np.float32 else pa[k],
                                          np.asarray(pb[k]).view(np.uint32) if np.asarray(pb[k]).dtype == np.float32 else pb[k]):
                        print(f"  lanes={lanes} rng={rng_mode} level={level} gen={gen} game={g} seat={seat} field={k}\n    gpu={pa[k]}\n    emu={pb[k]}")
                        bad = bad or "peek"
        if bad:
            print(f"FAIL lanes={lanes} rng={rng_mode} level={level} first divergence gen={gen} kind={bad} err={a.stats().device_error}")
            if bad == "canon":
                d = np.flatnonzero((ca[oa] != cb[ob]).reshape(len(ia), -1).any(1))
                print("   rows", d, "ids", ia[oa][d])
            a.close(); b.close(); return False
        v, pi = ph.fake_net(ca)
        a.submit_eval_host(ia, v, pi)
        v, pi = ph.fake_net(cb)
        b.submit_eval_host(ib, v, pi)
    print(f"ok lanes={lanes} rng={rng_mode} level={level} gens={gen}")
    a.close(); b.close(); return True

if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "all"
    if mode == "all":
        for level in (0, 1):
            for rng_mode in (1, 0):
                for lanes in (1, 8, 32, 4):
                    run(4, 6, 32, level, 12345, lanes, rng_mode, maxgen=200)
    elif mode == "pool":
        for pool in (0, 64 * 16 * 256):
            for G in (1, 2, 4, 9):
                print("pool", pool, "G", G)
                run(G, G + 2, 32, 0, 12345, 8, 0, maxgen=120, pool=pool)
    elif mode == "one":
        run(4, 6, 32, 0, 12345, 8, 0, maxgen=80)
