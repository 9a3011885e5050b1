"""Stage-by-stage GPU-vs-oracle comparison used while bringing kernels up (run under gpurun)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle_lib as O
from mpcith_kyber_kosk_b200 import KoskContext

def main():
    ks = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2, 3, 4]
    nseeds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    allok = True
    for k in ks:
        L = O.layout(k)
        ctx = KoskContext(k, 0, 8)
        # components
        rng = np.random.default_rng(k)
        y = rng.integers(0, 3329, size=(5, 407), dtype=np.uint16)
        sh = ctx.share_eval(y)
        okc = all((sh[i] == O.oracle_share(y[i])).all() for i in range(5))
        a = rng.integers(0, 3329, size=(4, 256), dtype=np.uint16)
        okn = (ctx.ntt_rows(a) == np.stack([O.oracle_ntt(r) for r in a])).all()
        import hashlib
        rows = rng.integers(0, 256, size=(7, 308), dtype=np.uint8)
        okh = all(bytes(h) == hashlib.sha3_256(bytes(r)).digest() for h, r in zip(ctx.sha3_256_rows(rows), rows))
        print(f"K={k} components: share_eval={okc} ntt={okn} sha3={okh}", flush=True)
        allok &= okc and okn and okh
        seeds = np.stack([np.frombuffer(O.seed_of(i), np.uint8) for i in range(nseeds)])
        t0 = time.time(); pk, sk, pi = ctx.prove_batch(seeds); t1 = time.time() - t0
        for i in range(nseeds):
            opk, osk, opi = O.oracle_prove(k, seeds[i])
            tr = O.oracle_trace()
            same = ((pk[i] == opk).all(), (sk[i] == osk).all(), (pi[i] == opi).all())
            print(f"K={k} seed{i}: pk={same[0]} sk={same[1]} proof={same[2]} ({t1:.3f}s batch)", flush=True)
            allok &= all(same)
            if not all(same):
                offs = [(n, getattr(L, n)) for n in O.FIELDS] + [("end", L.proof_bytes)]
                for (n, o), (_, e) in zip(offs[:-1], offs[1:]):
                    d = np.nonzero(pi[i][o:e] != opi[o:e])[0]
                    if d.size: print(f"   field {n}: {d.size}/{e-o} bytes differ, first at +{d[0]}")
                if i == nseeds - 1 or True:
                    F = L.F; NA = 70 + 2 * k
                    pw = ctx.debug_fetch("alpha_pow", 8 * NA * F * 2, np.uint16).reshape(8, NA, F)
                    print("   alpha ok:", (pw[i, :, 1] == np.array(tr.alpha[:NA])).all())
                    I = ctx.debug_fetch("I", 8 * 150 * 2, np.uint16).reshape(8, 150)
                    print("   I ok:", (I[i] == np.array(tr.I[:])).all())
                    tc = ctx.debug_fetch("tcomm", 8 * 1454 * 32).reshape(8, 1454, 32)
                    print("   tcomm0 ok:", bytes(tc[i, 0]) == bytes(tr.tcomm0))
                    vw = ctx.debug_fetch("views", 8 * 1454 * 32).reshape(8, 1454, 32)
                    print("   view0 ok:", bytes(vw[i, 0]) == bytes(tr.view0))
                    n2 = 2 * F + 2 * k * L.E + 2 * k + 2 * k * L.M + 4 * k
                    Y = ctx.debug_fetch("Y", 8 * n2 * 416 * 2, np.uint16).reshape(8, n2, 416)
                    print("   f0 secret ok:", (Y[i, 0, :256] == np.array(tr.first_secret[:])).all(), " ntt ok:", (Y[i, F, :256] == np.array(tr.first_ntt[:])).all())
                    nslot = n2 + 2 * k * 2 + 2 * k + 2 * k * L.M + 2 * k + 16
                    P = ctx.debug_fetch("planes", 8 * nslot * 1456 * 2, np.uint16).reshape(8, nslot, 1456)
                    fs = np.array(tr.first_share[:])
                    print("   f0 tail ok:", (P[i, 0, 1:152] == fs[:151]).all(), " f0 evals ok:", (P[i, 0, 152:1455] == fs[151:]).all())
        # verifier
        opk, osk, opi = O.oracle_prove(k, seeds[0])
        res = ctx.kosk_verify(bytes(opi), bytes(opk))
        fl = ctx.debug_fetch("vflags", 4, np.int32)[0]
        print(f"K={k} verify(honest)={res} flags={fl}", flush=True)
        allok &= bool(res)
        # tamper matrix: first and last element of each field, bit 0
        offs = [(n, getattr(L, n)) for n in O.FIELDS] + [("end", L.proof_bytes)]
        cases = []
        for (n, o), (_, e) in zip(offs[:-1], offs[1:]):
            step = 1 if n in ("o_Tcomm", "o_comm") else 2
            cases.append((n + ":first", o)); cases.append((n + ":last", e - step))
        tp = np.repeat(opi[None, :], len(cases), axis=0).copy()
        for ci, (_, off) in enumerate(cases): tp[ci, off] ^= 1
        got = ctx.verify_batch(tp, np.repeat(opk[None, :], len(cases), axis=0))
        exp = np.array([O.oracle_verify(k, tp[ci], opk) for ci in range(len(cases))])
        bad = [(cases[ci][0], bool(got[ci]), bool(exp[ci])) for ci in range(len(cases)) if got[ci] != exp[ci]]
        print(f"K={k} tamper matrix: {len(cases)} cases, accepted by oracle: {int(exp.sum())}, mismatches: {bad}", flush=True)
        allok &= not bad
        tpk = opk.copy(); tpk[5] ^= 1
        g2 = ctx.kosk_verify(bytes(opi), bytes(tpk)); e2 = O.oracle_verify(k, opi, tpk)
        print(f"K={k} tampered pk: gpu={g2} oracle={e2}", flush=True)
        allok &= (g2 == e2)
        ctx.close()
    print("ALL OK" if allok else "MISMATCH", flush=True)
    return 0 if allok else 1

if __name__ == "__main__":
    sys.exit(main())
