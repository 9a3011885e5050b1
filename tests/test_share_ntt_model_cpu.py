"""CPU suite: the identity behind csrc/share_ntt.cuh, checked against the oracle's table mat-vec (ko_share_ddeg, reference ss.cpp:76-99)
in plain Python integers -- no GPU, no Montgomery form: a packed-Shamir sharing is P(x) * sum_j c[x + 407 - j] * (w_j y_j) with c[m] = 1/m,
and that Toeplitz product equals the blocked length-256 cyclic convolutions over GF(3329) (4 input blocks, 11 output blocks, root 17)."""
import numpy as np

import oracle_lib as O

Q, D1, NX = 3329, 407, 1303


def _inv(a):
    return pow(int(a) % Q, Q - 2, Q)


def _ntt(v, root):
    n = len(v)
    pw = [pow(root, i, Q) for i in range(n)]
    return [sum(v[t] * pw[(t * k) % n] for t in range(n)) % Q for k in range(n)]


def test_blocked_ntt_convolution_equals_the_share_table():
    assert pow(17, 256, Q) == 1 and pow(17, 128, Q) == Q - 1          # 17 generates the 256-th roots of unity (the Kyber NTT's root)
    w, P = [], []
    for j in range(D1):
        d = 1
        for m in range(D1):
            if m != j:
                d = d * ((j - m) % Q) % Q
        w.append(_inv(d))
    for x in range(NX):
        p = 1
        for m in range(D1):
            p = p * ((x + D1 - m) % Q) % Q
        P.append(p)
    c = [0] + [_inv(m) for m in range(1, Q)]
    rng = np.random.default_rng(3)
    y = rng.integers(0, Q, size=D1, dtype=np.uint16)
    want = O.oracle_share(y)                                            # parties 0..150 = tail, 151.. = S y
    assert (want[:151] == y[256:]).all()
    u = [w[j] * int(y[j]) % Q for j in range(D1)] + [0] * (512 - D1)
    U = [_ntt(u[128 * i:128 * i + 128] + [0] * 128, 17) for i in range(4)]
    spectra = {}
    for dl in range(-3, 11):                                            # kernel segment K_dl[d] = c[128 dl + 407 + d], d in [-127, 127]
        K = [0] * 256
        for t in range(256):
            d = t if t < 128 else t - 256
            m = 128 * dl + D1 + d
            K[t] = c[m] if (t != 128 and 1 <= m < Q) else 0
        spectra[dl] = _ntt(K, 17)
    i256, iom = _inv(256), _inv(17)
    got = np.zeros(NX, np.int64)
    for o in (0, 5, 10):                                                # first, a middle and the last (partial) output block
        acc = [sum(spectra[o - i][n] * U[i][n] for i in range(4)) % Q for n in range(256)]
        blk = _ntt(acc, iom)
        for t in range(128):
            x = 128 * o + t
            if x < NX:
                got[x] = blk[t] * i256 % Q * P[x] % Q
                assert got[x] == want[151 + x], (o, t)


def test_unequal_blocks_and_limb_split_equal_the_share_table():
    """k_share_ntt2's form of the same identity: 126-wide input blocks x 131-wide output blocks (126 + 131 - 1 = 256, so ten output blocks
    cover the 1303 shares), segment (o, i) K[d] = c[131 o - 126 i + 407 + d] for d in [-125, 130], and the pointwise stage on signed limbs
    k = 64 k1 + k0 of the centered segment spectra (the IDP.2A operands): sum_i u_i (k0_i + 64 k1_i) = the product over GF(3329)."""
    BI, BO, NIN, NOUT = 126, 131, 4, 10
    assert BI + BO - 1 == 256 and NIN * BI >= D1 and NOUT * BO >= NX
    w, P = [], []
    for j in range(D1):
        d = 1
        for m in range(D1):
            if m != j:
                d = d * ((j - m) % Q) % Q
        w.append(_inv(d))
    for x in range(NX):
        p = 1
        for m in range(D1):
            p = p * ((x + D1 - m) % Q) % Q
        P.append(p)
    rng = np.random.default_rng(5)
    y = rng.integers(0, Q, size=D1, dtype=np.uint16)
    want = O.oracle_share(y)
    u = [w[j] * int(y[j]) % Q for j in range(D1)] + [0] * (NIN * BI - D1)
    U = [_ntt(u[BI * i:BI * (i + 1)] + [0] * (256 - BI), 17) for i in range(NIN)]
    i256, iom = _inv(256), _inv(17)
    for o in (0, 4, 9):
        acc0, acc1 = [0] * 256, [0] * 256
        for i in range(NIN):
            K = [0] * 256
            for t in range(256):
                d = t if t < BO else t - 256
                mm = (BO * o - BI * i + D1 + d) % Q
                K[t] = _inv(mm) if mm else 0
            spec = [v * i256 % Q for v in _ntt(K, 17)]
            for n in range(256):
                v = spec[n] - Q if spec[n] > Q // 2 else spec[n]      # centered, then limbs as in share_ntt_tables()
                k0 = ((v + 32) & 63) - 32
                k1 = (v - k0) // 64
                assert -32 <= k0 < 32 and -27 <= k1 <= 26 and 64 * k1 + k0 == v
                acc0[n] += U[i][n] * k0
                acc1[n] += U[i][n] * k1
        blk = _ntt([(a0 + 64 * a1) % Q for a0, a1 in zip(acc0, acc1)], iom)
        for t in range(BO):
            x = BO * o + t
            if x < NX:
                assert blk[t] * P[x] % Q == want[151 + x], (o, t)


def test_dft_network_masks_keep_every_intermediate_below_2_31():
    """The lazy / Shoup assignment hard-coded in share_ntt.cuh (SN_FFT_SMALL, SN_FFT_BIG) against the interval model of tools/exp/fft16_plan.py,
    for every place the kernels use it and the largest inputs that can reach it."""
    import importlib.util
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("fft16_plan", os.path.join(root, "tools", "exp", "fft16_plan.py"))
    plan = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(plan)
    src = open(os.path.join(root, "mpcith_kyber_kosk_b200", "csrc", "share_ntt.cuh")).read()
    m = re.search(r"SN_FFT_SMALL = (0x[0-9a-f]+)u, SN_FFT_BIG = (0x[0-9a-f]+)u", src)
    small, big = int(m.group(1), 16), int(m.group(2), 16)
    every = set(range(16))
    cases = [
        ("forward stage 1, factor applied (Shoup output)", set(range(8)), 1.25 * Q, range(16), small),
        ("forward stage 1, raw u16 rows of the verifier's products", set(range(8)), 65535, range(16), small),
        ("forward stage 2 / inverse stage 2 (twiddle Shoup output)", every, 1.25 * Q, range(16), small),
        ("inverse stage 1, k_share_ntt2 (4 blocks)", every, 4 * 6700 * 1664, range(16), big),
        ("inverse stage 1, k_conv_ntt<8, 2> (8 blocks)", every, 8 * 6700 * 1664, range(16), big),
    ]
    for name, nz, bound, outs, mask in cases:
        slots, ok, mx = plan.plan(nz, bound, outs, plan.LIM, mask)
        assert ok and mx < 2 ** 31, name
    # the cheapest valid assignment for reduced inputs costs 32 issue slots, and the hard-coded one is such an assignment
    assert plan.plan(every, 1.25 * Q, range(16), plan.LIM, small)[0] == 32
    assert plan.plan(every, 4 * 6700 * 1664, range(16), plan.LIM, big)[0] == 68
