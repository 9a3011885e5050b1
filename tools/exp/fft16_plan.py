"""Which of the 17 twiddle multiplications of a radix-2 decimation-in-time 16-point DFT over GF(3329) may stay LAZY (one IMAD, unreduced
int32 product with a centered constant |w| <= 1664) and which must be SHOUP multiplications (IMAD.HI + 2 IMAD = 4 FMA-heavy issue slots,
result below 1.25 q for any |s| < 2^31)?  Exhaustive search over the 2^17 masks with interval bounds: every intermediate must stay below
2^31.  Prints the masks `share_ntt.cuh` hard-codes (SN_FFT_SMALL, SN_FFT_BIG).  No GPU needed: python tools/exp/fft16_plan.py

Multiplication ids (as in SnDft): size-4 transform at offset o: o; size-8 transform at offset o, k = 1..3: 4 + 3 o + k - 1; size 16,
k = 1..7: 9 + k."""
import math

Q = 3329
LIM = 2 ** 31 - 1


def shoup_bound(s):
    return Q * (1 + s / 2 ** 33)


def plan(nz_inputs, in_bound, need_outputs, out_limit, mask):
    slots, ok = 0, True

    def dft(n, off, stride):
        nonlocal slots, ok
        if n == 1:
            return [in_bound if off in nz_inputs else 0]
        ev, od = dft(n // 2, off, 2 * stride), dft(n // 2, off + stride, 2 * stride)
        out = [0] * n
        for k in range(n // 2):
            t = od[k]
            if k and t:
                mid = off if n == 4 else 4 + 3 * off + k - 1 if n == 8 else 9 + k
                if mask >> mid & 1:
                    ok &= t <= LIM
                    t, slots = shoup_bound(t), slots + 4
                else:
                    t, slots = t * 1664, slots + 1
            out[k] = out[k + n // 2] = ev[k] + t
            ok &= out[k] <= LIM
        return out
    out = dft(16, 0, 1)
    mx = max(out[k] for k in need_outputs)
    return slots, ok and mx <= out_limit, mx


def best(name, nz_inputs, in_bound, need_outputs, out_limit=LIM):
    res = None
    for mask in range(1 << 17):
        s, ok, mx = plan(nz_inputs, in_bound, need_outputs, out_limit, mask)
        if ok and (res is None or (s, mx) < (res[0], res[2])):
            res = (s, mask, mx)
    print(f"{name}: {res[0]} issue slots, mask {res[1]:#x}, outputs below 2^{math.log2(res[2]):.2f}")


if __name__ == "__main__":
    every = set(range(16))
    best("forward stage 1 (8 non-zero inputs below 1.04 q)", set(range(8)), 1.04 * Q, range(16))
    best("forward stage 2 (inputs below 1.25 q, outputs go to Barrett)", every, 1.25 * Q, range(16))
    best("inverse stage 1 (unreduced pointwise sums, 4 x 6700 x 1664)", every, 4 * 6700 * 1664, range(16))
    best("inverse stage 2 (inputs below 1.25 q, outputs 0..8)", every, 1.25 * Q, range(9))
