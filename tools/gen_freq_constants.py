#!/usr/bin/env python3
"""Frequency-domain form of the Poseidon MDS layer for the fp64 pipe: tables + exactness proof.

The MDS matrix of plonky2's Poseidon is M = C + 8 e0 e0^T with C circulant, so y = C x is a cyclic convolution of
length 12.  With  t^12 - 1 = (t^3 - 1)(t^3 + 1)(t^6 + 1)  it splits (CRT, only additions) into a cyclic 3 x 3, a
negacyclic 3 x 3 and a negacyclic 6 x 6 product: 18 + 54 + 18 = 90 fp64 operations per 32-bit half of the state instead
of the 144 multiply-accumulates of the matrix form.  The divisions by 2 of the inverse transform are folded into the
kernel's frequency coefficients (powers of two: exact in binary floating point), the round constants enter as the
initial values of the products (their forward transform, precomputed here), and 2^52 is added last so that the mantissa
of every output IS the integer (poseidon.cuh combine_magic_c).

Paired partial rounds (gen_constants.py::derive_paired):  z = A s' + M[:,0] x + K.  Here
    z = C^2 s' + C[:,0] (8 s'_0 + x - y0) + 8 x e0 + (C c1 + c2),        y0 = M[0,:] s' + c1_0,  x = sbox(y0)
(derivation in DESIGN.md 4.4): C^2 is circulant too, so the 288 multiply-accumulates of A s' become 2 x 90 operations
and the rank-one terms cost 2 x 16.

Everything is verified here with exact rational arithmetic against the matrix forms, and every intermediate value of
the device code (poseidon_freq.cuh, same operation order) is bounded symbolically: |value| / granularity < 2^53, i.e.
exactly representable in a double, for ALL inputs (32-bit halves).  Emits
    plonky2_merkle_trees_b200/csrc/poseidon_freq_constants.cuh
"""
import math
import os
import random
import sys
from fractions import Fraction as Fr

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_constants as G  # noqa: E402

W = G.WIDTH
MAGIC = 2 ** 52
HALF_MAX = 2 ** 32 - 1


# ---- the transform (any ring: ints, Fractions, Lin) -------------------------------------------------------------------
def fwd(x):
    u = [x[j] + x[j + 6] for j in range(6)]
    v = [x[j] - x[j + 6] for j in range(6)]
    return [u[j] + u[j + 3] for j in range(3)] + [u[j] - u[j + 3] for j in range(3)] + v


def inv(o):
    uu = [o[j] + o[3 + j] for j in range(3)] + [o[j] - o[3 + j] for j in range(3)]
    return [uu[j] + o[6 + j] for j in range(6)] + [uu[j] - o[6 + j] for j in range(6)]


def product_coefficients(kernel):
    """12 coefficients [ka 3 | kb 3 | kv 6]: the kernel's residues with the inverse transform's 1/4, 1/4, 1/2 folded in.
    The product matrices are Toeplitz in them (cyclic / negacyclic), the device code applies the signs as operand negations."""
    f = fwd([Fr(k) for k in kernel])
    return [x / 4 for x in f[0:6]] + [x / 2 for x in f[6:12]]


def freq_mul(f, pc, fma):
    """the three residue products; `fma(a, b, c)` = a * b + c in the caller's ring, in the device code's operation order
    (the first term of each sum is a plain multiplication: fma with c = 0)."""
    ka, kb, kv = pc[0:3], pc[3:6], pc[6:12]
    o = []
    for k in range(3):
        acc = 0
        for i in range(3):
            acc = fma(f[i], ka[(k - i) % 3], acc)
        o.append(acc)
    for k in range(3):
        acc = 0
        for i in range(3):
            acc = fma(f[3 + i], kb[k - i] if i <= k else -kb[k - i + 3], acc)
        o.append(acc)
    for k in range(6):
        acc = 0
        for i in range(6):
            acc = fma(f[6 + i], kv[k - i] if i <= k else -kv[k - i + 6], acc)
        o.append(acc)
    return o


def cyclic_conv(x, k):
    return [sum(x[j] * k[(r - j) % W] for j in range(W)) for r in range(W)]


# ---- symbolic bound: value = sum coef_i * input_i + const, inputs in [0, 2^32 - 1] -----------------------------------
class Lin:
    nodes = []   # every value the device code produces

    def __init__(self, coefs, const=Fr(0), record=True):
        self.c = dict((k, Fr(v)) for k, v in coefs.items() if v != 0)
        self.k = Fr(const)
        if record:
            Lin.nodes.append(self)

    @staticmethod
    def _lift(o):
        return o if isinstance(o, Lin) else Lin({}, o, record=False)

    def __add__(self, o):
        o = Lin._lift(o)
        keys = set(self.c) | set(o.c)
        return Lin({k: self.c.get(k, 0) + o.c.get(k, 0) for k in keys}, self.k + o.k)

    __radd__ = __add__

    def __sub__(self, o):
        o = Lin._lift(o)
        keys = set(self.c) | set(o.c)
        return Lin({k: self.c.get(k, 0) - o.c.get(k, 0) for k in keys}, self.k - o.k)

    def scaled(self, s):
        return Lin({k: v * s for k, v in self.c.items()}, self.k * s, record=False)

    def bound(self):
        lo = self.k + sum(v for v in self.c.values() if v < 0) * HALF_MAX
        hi = self.k + sum(v for v in self.c.values() if v > 0) * HALF_MAX
        return lo, hi

    def granularity(self):
        g = None
        for v in list(self.c.values()) + [self.k]:
            if v == 0:
                continue
            q = Fr(1)
            while (v / q).denominator != 1:
                q /= 2
            while (v / (q * 2)).denominator == 1:
                q *= 2
            g = q if g is None or q < g else g
        return g or Fr(1)

    def representable(self):
        lo, hi = self.bound()
        return max(abs(lo), abs(hi)) / self.granularity() < 2 ** 53


def lin_fma(a, b, c):   # a: Lin, b: constant, c: Lin or constant
    return a.scaled(Fr(b)) + c


# ---- the two layers in the device code's operation order, generic in the ring -------------------------------------------
def full_layer_half(x, pm, km, fma):
    """km: 2^52 + this half of the constants the layer adds"""
    x0 = x[0]
    y = inv(freq_mul(fwd(x), pm, fma))
    return [y[0] + fma(x0, 8, km[0])] + [y[j] + km[j] for j in range(1, W)]


def pair_half(x, xs, pm2, kpm, row0, col0, c1_half_magic, fma, magic=MAGIC):
    """x: 12 halves of s', xs: the half of x = sbox(y0).  Returns (L0 = 2^52 + this half of y0, outputs 2^52 + z_half)."""
    l0m = c1_half_magic
    for i in range(W):
        l0m = fma(x[i], row0[i], l0m)
    x0 = x[0]
    y = inv(freq_mul(fwd(x), pm2, fma))
    t = fma(x0, 8, xs)
    d = t - (l0m - magic)
    tt = [fma(d, col0[j], kpm[j]) for j in range(W)]     # kpm: 2^52 + this half of C c1 + c2
    tt[0] = fma(xs, 8, tt[0])
    return l0m, [y[j] + tt[j] for j in range(W)]


def main():
    rc = G.generate_round_constants()
    m = G.mds_matrix()
    circ = [[m[r][c] - (8 if r == c == 0 else 0) for c in range(W)] for r in range(W)]
    kern1 = [G.MDS_CIRC[(-k) % W] for k in range(W)]       # y = C x  <=>  y = x (*) kern1  (cyclic convolution)
    kern2 = cyclic_conv(kern1, kern1)                      # C^2
    rnd = random.Random(99)
    for _ in range(20):
        x = [rnd.randrange(2 ** 32) for _ in range(W)]
        assert cyclic_conv(x, kern1) == [sum(circ[r][c] * x[c] for c in range(W)) for r in range(W)]
        c2x = [sum(circ[r][c] * v for c, v in enumerate([sum(circ[q][c] * x[c] for c in range(W)) for q in range(W)])) for r in range(W)]
        assert cyclic_conv(x, kern2) == c2x
    pm1, pm2 = product_coefficients(kern1), product_coefficients(kern2)
    ffma = lambda a, b, c: a * b + c
    zero = [Fr(0)] * W
    for _ in range(20):   # the transform computes the convolution, exactly
        x = [Fr(rnd.randrange(2 ** 32)) for _ in range(W)]
        assert inv(freq_mul(fwd(x), pm1, ffma)) == cyclic_conv(x, kern1)
        assert inv(freq_mul(fwd(x), pm2, ffma)) == cyclic_conv(x, kern2)

    halves = (lambda v: v & 0xFFFFFFFF, lambda v: v >> 32)
    # ---- full rounds: layer j = 0..7 ends full round r (0..3, 26..29) and adds the constants of round r + 1 ----------
    full_rounds = list(range(G.N_FULL_HALF)) + list(range(G.N_FULL_HALF + G.N_PARTIAL, G.N_ROUNDS))
    kfull = []
    for r in full_rounds:
        nxt = rc[W * (r + 1): W * (r + 2)] if r + 1 < G.N_ROUNDS else [0] * W
        for h in halves:
            init = [h(c) + MAGIC for c in nxt]
            kfull.append(init)
            x = [Fr(rnd.randrange(2 ** 32)) for _ in range(W)]
            want = [sum(m[q][c] * x[c] for c in range(W)) + h(nxt[q]) + MAGIC for q in range(W)]
            assert full_layer_half(x, pm1, init, ffma) == want
    # ---- pairs ------------------------------------------------------------------------------------------------------
    pp = G.derive_paired(rc)
    row0 = [m[0][i] for i in range(W)]
    col0 = [m[j][0] - (8 if j == 0 else 0) for j in range(W)]      # C[:,0]
    kpair = []
    for pair in range(G.N_PARTIAL // 2):
        r = G.N_FULL_HALF + 2 * pair
        c1 = rc[W * (r + 1): W * (r + 2)]
        c2 = rc[W * (r + 2): W * (r + 3)]
        for h in halves:
            kp = [sum(circ[j][i] * h(c1[i]) for i in range(W)) + h(c2[j]) for j in range(W)]     # (C c1 + c2), this half
            init = [v + MAGIC for v in kp]
            kpair.append(init)
            x = [Fr(rnd.randrange(2 ** 32)) for _ in range(W)]
            xs = Fr(rnd.randrange(2 ** 32))
            l0m, z = pair_half(x, xs, pm2, init, row0, col0, h(c1[0]) + MAGIC, ffma)
            assert l0m == sum(row0[i] * x[i] for i in range(W)) + h(c1[0]) + MAGIC
            # the existing paired form, half by half (all maps are integer-linear, so halves can be treated separately;
            # K there is reduced mod p as a whole word, so compare against the un-reduced integer form)
            k_int = [sum(m[j][i] * h(c1[i]) for i in range(1, W)) + h(c2[j]) for j in range(W)]
            want = [sum(pp["a"][j][i] * x[i] for i in range(W)) + m[j][0] * xs + k_int[j] + MAGIC for j in range(W)]
            assert z == want, (pair, z[0] - want[0])

    # ---- exactness of every intermediate value, symbolically, worst case over all inputs ---------------------------------
    worst = Fr(0)
    def check_nodes(what):
        nonlocal worst
        for nd in Lin.nodes:
            lo, hi = nd.bound()
            ratio = max(abs(lo), abs(hi)) / nd.granularity()
            worst = max(worst, ratio)
            assert ratio < 2 ** 53, (what, float(ratio))
        Lin.nodes = []
    for j in range(len(full_rounds)):
        for hh in range(2):
            x = [Lin({i: 1}) for i in range(W)]
            out = full_layer_half(x, pm1, kfull[2 * j + hh], lin_fma)
            for o in out:   # the outputs must leave the integer in the mantissa: 2^52 <= value < 2^53
                lo, hi = o.bound()
                assert lo >= MAGIC and hi < 2 * MAGIC and o.granularity() >= 1
            check_nodes("full")
    for pair in range(G.N_PARTIAL // 2):
        r = G.N_FULL_HALF + 2 * pair
        for hh, h in enumerate(halves):
            x = [Lin({i: 1}) for i in range(W)]
            xs = Lin({"x": 1})
            l0m, z = pair_half(x, xs, pm2, kpair[2 * pair + hh], row0, col0, h(rc[W * (r + 1)]) + MAGIC, lin_fma)
            lo, hi = l0m.bound()
            assert lo >= MAGIC and hi < 2 * MAGIC
            for o in z:
                lo, hi = o.bound()
                # the symbolic lower bound ignores that y0 and s' are correlated (d = 8 s'_0 + x - y0 enters with a
                # minus sign); the true value is A s' + M[:,0] x + K >= 0 (asserted above against the matrix form), and
                # the upper bound holds
                assert hi < 2 * MAGIC and o.granularity() >= 1 and lo > MAGIC // 2
            check_nodes("pair")
    print("exactness: worst |value| / granularity = 2^%.2f (< 2^53)" % math.log2(float(worst)))

    # ---- emit ---------------------------------------------------------------------------------------------------------------
    def fmt(v):
        v = Fr(v)
        assert v.denominator in (1, 2, 4)
        return "%s" % (repr(float(v)))
    out = []
    out.append("// GENERATED by tools/gen_freq_constants.py -- do not edit.")
    out.append("// Frequency-domain form of the Poseidon MDS layer (poseidon_freq.cuh): residues [ka 3 | kb 3 | kv 6] (pre-scaled) of")
    out.append("// the circulant part C of the MDS matrix and of C^2, and the round constants' forward transforms.  Exact binary")
    out.append("// fractions (multiples of 1/4); exactness of every intermediate proven by the generator.")
    out.append("#ifndef PMT_POSEIDON_FREQ_CONSTANTS_CUH\n#define PMT_POSEIDON_FREQ_CONSTANTS_CUH")
    out.append("#ifndef PMT_FQ_QUAL\n#define PMT_FQ_QUAL static __device__ __constant__ double\n#endif")
    out.append("PMT_FQ_QUAL PMT_FQ_P1[12] = {%s};" % ", ".join(fmt(v) for v in pm1))
    out.append("PMT_FQ_QUAL PMT_FQ_P2[12] = {%s};" % ", ".join(fmt(v) for v in pm2))
    out.append("// full layers add PMT_RC_DM (poseidon_constants.cuh: 2^52 + lo, 2^52 + hi of the next round's constants)")
    out.append("// [pair 0..10][lane][lo, hi]: 2^52 + the halves of C c1 + c2")
    kp_il = []
    for pair in range(G.N_PARTIAL // 2):
        for j in range(W):
            kp_il += [kpair[2 * pair][j], kpair[2 * pair + 1][j]]
    out.append("PMT_FQ_QUAL PMT_FQ_KPAIR_DM[%d] = {%s};" % (len(kp_il), ", ".join(fmt(v) for v in kp_il)))
    out.append("// MDS_MATRIX_CIRC and, at index 12, CIRC[0] + DIAG[0]: M[0,i] = CIRC13[i == 0 ? 12 : i], C[j,0] = CIRC13[(12 - j) % 12]")
    assert row0 == [G.MDS_CIRC[0] + 8] + G.MDS_CIRC[1:] and col0 == [G.MDS_CIRC[(12 - j) % 12] for j in range(W)]
    out.append("PMT_FQ_QUAL PMT_FQ_CIRC13[13] = {%s};" % ", ".join(fmt(v) for v in G.MDS_CIRC + [G.MDS_CIRC[0] + 8]))
    out.append("#endif")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "plonky2_merkle_trees_b200", "csrc",
                        "poseidon_freq_constants.cuh")
    text = "\n".join(out) + "\n"
    # an unchanged table is not rewritten: the header is a build dependency of libpmt.so, and touching it (the CPU test suite
    # runs this script) would make every later import rebuild the library
    if os.path.exists(path) and open(path).read() == text:
        print("ok: frequency-domain tables verified against the matrix forms; %s is up to date" % os.path.relpath(path))
    else:
        with open(path, "w") as f:
            f.write(text)
        print("ok: frequency-domain tables verified against the matrix forms; %s written" % os.path.relpath(path))


if __name__ == "__main__":
    main()
