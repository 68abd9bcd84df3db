#!/usr/bin/env python3
"""Codelet generator for the pruned 64-point transforms of the FNO spectral layer.

The spectral layer of the reference (src/models/fno/fno2d.py:59-82) keeps 24x12 of the 64x33
rfft2 modes, so every 1-D transform on the hot path is *pruned*: few outputs of many inputs
(forward) or few non-zero inputs (inverse).  Instead of hand-writing butterflies, this script
builds each transform symbolically -- a hash-consed expression DAG over real scalars with
algebraic simplification (x+0, 1*x, -(-x), constant folding of twiddles from float64 tables) and
*demand-driven* evaluation, so unused outputs are never computed and zero inputs vanish -- and
emits straight-line C++ (`__host__ __device__`, templated on the scalar type) into
`fft_codelets.cuh`.  All array indices are compile-time constants, so in a kernel every operand
lives in a register and every twiddle is an FFMA immediate.

Codelets (N = 64 everywhere):
  rfft64_lo13        64 real in            -> bins 0..12 (complex)           forward  e^{-i..}
  cfft64_r<j>        64 complex in         -> bins {0..11, 53..63} = j mod 4  forward
  icfft64_in24_r<r>  24 complex in (bins 0..11, 52..63) -> outputs h = 8h'+r, h'=0..7   inverse
  icfft64_in24_full  24 complex in (bins 0..11, 52..63) -> all 64 outputs               inverse
  c2r64_in12         12 complex in (bins 0..11; Im of bin 0 ignored) -> 64 real out   inverse
                     y[w] = Re sum_k Z[k] e^{+2 pi i k w/64}   (caller pre-scales Z by c_ky/HW)

Run `python gen_codelets.py` to regenerate the header; `python gen_codelets.py --selftest`
evaluates every DAG numerically against numpy's FFT.
"""
from __future__ import annotations

import argparse
import math
import os
import sys

import numpy as np

# ------------------------------------------------------------------------------------------ DAG


class Node:
    __slots__ = ("op", "a", "b", "c", "id", "name")

    def __init__(self, op, a=None, b=None, c=None, name=None):
        self.op, self.a, self.b, self.c, self.name = op, a, b, c, name
        self.id = None


class Graph:
    """Hash-consed real-valued expression DAG."""

    def __init__(self):
        self.table = {}
        self.nodes = []

    def _mk(self, op, a=None, b=None, c=None, name=None):
        key = (op, a.id if isinstance(a, Node) else a, b.id if isinstance(b, Node) else b, c, name)
        n = self.table.get(key)
        if n is None:
            n = Node(op, a, b, c, name)
            n.id = len(self.nodes)
            self.nodes.append(n)
            self.table[key] = n
        return n

    def inp(self, name):
        return self._mk("in", name=name)

    # None represents an exact zero
    def neg(self, x):
        if x is None:
            return None
        if x.op == "neg":
            return x.a
        if x.op == "mul":
            return self._mk("mul", x.a, None, -x.c)
        if x.op == "sub":
            return self._mk("sub", x.b, x.a)
        return self._mk("neg", x)

    def add(self, x, y):
        if x is None:
            return y
        if y is None:
            return x
        if y.op == "neg":
            return self.sub(x, y.a)
        if x.op == "neg":
            return self.sub(y, x.a)
        if x.id > y.id:
            x, y = y, x
        return self._mk("add", x, y)

    def sub(self, x, y):
        if y is None:
            return x
        if x is None:
            return self.neg(y)
        if y.op == "neg":
            return self.add(x, y.a)
        if x is y:
            return None
        return self._mk("sub", x, y)

    def mul(self, c, x):
        if x is None or c == 0.0:
            return None
        if c == 1.0:
            return x
        if c == -1.0:
            return self.neg(x)
        if x.op == "neg":
            return self.mul(-c, x.a)
        if x.op == "mul":
            return self.mul(c * x.c, x.a)
        return self._mk("mul", x, None, float(c))


def _snap(v: float) -> float:
    """Snap twiddle components that are exactly 0/+-1 in exact arithmetic."""
    for t in (0.0, 1.0, -1.0):
        if abs(v - t) < 1e-15:
            return t
    return v


def twiddle(k: int, n: int, sign: int):
    """e^{sign * 2 pi i k / n} from float64."""
    k %= n
    ang = 2.0 * math.pi * k / n
    return _snap(math.cos(ang)), _snap(sign * math.sin(ang))


class C:
    """Complex value made of two DAG nodes."""
    __slots__ = ("re", "im")

    def __init__(self, re, im):
        self.re, self.im = re, im


def cadd(g, x, y):
    return C(g.add(x.re, y.re), g.add(x.im, y.im))


def csub(g, x, y):
    return C(g.sub(x.re, y.re), g.sub(x.im, y.im))


def cconj(g, x):
    return C(x.re, g.neg(x.im))


def cmulc(g, x, wr, wi):
    """x * (wr + i wi) with constant folding."""
    re = g.sub(g.mul(wr, x.re), g.mul(wi, x.im))
    im = g.add(g.mul(wi, x.re), g.mul(wr, x.im))
    return C(re, im)


# ------------------------------------------------------------------------- lazy transforms

def lazy_cfft(g, xs, sign):
    """Demand-driven radix-2 DIT complex DFT of the list xs (entries may be C with None parts).
    Returns a function bin(k) -> C.  sign=-1 forward, +1 inverse (unnormalised)."""
    n = len(xs)
    memo = {}
    if n == 1:
        return lambda k: xs[0]
    if n == 2:
        def bin2(k):
            k %= 2
            if k not in memo:
                memo[k] = cadd(g, xs[0], xs[1]) if k == 0 else csub(g, xs[0], xs[1])
            return memo[k]
        return bin2
    ev = lazy_cfft(g, xs[0::2], sign)
    od = lazy_cfft(g, xs[1::2], sign)
    h = n // 2

    def binn(k):
        k %= n
        if k not in memo:
            wr, wi = twiddle(k % h, n, sign)
            t = cmulc(g, od(k % h), wr, wi)
            memo[k] = cadd(g, ev(k % h), t) if k < h else csub(g, ev(k % h), t)
        return memo[k]
    return binn


def lazy_rfft(g, xs, sign=-1):
    """Demand-driven DFT of a REAL list xs (nodes), exploiting Hermitian symmetry at every level:
    bins above n/2 are conjugates of their mirrors, so only bins 0..n/2 are ever built."""
    n = len(xs)
    memo = {}
    if n == 1:
        return lambda k: C(xs[0], None)
    if n == 2:
        def bin2(k):
            k %= 2
            if k not in memo:
                memo[k] = C(g.add(xs[0], xs[1]), None) if k == 0 else C(g.sub(xs[0], xs[1]), None)
            return memo[k]
        return bin2
    ev = lazy_rfft(g, xs[0::2], sign)
    od = lazy_rfft(g, xs[1::2], sign)
    h = n // 2

    def binn(k):
        k %= n
        if k > h:
            return cconj(g, binn(n - k))
        if k not in memo:
            wr, wi = twiddle(k, n, sign)
            t = cmulc(g, od(k % h), wr, wi)
            memo[k] = cadd(g, ev(k % h), t)
        return memo[k]
    return binn


# ------------------------------------------------------------------------------ codelets

N = 64
FWD_BINS = list(range(0, 12)) + list(range(53, 64))        # bins -11..11 of the w-transform
KEPT_KX = list(range(0, 12)) + list(range(52, 64))         # kept rows of the h-transform
INV_R = 8                                                   # K3 splits rows h = 8h'+r


def build_rfft64_lo13():
    g = Graph()
    xs = [g.inp(f"x[{i}]") for i in range(N)]
    f = lazy_rfft(g, xs, -1)
    outs = []
    for k in range(13):
        v = f(k)
        outs.append((f"ore[{k}]", v.re))
        outs.append((f"oim[{k}]", v.im))
    return g, outs


def cfft64_bins(j):
    return [k for k in FWD_BINS if k % 4 == j]


def build_cfft64_r(j):
    """DIF split by output residue: X[4k'+j] = sum_{n<16} W16^{n k'} W64^{n j} sum_m x[n+16m] (-i)^{m j}."""
    g = Graph()
    xs = [C(g.inp(f"xre[{i}]"), g.inp(f"xim[{i}]")) for i in range(N)]
    folded = []
    for n in range(16):
        acc = None
        for m in range(4):
            wr, wi = twiddle(m * j, 4, -1)
            t = cmulc(g, xs[n + 16 * m], wr, wi)
            acc = t if acc is None else cadd(g, acc, t)
        wr, wi = twiddle(n * j, 64, -1)
        folded.append(cmulc(g, acc, wr, wi))
    f = lazy_cfft(g, folded, -1)
    outs = []
    for idx, k in enumerate(cfft64_bins(j)):
        v = f((k - j) // 4)
        outs.append((f"ore[{idx}]", v.re))
        outs.append((f"oim[{idx}]", v.im))
    return g, outs


def build_icfft64_in24_r(r):
    """Inverse along kx with 24 non-zero inputs, outputs h = 8h'+r:
    Z[8h'+r] = sum_{m<8} W8^{-m h'} sum_{kx = m mod 8} Y[kx] W64^{-kx r}  (W = e^{-2 pi i/.})."""
    g = Graph()
    ys = {kx: C(g.inp(f"yre[{i}]"), g.inp(f"yim[{i}]")) for i, kx in enumerate(KEPT_KX)}
    folded = []
    for m in range(INV_R):
        acc = None
        for kx in KEPT_KX:
            if kx % INV_R != m:
                continue
            wr, wi = twiddle(kx * r, 64, +1)
            t = cmulc(g, ys[kx], wr, wi)
            acc = t if acc is None else cadd(g, acc, t)
        folded.append(acc if acc is not None else C(None, None))
    f = lazy_cfft(g, folded, +1)
    outs = []
    for hp in range(N // INV_R):
        v = f(hp)
        outs.append((f"ore[{hp}]", v.re))
        outs.append((f"oim[{hp}]", v.im))
    return g, outs


def build_icfft64_in24_full():
    """Inverse along kx with 24 non-zero inputs (bins 0..11, 52..63), all 64 outputs."""
    g = Graph()
    zero = C(None, None)
    full = [zero] * N
    for i, kx in enumerate(KEPT_KX):
        full[kx] = C(g.inp(f"yre[{i}]"), g.inp(f"yim[{i}]"))
    f = lazy_cfft(g, full, +1)
    outs = []
    for h in range(N):
        v = f(h)
        outs.append((f"ore[{h}]", v.re))
        outs.append((f"oim[{h}]", v.im))
    return g, outs


def build_c2r64_in12():
    """y[w] = Re sum_{k<12} Z[k] e^{+2 pi i k w/64}, Im Z[0] ignored (C2R semantics of
    torch.fft.irfft2's last axis, reference fno2d.py:81).  Packed-pair algorithm: with
    Zh[k] = Z[k]/1 for the 'half spectrum' convention y = Re(sum) we build
    V[k] = A[k] + i e^{+2 pi i k/64} B[k] over k<32 with A = Z[k] + conj(Z[32-k])... specialised to
    Z[k]=0 for k>=12, then a 32-point inverse DFT yields v[n] = y[2n] + i y[2n+1]."""
    g = Graph()
    z = [C(g.inp(f"zre[{k}]"), g.inp(f"zim[{k}]") if k > 0 else None) for k in range(12)]
    zero = C(None, None)

    def Z(k):  # spectrum entry (zero beyond the kept 12)
        return z[k] if 0 <= k < 12 else zero

    # y[w] = Re sum_k Z[k] e^{i th k w}.  Let E[n] = y[2n], O[n] = y[2n+1].
    # y[w] = 1/2 sum_k (Z[k] e^{+} + conj(Z[k]) e^{-}) = sum over a Hermitian 64-spectrum
    # S[k] = Z[k]/2 (0<k<32), S[64-k] = conj(Z[k])/2, S[0] = Re Z[0].
    # Standard C2R packing: v[n] = E[n] + i O[n] = IDFT32( S[k] + S[k+32] + i w^k (S[k] - S[k+32]) ),
    # w = e^{+2 pi i/64}.  With S[k+32] = conj(S[32-k]):
    def S(k):
        k %= 64
        if k == 0:
            return C(z[0].re, None)
        if k < 32:
            zz = Z(k)
            return C(g.mul(0.5, zz.re), g.mul(0.5, zz.im))
        if k == 32:
            return zero
        zz = Z(64 - k)
        return C(g.mul(0.5, zz.re), g.neg(g.mul(0.5, zz.im)))

    vs = []
    for k in range(32):
        a = cadd(g, S(k), S(k + 32))
        b = csub(g, S(k), S(k + 32))
        wr, wi = twiddle(k, 64, +1)
        bw = cmulc(g, b, wr, wi)
        ib = C(g.neg(bw.im), bw.re)  # i * bw
        vs.append(cadd(g, a, ib))
    f = lazy_cfft(g, vs, +1)
    outs = []
    for n in range(32):
        v = f(n)
        outs.append((f"y[{2 * n}]", v.re))
        outs.append((f"y[{2 * n + 1}]", v.im))
    return g, outs


# --------------------------------------------------------------------------- evaluation / emit

def needed(outs):
    seen, order = set(), []

    def visit(n):
        if n is None or n.id in seen:
            return
        # iterative DFS to stay clear of recursion limits
        stack = [(n, False)]
        while stack:
            node, done = stack.pop()
            if node is None:
                continue
            if done:
                order.append(node)
                continue
            if node.id in seen:
                continue
            seen.add(node.id)
            stack.append((node, True))
            for ch in (node.b, node.a):
                if isinstance(ch, Node) and ch.id not in seen:
                    stack.append((ch, False))
    for _, n in outs:
        visit(n)
    return order


def evaluate(outs, inputs: dict):
    vals = {}
    for n in needed(outs):
        if n.op == "in":
            vals[n.id] = inputs[n.name]
        elif n.op == "add":
            vals[n.id] = vals[n.a.id] + vals[n.b.id]
        elif n.op == "sub":
            vals[n.id] = vals[n.a.id] - vals[n.b.id]
        elif n.op == "mul":
            vals[n.id] = n.c * vals[n.a.id]
        elif n.op == "neg":
            vals[n.id] = -vals[n.a.id]
    return {name: (0.0 if n is None else vals[n.id]) for name, n in outs}


def op_counts(outs):
    cnt = {"add": 0, "sub": 0, "mul": 0, "neg": 0, "in": 0}
    for n in needed(outs):
        cnt[n.op] += 1
    return cnt


def fmt_const(c: float) -> str:
    return f"T({c!r})"


def emit(name, signature, outs, doc):
    """Straight-line code.  A constant multiple used exactly once, by an add or a sub, is fused into it:
    `fno_fma(c, a, x)` = c*a + x and `fno_fms(c, a, x)` = c*a - x  (x - c*a becomes fno_fma(-c, a, x)).  For scalar T the
    compiler would contract these anyway; spelling them out lets a packed two-lane type (f32x2 in fno_dft_fwd.cu) map
    every fused pair onto ONE FFMA2 instead of FMUL2 + FADD2."""
    order = needed(outs)
    uses, addsub_uses = {}, {}
    for n in order:
        for opnd in (n.a, n.b):
            if isinstance(opnd, Node):
                uses[opnd.id] = uses.get(opnd.id, 0) + 1
                if n.op in ("add", "sub"):
                    addsub_uses[opnd.id] = addsub_uses.get(opnd.id, 0) + 1
    for _, n in outs:
        if n is not None:
            uses[n.id] = uses.get(n.id, 0) + 1
    # a multiply whose every consumer is an add / sub can be folded into each of them (a +- c*t butterflies: two FMAs
    # instead of one multiply and two adds); a consumer can absorb only one of its operands
    cand = lambda m: isinstance(m, Node) and m.op == "mul" and uses.get(m.id, 0) == addsub_uses.get(m.id, 0)  # noqa: E731
    choice, chosen = {}, {}
    for n in order:
        if n.op in ("add", "sub"):
            pick = n.b if cand(n.b) else (n.a if cand(n.a) else None)
            if pick is not None:
                choice[n.id] = pick.id
                chosen[pick.id] = chosen.get(pick.id, 0) + 1
    fused = {mid for mid, k in chosen.items() if k == uses[mid]}   # muls that need no instruction of their own
    for nid in list(choice):   # consumers of a mul that must be emitted anyway use it directly
        if choice[nid] not in fused:
            del choice[nid]
    cnt = op_counts(outs)
    n_fma = len(choice)
    total = cnt["add"] + cnt["sub"] + cnt["mul"] + cnt["neg"] - len(fused)
    lines = [f"// {doc}", f"// ops: {cnt}; emitted: {total} instructions, {n_fma} of them fma ({len(fused)} multiplies folded)",
             "template <typename T>", f"FNO_HD void {name}({signature}) {{"]
    ref = {}
    for n in order:
        if n.op == "in":
            ref[n.id] = n.name
            continue
        if n.op == "mul" and n.id in fused:
            continue
        v = f"t{n.id}"
        if n.op in ("add", "sub"):
            a, b = n.a, n.b
            if choice.get(n.id) == b.id:
                # x + c*t  /  x - c*t
                c = b.c if n.op == "add" else -b.c
                e = f"fno_fma({fmt_const(c)}, {ref[b.a.id]}, {ref[a.id]})"
            elif choice.get(n.id) == a.id:
                # c*t + y  /  c*t - y
                e = (f"fno_fma({fmt_const(a.c)}, {ref[a.a.id]}, {ref[b.id]})" if n.op == "add"
                     else f"fno_fms({fmt_const(a.c)}, {ref[a.a.id]}, {ref[b.id]})")
            else:
                e = f"{ref[a.id]} {'+' if n.op == 'add' else '-'} {ref[b.id]}"
        elif n.op == "mul":
            e = f"{fmt_const(n.c)} * {ref[n.a.id]}"
        else:
            e = f"-{ref[n.a.id]}"
        lines.append(f"  const T {v} = {e};")
        ref[n.id] = v
    for oname, n in outs:
        lines.append(f"  {oname} = {'T(0)' if n is None else ref[n.id]};")
    lines.append("}")
    return "\n".join(lines)


def all_codelets():
    cl = []
    g, o = build_rfft64_lo13()
    cl.append(("rfft64_lo13", "const T* __restrict__ x, T* __restrict__ ore, T* __restrict__ oim", o,
               "64 real in -> bins 0..12 of the forward DFT"))
    for j in range(4):
        g, o = build_cfft64_r(j)
        cl.append((f"cfft64_r{j}",
                   "const T* __restrict__ xre, const T* __restrict__ xim, T* __restrict__ ore, T* __restrict__ oim", o,
                   f"64 complex in -> forward DFT bins {cfft64_bins(j)} (in this order)"))
    for r in range(INV_R):
        g, o = build_icfft64_in24_r(r)
        cl.append((f"icfft64_in24_r{r}",
                   "const T* __restrict__ yre, const T* __restrict__ yim, T* __restrict__ ore, T* __restrict__ oim", o,
                   f"24 complex in (bins 0..11,52..63) -> inverse DFT outputs h=8h'+{r}, h'=0..7"))
    g, o = build_icfft64_in24_full()
    cl.append(("icfft64_in24_full",
               "const T* __restrict__ yre, const T* __restrict__ yim, T* __restrict__ ore, T* __restrict__ oim", o,
               "24 complex in (bins 0..11,52..63) -> inverse DFT, all 64 outputs"))
    g, o = build_c2r64_in12()
    cl.append(("c2r64_in12", "const T* __restrict__ zre, const T* __restrict__ zim, T* __restrict__ y", o,
               "12 complex in (Im of bin 0 ignored) -> y[w] = Re sum_k Z[k] e^{+2 pi i k w/64}, w=0..63"))
    return cl


def selftest():
    rng = np.random.default_rng(0)
    # rfft
    g, o = build_rfft64_lo13()
    x = rng.standard_normal(64)
    res = evaluate(o, {f"x[{i}]": x[i] for i in range(64)})
    ref = np.fft.fft(x)
    for k in range(13):
        assert abs(res[f"ore[{k}]"] - ref[k].real) < 1e-12 and abs(res[f"oim[{k}]"] - ref[k].imag) < 1e-12
    print("rfft64_lo13 ok", op_counts(o))
    # cfft residues
    z = rng.standard_normal(64) + 1j * rng.standard_normal(64)
    ref = np.fft.fft(z)
    for j in range(4):
        g, o = build_cfft64_r(j)
        inp = {f"xre[{i}]": z[i].real for i in range(64)}
        inp.update({f"xim[{i}]": z[i].imag for i in range(64)})
        res = evaluate(o, inp)
        for idx, k in enumerate(cfft64_bins(j)):
            assert abs(res[f"ore[{idx}]"] - ref[k].real) < 1e-12 and abs(res[f"oim[{idx}]"] - ref[k].imag) < 1e-12
        print(f"cfft64_r{j} ok", op_counts(o))
    # inverse along kx
    y = rng.standard_normal(24) + 1j * rng.standard_normal(24)
    full = np.zeros(64, dtype=complex)
    full[KEPT_KX] = y
    ref = np.fft.ifft(full) * 64
    for r in range(INV_R):
        g, o = build_icfft64_in24_r(r)
        inp = {f"yre[{i}]": y[i].real for i in range(24)}
        inp.update({f"yim[{i}]": y[i].imag for i in range(24)})
        res = evaluate(o, inp)
        for hp in range(8):
            v = ref[8 * hp + r]
            assert abs(res[f"ore[{hp}]"] - v.real) < 1e-12 and abs(res[f"oim[{hp}]"] - v.imag) < 1e-12
        print(f"icfft64_in24_r{r} ok", op_counts(o))
    g, o = build_icfft64_in24_full()
    res = evaluate(o, inp)
    for h in range(64):
        assert abs(res[f"ore[{h}]"] - ref[h].real) < 1e-12 and abs(res[f"oim[{h}]"] - ref[h].imag) < 1e-12
    print("icfft64_in24_full ok", op_counts(o))
    # c2r
    zz = rng.standard_normal(12) + 1j * rng.standard_normal(12)
    g, o = build_c2r64_in12()
    inp = {f"zre[{k}]": zz[k].real for k in range(12)}
    inp.update({f"zim[{k}]": zz[k].imag for k in range(1, 12)})
    res = evaluate(o, inp)
    w = np.arange(64)
    zz0 = zz.copy()
    zz0[0] = zz0[0].real
    ref = np.real(sum(zz0[k] * np.exp(2j * np.pi * k * w / 64) for k in range(12)))
    for i in range(64):
        assert abs(res[f"y[{i}]"] - ref[i]) < 1e-12, (i, res[f"y[{i}]"], ref[i])
    print("c2r64_in12 ok", op_counts(o))


HEADER = '''// GENERATED by gen_codelets.py -- do not edit.  Pruned 64-point DFT codelets for the FNO
// spectral layer (reference src/models/fno/fno2d.py:59-82 keeps 24x12 of 64x33 rfft2 modes).
// Straight-line code, twiddles are float64-derived literals, indices are compile-time constants.
#pragma once
#ifndef FNO_HD
#if defined(__CUDACC__)
#define FNO_HD __host__ __device__ __forceinline__
#else
#define FNO_HD inline
#endif
#endif

namespace fno_codelets {

// c*a + x and c*a - x.  Scalar types: plain expressions (the compiler contracts them to one FMA).  A packed two-lane type
// provides its own overloads (found by argument-dependent lookup), e.g. f32x2 in fno_dft_fwd.cu -> one FFMA2.
template <typename T>
FNO_HD T fno_fma(T c, T a, T x) { return c * a + x; }
template <typename T>
FNO_HD T fno_fms(T c, T a, T x) { return c * a - x; }

'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--selftest", action="store_true")
    ap.add_argument("-o", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "fft_codelets.cuh"))
    a = ap.parse_args()
    if a.selftest:
        selftest()
        return
    parts = [HEADER]
    for name, sig, outs, doc in all_codelets():
        parts.append(emit(name, sig, outs, doc))
        parts.append("")
    parts.append("}  // namespace fno_codelets\n")
    with open(a.o, "w") as f:
        f.write("\n".join(parts))
    print("wrote", a.o)


if __name__ == "__main__":
    sys.setrecursionlimit(10000)
    main()
