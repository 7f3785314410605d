#!/usr/bin/env python
"""Generator for halo2_gpu_specific_b200/csrc/fp_gen.cuh: Montgomery squaring and Karatsuba multiplication on
8 x 32-bit limbs in "separated operand scanning" form (full 512-bit product first, then 8 reduction rows).

  fp_sqr_sos   : 28 cross products (doubled) + 8 squares + 64 reduction MACs = 100 wide MACs (+ 8 narrow) instead of 128 (+ 8)
  fp_mul_kara  : one Karatsuba level on the 256-bit operands: 3 x (4 x 4 limbs) = 48 product MACs + 64 reduction MACs = 112

The reduction works on ABSOLUTE limb positions x0..x15: row i adds m_i * p at positions i..i+8 through two carry
chains (even / odd modulus limbs, each a run of mad.lo.cc / madc.hi.cc pairs that ptxas fuses into IMAD.WIDE), and the
carry out of each chain is parked in a separate small counter k_{i+8} / k_{i+9} instead of rippling through the high
half of the product; the counters are added back once at the end.  No accumulator shifting is needed.

Every carry chain is checked in a Python emulation of the PTX carry semantics (no instruction without .cc may lose a
carry unless marked as an intended wrap) against big-int arithmetic before the header is printed.

    python tools/gen_fp.py            # self-test
    python tools/gen_fp.py --write    # self-test + rewrite csrc/fp_gen.cuh
"""
import os
import random
import re
import sys
from collections import Counter

M32 = (1 << 32) - 1
Q_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


class Gen:
    def __init__(self, inputs):
        self.chains = []          # list of lists of (op, dst, *src)
        self.cur = None
        self.written = set(inputs)
        self.small = set()        # limbs that only ever received carries
        self.inputs = set(inputs)

    def begin(self):
        self.cur = []

    def end(self):
        if self.cur:
            self.chains.append(self.cur)
        self.cur = None

    def val(self, v):
        return v if v in self.written else '0'

    def emit(self, op, dst, *src):
        self.cur.append((op, dst) + tuple(src))
        is_carry_sink = op.startswith('addc') and src[1] == '0' and src[0] in ('0', dst)
        if is_carry_sink and (dst not in self.written or dst in self.small):
            self.small.add(dst)
        else:
            self.small.discard(dst)
        self.written.add(dst)

    def one(self, op, dst, *src):
        self.begin()
        self.emit(op, dst, *src)
        self.end()


def mac_chain(g, prods, acc, top_limit):
    """prods: [(pos, x, y)] ascending, pairs (pos, pos + 1) stepping by 2; acc(pos) -> variable name"""
    g.begin()
    n = len(prods)
    for k, (pos, x, y) in enumerate(prods):
        lo, hi = acc(pos), acc(pos + 1)
        g.emit('mad.lo.cc' if k == 0 else 'madc.lo.cc', lo, x, y, g.val(lo))
        if k < n - 1:
            g.emit('madc.hi.cc', hi, x, y, g.val(hi))
            continue
        hi_fresh = hi not in g.written
        if hi_fresh or pos + 2 > top_limit:
            g.emit('madc.hi', hi, x, y, g.val(hi))
        else:
            g.emit('madc.hi.cc', hi, x, y, g.val(hi))
            nxt = acc(pos + 2)
            assert nxt not in g.written or nxt in g.small, (nxt, 'ripple target neither fresh nor small')
            g.emit('addc', nxt, g.val(nxt), '0')
    g.end()


def add_chain(g, dsts, xs, ys, last_cc=False, wrap_last=False):
    g.begin()
    n = len(dsts)
    for i, (d, x, y) in enumerate(zip(dsts, xs, ys)):
        if i == 0:
            op = 'add.cc' if n > 1 or last_cc else 'add'
        elif i < n - 1 or last_cc:
            op = 'addc.cc'
        else:
            op = 'addc.wrap' if wrap_last else 'addc'
        g.emit(op, d, x, y)
    g.end()


def redc(g):
    """x0..x15 (merged product, pairs (2t, 2t+1)) -> r0..r7 = (x + sum m_i p 2^(32 i)) / 2^256, < 2p.

    IMAD.WIDE wants its 64-bit addend and destination in aligned register pairs, so every limb must keep ONE
    partner for its whole life or ptxas inserts moves (the first version of this function, which added both the even
    and the odd modulus limbs into x, cost 53 moves per product).  Two accumulators with fixed pairings: x (pairs at
    even positions) and o (pairs at odd positions, o1..o16, empty at the start).  A round clears TWO limbs
    (m_lo for position 2r, m_hi for position 2r + 1), so the alignment never flips and nothing is shifted:
        x += m_lo * p_even        o += m_lo * p_odd         (clears position 2r)
        o += m_hi * p_even        x += m_hi * p_odd         (clears position 2r + 1; its carry enters the x chain)
    Carries out of the chains are parked in small counters k_pos (positions >= 8 are never read by a later m) and
    added once at the end."""
    def chain(acc, first_pos, m, js, carry_in=None, first_addend=None, first_dst=None):
        g.begin()
        if carry_in is not None:
            g.emit('add.cc.wrap', carry_in[0], carry_in[1], carry_in[2])    # the sum is discarded; its carry is what matters
        for n_, j in enumerate(js):
            pos = first_pos + n_ * 2
            lo, hi = f'{acc}{pos}', f'{acc}{pos + 1}'
            first = (n_ == 0 and carry_in is None)
            src_lo = first_addend if (n_ == 0 and first_addend is not None) else g.val(lo)
            dst_lo = first_dst if (n_ == 0 and first_dst is not None) else lo
            g.emit('mad.lo.cc' if first else 'madc.lo.cc', dst_lo, f'p{j}', m, src_lo)
            g.emit('madc.hi.cc', hi, f'p{j}', m, g.val(hi))
        top = first_pos + 2 * len(js)
        g.emit('addc', f'k{top}', g.val(f'k{top}'), '0')
        g.end()

    for r_ in range(4):
        b = 2 * r_
        # position b holds x_b and (r > 0) o_b, the high limb of o's pair (b-1, b) whose low limb is already clear
        if r_ == 0:
            fb = f'x{b}'
        else:
            fb = f'fb{r_}'
            g.one('add.wrap', fb, f'x{b}', f'o{b}')
        g.one('mul.lo', f'ml{r_}', fb, 'inv')
        chain('o', b + 1, f'ml{r_}', [1, 3, 5, 7],                  # pairs (b+1, b+2) .. (b+7, b+8) -> k(b+9)
              carry_in=(f'dy{r_}', f'x{b}', f'o{b}') if r_ else None)
        chain('x', b, f'ml{r_}', [0, 2, 4, 6], first_addend=fb, first_dst=f'zb{r_}')   # -> k(b+8); zb == 0
        g.one('add.wrap', f'sh{r_}', f'x{b + 1}', f'o{b + 1}')
        g.one('mul.lo', f'mh{r_}', f'sh{r_}', 'inv')
        chain('o', b + 1, f'mh{r_}', [0, 2, 4, 6])                  # pairs (b+1, b+2) .. (b+7, b+8) -> k(b+9)
        chain('x', b + 2, f'mh{r_}', [1, 3, 5, 7],                  # pairs (b+2, b+3) .. (b+8, b+9) -> k(b+10)
              carry_in=(f'dz{r_}', f'x{b + 1}', f'o{b + 1}'))
    # r = x[8..15] + o[8..15] + k[8..15]   (x16, o16, k16 are zero: the total is below 2^512)
    add_chain(g, [f'v{j}' for j in range(8)], [g.val(f'x{8 + j}') for j in range(8)], [g.val(f'o{8 + j}') for j in range(8)])
    add_chain(g, [f'r{j}' for j in range(8)], [f'v{j}' for j in range(8)], [g.val(f'k{8 + j}') for j in range(8)])


def prod4(g, xs, ys, even_names, odd_prefix):
    """4 x 4 limb product into even_names[0..7] (after the merge the full product); the odd-aligned accumulator
    lives in odd_prefix + position.  even_names must be unwritten."""
    acc = lambda pos: even_names[pos] if False else None  # noqa: E731  (placeholder, replaced below)

    def acc_even(pos):
        return even_names[pos]

    def acc_odd(pos):
        return f'{odd_prefix}{pos}'

    for j in range(4):
        prods = [(i + j, xs[i], ys[j]) for i in range(4)]
        pe = [p for p in prods if p[0] % 2 == 0]
        po = [p for p in prods if p[0] % 2 == 1]
        # an even-aligned pair (pos, pos + 1) lives entirely in the even accumulator, an odd-aligned one in the odd one
        mac_chain(g, pe, acc_even, top_limit=7)
        mac_chain(g, po, acc_odd, top_limit=7)
    # merge: even[1..7] += odd[1..7]
    g.begin()
    for pos in range(1, 8):
        op = 'add.cc' if pos == 1 else ('addc.cc' if pos < 7 else 'addc')
        g.emit(op, even_names[pos], g.val(even_names[pos]), g.val(acc_odd(pos)))
    g.end()


def build_sqr():
    g = Gen([f'a{i}' for i in range(8)] + [f'p{i}' for i in range(8)] + ['inv'])
    acc = lambda pos: (f'ce{pos}' if pos % 2 == 0 else f'co{pos}')  # noqa: E731
    # pairs at even positions live in ce (ce_pos, ce_pos+1), odd ones in co
    def acc_e(pos): return f'ce{pos}'
    def acc_o(pos): return f'co{pos}'
    for i in range(7):
        prods = [(i + j, f'a{i}', f'a{j}') for j in range(i + 1, 8)]
        pe = [p for p in prods if p[0] % 2 == 0]
        po = [p for p in prods if p[0] % 2 == 1]
        if pe:
            mac_chain(g, pe, acc_e, top_limit=15)
        if po:
            mac_chain(g, po, acc_o, top_limit=15)
    # c = ce + co (positions 1..15), d = 2c
    g.begin()
    for pos in range(1, 16):
        op = 'add.cc' if pos == 1 else ('addc.cc' if pos < 15 else 'addc')
        g.emit(op, f'c{pos}', g.val(f'ce{pos}'), g.val(f'co{pos}'))
    g.end()
    g.begin()
    for pos in range(1, 16):
        op = 'add.cc' if pos == 1 else ('addc.cc' if pos < 15 else 'addc')
        g.emit(op, f'd{pos}', f'c{pos}', f'c{pos}')
    g.end()
    # x = d + sum a_i^2 2^(64 i)
    g.begin()
    for i in range(8):
        g.emit('mad.lo.cc' if i == 0 else 'madc.lo.cc', f'x{2 * i}', f'a{i}', f'a{i}', g.val(f'd{2 * i}') if i else '0')
        g.emit('madc.hi.cc' if i < 7 else 'madc.hi', f'x{2 * i + 1}', f'a{i}', f'a{i}', f'd{2 * i + 1}')
    g.end()
    redc(g)
    return g


def build_mul():
    g = Gen([f'a{i}' for i in range(8)] + [f'b{i}' for i in range(8)] + [f'p{i}' for i in range(8)] + ['inv'])
    A = [f'a{i}' for i in range(8)]
    B = [f'b{i}' for i in range(8)]
    prod4(g, A[:4], B[:4], [f'x{i}' for i in range(8)], 'z0o')
    prod4(g, A[4:], B[4:], [f'x{8 + i}' for i in range(8)], 'z2o')
    # |aL - aH|, |bL - bH| and their signs (mask = 0xffffffff when negative)
    for nm, V in (('a', A), ('b', B)):
        g.begin()
        for i in range(4):
            g.emit('sub.cc' if i == 0 else 'subc.cc', f'd{nm}{i}', V[i], V[4 + i])
        g.emit('subc.wrap', f'm{nm}', '0', '0')
        g.end()
        for i in range(4):
            g.one('xor', f't{nm}{i}', f'd{nm}{i}', f'm{nm}')
        g.begin()
        g.emit('add.cc', f'dump{nm}', f'm{nm}', f'm{nm}')          # carry = 1 when the difference was negative
        for i in range(4):
            g.emit('addc.cc' if i < 3 else 'addc', f'u{nm}{i}', f't{nm}{i}', '0')
        g.end()
    prod4(g, [f'ua{i}' for i in range(4)], [f'ub{i}' for i in range(4)], [f'mm{i}' for i in range(8)], 'mmo')
    # aL bH + aH bL = z0 + z2 - (aL - aH)(bL - bH) = z0 + z2 -/+ M  (minus when the signs are equal)
    g.one('xor', 'tsign', 'ma', 'mb')
    g.one('xor', 'usub', 'tsign', str(M32))                        # all ones when M is subtracted
    add_chain(g, [f's{i}' for i in range(9)], [f'x{i}' for i in range(8)] + ['0'], [f'x{8 + i}' for i in range(8)] + ['0'])
    for i in range(8):
        g.one('xor', f'mx{i}', f'mm{i}', 'usub')
    g.begin()
    g.emit('add.cc', 'dumps', 'usub', 'usub')
    for i in range(8):
        g.emit('addc.cc', f'mid{i}', f's{i}', f'mx{i}')
    g.emit('addc.wrap', 'mid8', 's8', 'usub')
    g.end()
    # x += mid << 128
    g.begin()
    for i in range(12):
        pos = 4 + i
        op = 'add.cc' if i == 0 else ('addc.cc' if i < 11 else 'addc')
        g.emit(op, f'x{pos}', f'x{pos}', f'mid{i}' if i < 9 else '0')
    g.end()
    redc(g)
    return g


def emulate(g, env):
    for ch in g.chains:
        cf = 0
        for ins in ch:
            op, dst = ins[0], ins[1]
            s = [env[x] if not x.isdigit() else int(x) for x in ins[2:]]
            wrap = '.wrap' in op
            base = op.replace('.cc', '').replace('.wrap', '')
            if base == 'add': v = s[0] + s[1]
            elif base == 'addc': v = s[0] + s[1] + cf
            elif base == 'sub': v = s[0] - s[1]
            elif base == 'subc': v = s[0] - s[1] - cf
            elif base == 'mad.lo': v = ((s[0] * s[1]) & M32) + s[2]
            elif base == 'madc.lo': v = ((s[0] * s[1]) & M32) + s[2] + cf
            elif base == 'mad.hi': v = ((s[0] * s[1]) >> 32) + s[2]
            elif base == 'madc.hi': v = ((s[0] * s[1]) >> 32) + s[2] + cf
            elif base == 'mul.lo': v = (s[0] * s[1]) & M32
            elif base == 'xor': v = s[0] ^ s[1]
            else: raise ValueError(op)
            if '.cc' in op:
                cf = 1 if (v >> 32) != 0 else 0          # carry out, or borrow for sub (v negative)
            elif not wrap and (v < 0 or v >> 32):
                raise AssertionError(('lost carry', ins, hex(v)))
            env[dst] = v & M32
    return env


def count(g):
    c = Counter()
    for ch in g.chains:
        for ins in ch:
            c[ins[0].replace('.cc', '').replace('.wrap', '').replace('madc', 'mad').replace('addc', 'add').replace('subc', 'sub')] += 1
    return c


def selftest(n_random=5000):
    out = {}
    for name, g, two in (('sqr', build_sqr(), False), ('mul', build_mul(), True)):
        for p, inv in ((Q_MOD, 0xe4866389), (R_MOD, 0xefffffff)):
            assert (p * inv + 1) % (1 << 32) == 0
            rnd = random.Random(7)
            edge = [0, 1, 2, p - 1, p - 2, (1 << 253), M32, (1 << 64) - 1, (1 << 128) - 1, (1 << 128), (1 << 128) + 1,
                    ((1 << 128) - 1) << 128 & ((1 << 254) - 1), p >> 1, int('ffffffff00000000' * 4, 16) % p,
                    int('00000000ffffffff' * 4, 16) % p, (M32 << 128) | M32, ((1 << 127) << 128) % p | 5]
            edge = [e % p for e in edge]
            cases = [(x, y) for x in edge for y in edge] + [(rnd.randrange(p), rnd.randrange(p)) for _ in range(n_random)]
            for x, y in cases:
                env = {'inv': inv}
                for i in range(8):
                    env[f'a{i}'] = (x >> 32 * i) & M32
                    env[f'b{i}'] = (y >> 32 * i) & M32
                    env[f'p{i}'] = (p >> 32 * i) & M32
                emulate(g, env)
                res = sum(env[f'r{i}'] << 32 * i for i in range(8))
                yy = y if two else x
                assert res < 2 * p and (res << 256) % p == (x * yy) % p, (name, hex(x), hex(yy))
                for t_ in ('x16', 'x17', 'o16', 'o17', 'k16', 'k17', 'k18'):
                    assert env.get(t_, 0) == 0, t_
                for t_ in ('zb0', 'zb1', 'zb2', 'zb3'):
                    assert env[t_] == 0, t_
        out[name] = dict(count(g))
    return out


# ------------------------------------------------------------------------------------------------ CUDA printer
PTX = {'add.cc.wrap': 'add.cc.u32', 'add.wrap': 'add.u32', 'add.cc': 'add.cc.u32', 'addc.cc': 'addc.cc.u32', 'addc': 'addc.u32', 'addc.wrap': 'addc.u32', 'add': 'add.u32',
       'sub.cc': 'sub.cc.u32', 'subc.cc': 'subc.cc.u32', 'subc': 'subc.u32', 'subc.wrap': 'subc.u32',
       'mad.lo.cc': 'mad.lo.cc.u32', 'madc.lo.cc': 'madc.lo.cc.u32', 'madc.lo': 'madc.lo.u32', 'mad.lo': 'mad.lo.u32',
       'mad.hi.cc': 'mad.hi.cc.u32', 'madc.hi.cc': 'madc.hi.cc.u32', 'madc.hi': 'madc.hi.u32'}


def cexpr(v):
    m = re.fullmatch(r'([a-z]+?)(\d+)', v)
    if m and m.group(1) in ('a', 'b', 'r') and len(m.group(2)) == 1:
        return f'{m.group(1)}.v[{m.group(2)}]'
    return v


def is_imm(v):
    return v == 'inv' or re.fullmatch(r'p\d', v) is not None


def render_function(g, name, args, doc):
    local = []
    for ch in g.chains:
        for ins in ch:
            d = ins[1]
            if cexpr(d) == d and d not in local:
                local.append(d)
    lines = [doc, 'template <class P>', f'__device__ __forceinline__ Fp<P> {name}({args}) {{',
             '    Fp<P> r;', '    uint32_t ' + ', '.join(local) + ';']
    for ch in g.chains:
        if len(ch) == 1 and ch[0][0] in ('xor', 'mul.lo'):
            op, d, x, y = ch[0]
            rhs = lambda s: (f'P::{s}' if is_imm(s) else (f'0x{int(s):x}u' if s.isdigit() else cexpr(s)))  # noqa: E731
            lines.append(f'    {cexpr(d)} = {rhs(x)} {"^" if op == "xor" else "*"} {rhs(y)};')
            continue
        order, writes, rbw = [], set(), set()
        for ins in ch:
            for s in ins[2:]:
                if s.isdigit():
                    continue
                if s not in order:
                    order.append(s)
                if s not in writes:
                    rbw.add(s)
            if ins[1] not in order:
                order.append(ins[1])
            writes.add(ins[1])
        outs = [v for v in order if v in writes]
        ins_ = [v for v in order if v not in writes]
        idx, ops_out, ops_in = {}, [], []
        for v in outs:
            idx[v] = len(idx)
            ops_out.append(('"+r"' if v in rbw else '"=&r"') + f'({cexpr(v)})')
        for v in ins_:
            idx[v] = len(idx)
            ops_in.append(f'"n"(P::{v})' if is_imm(v) else f'"r"({cexpr(v)})')
        body = []
        for ins in ch:
            a_ = [f'%{idx[ins[1]]}'] + [(s if s.isdigit() else f'%{idx[s]}') for s in ins[2:]]
            body.append(f'{PTX[ins[0]]} {", ".join(a_)};')
        text = '"' + '\\n\\t"\n        "'.join(body) + '"'
        lines.append(f'    asm({text}\n        : {", ".join(ops_out)}\n        : {", ".join(ops_in) if ops_in else ""});')
    lines += ['    fp_reduce_once<P>(r.v);', '    return r;', '}']
    return '\n'.join(lines)


def render():
    hdr = '''// fp_gen.cuh -- GENERATED by tools/gen_fp.py; do not edit by hand.
//
// Montgomery squaring and Karatsuba multiplication in separated-operand-scanning form: the 512-bit product is
// formed first (squaring: 28 doubled cross products + 8 squares; multiplication: one Karatsuba level, three
// 4 x 4-limb products), then reduced by 8 rows m_i * p on absolute limb positions whose chain carries are parked
// in small counters and added back once.  100 / 112 wide multiply-accumulates instead of the 128 of the
// interleaved (CIOS) product in fp.cuh, paid for with additions on the otherwise idle ALU pipe.
#pragma once
#include "fp.cuh"

namespace b2 {

'''
    s = hdr
    s += render_function(build_sqr(), 'fp_sqr_sos', 'const Fp<P>& a',
                         '// r = a*a*2^-256 mod p, fully reduced.  a < p.') + '\n\n'
    s += render_function(build_mul(), 'fp_mul_kara', 'const Fp<P>& a, const Fp<P>& b',
                         '// r = a*b*2^-256 mod p, fully reduced.  a, b < p.') + '\n\n'
    s += '}  // namespace b2\n'
    return s


if __name__ == '__main__':
    print(selftest())
    if '--write' in sys.argv:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'halo2_gpu_specific_b200', 'csrc',
                            'fp_gen.cuh')
        open(path, 'w').write(render())
        print('wrote', path)
