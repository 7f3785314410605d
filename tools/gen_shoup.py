#!/usr/bin/env python
"""Generator for halo2_gpu_specific_b200/csrc/fp_shoup.cuh (Shoup multiplication by a precomputed constant).

The carry chains are laid out here, checked limb by limb in a Python emulation of the PTX add/mad
carry semantics (every chain: no carry may be lost except out of bit 256), and then printed as inline
asm blocks (one asm statement per carry chain so that ptxas fuses each mad.lo.cc / madc.hi.cc pair into
one IMAD.WIDE).

    python tools/gen_shoup.py            # self-test (emulation against big-int arithmetic)
    python tools/gen_shoup.py --write    # self-test + rewrite csrc/fp_shoup.cuh
"""

import random, sys
M32 = (1 << 32) - 1

class Gen:
    def __init__(self):
        self.chains = []      # list of list of instr
        self.cur = None
        self.fresh = {}       # var -> bool written
        self.small = set()    # limbs that only ever received carries
    def begin(self): self.cur = []
    def end(self):
        if self.cur: self.chains.append(self.cur)
        self.cur = None
    def val(self, v):  # addend operand: 0 if never written
        return v if self.fresh.get(v) else '0'
    def emit(self, op, dst, *src):
        self.cur.append((op, dst) + tuple(src))
        if op.startswith('addc') and src[1] == '0' and (src[0] == '0' or src[0] == dst) and (not self.fresh.get(dst) or dst in self.small):
            self.small.add(dst)
        else:
            self.small.discard(dst)
        self.fresh[dst] = True

def mac_chain(g, prods, acc, top_limit, drop_top_carry=False, lo_only_pos=None):
    """prods: list of (pos, x, y) sorted ascending by pos, contiguous pairs (pos, pos+1) stepping by 2.
    acc(pos) -> variable name.  A trailing product at lo_only_pos contributes only its low half."""
    g.begin()
    first = True
    last_hi_fresh = True
    last_pos = None
    n = len(prods)
    for k, (pos, x, y) in enumerate(prods):
        lo_only = (lo_only_pos is not None and pos == lo_only_pos)
        lo, hi = acc(pos), acc(pos + 1) if not lo_only else None
        final = (k == n - 1)
        if lo_only:
            assert final
            op = 'mad.lo' if first else 'madc.lo'
            g.emit(op, lo, x, y, g.val(lo))       # no carry out (discarded)
            first = False
            last_pos = None
            break
        g.emit(('mad.lo.cc' if first else 'madc.lo.cc'), lo, x, y, g.val(lo))
        first = False
        hi_was_fresh = not g.fresh.get(hi)
        if final:
            if drop_top_carry or hi_was_fresh or pos + 2 > top_limit:
                g.emit('madc.hi', hi, x, y, g.val(hi))
                last_pos = None
            else:
                g.emit('madc.hi.cc', hi, x, y, g.val(hi))
                nxt = acc(pos + 2)
                assert (not g.fresh.get(nxt)) or nxt in g.small, (nxt, 'ripple target neither fresh nor small')
                g.emit('addc', nxt, g.val(nxt), '0')
        else:
            g.emit('madc.hi.cc', hi, x, y, g.val(hi))
    g.end()

def build():
    g = Gen()
    for i in range(8):
        for nm in ('a', 'w', 'wp'): g.fresh[f'{nm}{i}'] = True
    # ---- part 1: q = floor(S / 2^256)
    QE = lambda pos: f'qe{pos}'
    QO = lambda pos: f'qo{pos}'
    accq = lambda pos: (QE(pos) if pos % 2 == 0 else QO(pos))
    # hi32 of the i+j == 6 products, summed into (qo7, qo8)
    g.begin()
    pairs6 = [(i, 6 - i) for i in range(7)]
    i, j = pairs6[0]
    g.emit('mul.hi', 'qo7', f'a{i}', f'wp{j}')
    g.end()
    first = True
    for (i, j) in pairs6[1:]:
        g.begin()
        g.emit('mad.hi.cc', 'qo7', f'a{i}', f'wp{j}', 'qo7')
        g.emit('addc', 'qo8', g.val('qo8'), '0')
        g.end()
    for j in range(8):
        prods = [(i + j, f'a{i}', f'wp{j}') for i in range(7 - j, 8)]
        for par, acc in ((1, QO), (0, QE)):
            pr = [p for p in prods if p[0] % 2 == par]
            if pr: mac_chain(g, pr, acc, top_limit=15)
    # q[k] = qe[8+k] + qo[8+k]
    g.begin()
    for k in range(8):
        op = 'add.cc' if k == 0 else ('addc.cc' if k < 7 else 'addc')
        g.emit(op, f'q{k}', g.val(QE(8 + k)) if (8 + k) % 2 == 0 or True else '0', g.val(QO(8 + k)))
    g.end()
    # ---- part 2: r = lo256(a*w + q*pp)
    RE = lambda pos: f're{pos}'
    RO = lambda pos: f'ro{pos}'
    for (xs, ys) in (('a', 'w'), ('pp', 'q')):
        for i in range(8):     # scalar ys_i, vector xs_j, j <= 7 - i
            prods = [(i + j, f'{xs}{j}', f'{ys}{i}') for j in range(0, 8 - i)]
            pe = [p for p in prods if p[0] % 2 == 0]
            po = [p for p in prods if p[0] % 2 == 1]
            if pe: mac_chain(g, pe, RE, top_limit=7, drop_top_carry=True)
            if po: mac_chain(g, po, RO, top_limit=7, drop_top_carry=True, lo_only_pos=7)
    g.begin()
    g.emit('add.cc', 'r1', 're1', 'ro1')
    for k in range(2, 8):
        g.emit('addc.cc' if k < 7 else 'addc', f'r{k}', f're{k}', f'ro{k}')
    g.end()
    return g

def emulate(g, env):
    for ch in g.chains:
        cf = 0
        for ins in ch:
            op, dst = ins[0], ins[1]
            s = [env[x] if not x.isdigit() else int(x) for x in ins[2:]]
            base = op.replace('.cc', '')
            cin = cf if base in ('addc', 'madc.lo', 'madc.hi') else 0
            if base in ('add', 'addc'): v = s[0] + s[1] + cin
            elif base in ('mad.lo', 'madc.lo'): v = ((s[0] * s[1]) & M32) + s[2] + cin
            elif base in ('mad.hi', 'madc.hi'): v = ((s[0] * s[1]) >> 32) + s[2] + cin
            elif base == 'mul.hi': v = (s[0] * s[1]) >> 32
            else: raise ValueError(op)
            if op.endswith('.cc'): cf = v >> 32
            else:
                # an instruction without .cc must not lose a carry unless it is a discarded top limb
                if (v >> 32) and not (dst in ('re7', 'ro7', 'r7')): raise AssertionError(('lost carry', ins))
            env[dst] = v & M32
    return env

def count(g):
    from collections import Counter
    c = Counter()
    for ch in g.chains:
        for ins in ch: c[ins[0].replace('.cc', '').replace('madc', 'mad').replace('addc', 'add')] += 1
    return c

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
BETA = 1 << 256
PTX = {'add.cc': 'add.cc.u32', 'addc.cc': 'addc.cc.u32', 'addc': 'addc.u32', 'add': 'add.u32',
       'mad.lo.cc': 'mad.lo.cc.u32', 'madc.lo.cc': 'madc.lo.cc.u32', 'madc.lo': 'madc.lo.u32', 'mad.lo': 'mad.lo.u32',
       'mad.hi.cc': 'mad.hi.cc.u32', 'madc.hi.cc': 'madc.hi.cc.u32', 'madc.hi': 'madc.hi.u32', 'mul.hi': 'mul.hi.u32'}

def cname(v):
    # map generator variable to C expression
    import re
    m = re.fullmatch(r'([a-z]+)(\d+)', v)
    nm, i = m.group(1), int(m.group(2))
    if nm == 'a': return f'a.v[{i}]'
    if nm == 'w': return f'w.v[{i}]'
    if nm == 'wp': return f'wp.v[{i}]'
    if nm == 'pp': return None  # immediate
    if nm in ('qe', 'qo'): return f'{nm}[{i - 7}]'
    if nm in ('re', 'ro', 'q'): return f'{nm}[{i}]'
    if nm == 'r': return f'r.v[{i}]'
    raise ValueError(v)

def emit_function(g, pp):
    out = []
    written = set()
    for ch in g.chains:
        # classify variables
        order = []
        reads_before_write, writes = set(), set()
        for ins in ch:
            for s in ins[2:]:
                if s.isdigit(): continue
                if s not in order: order.append(s)
                if s not in writes: reads_before_write.add(s)
            d = ins[1]
            if d not in order: order.append(d)
            writes.add(d)
        outs = [v for v in order if v in writes]
        ins_ = [v for v in order if v not in writes]
        idx = {}
        ops_out, ops_in = [], []
        for v in outs:
            idx[v] = len(idx)
            ops_out.append(('"+r"' if v in reads_before_write else '"=&r"') + f'({cname(v)})')
        for v in ins_:
            idx[v] = len(idx)
            if v.startswith('pp'):
                ops_in.append(f'"n"(0x{(pp >> 32 * int(v[2:])) & 0xffffffff:08x}u)')
            else:
                ops_in.append(f'"r"({cname(v)})')
        lines = []
        for ins in ch:
            args = [f'%{idx[ins[1]]}'] + [(s if s.isdigit() else f'%{idx[s]}') for s in ins[2:]]
            lines.append(f'{PTX[ins[0]]} {", ".join(args)};')
        body = '"' + '\\n\\t"\n        "'.join(lines) + '"'
        out.append(f'    asm({body}\n        : {", ".join(ops_out)}\n        : {", ".join(ops_in)});')
    return '\n'.join(out)


def render(g):
    r = R_MOD
    pp = BETA - r
    code = emit_function(g, pp)
    hdr = f'''// fp_shoup.cuh -- GENERATED by tools/gen_shoup.py; do not edit by hand.
//
// Multiplication of an Fr element by a constant w known in advance (NTT twiddles), after Shoup /
// Harvey: with wp = floor(w * 2^256 / r) precomputed,
//     q  ~ floor(a * wp / 2^256)            (upper half only: partial products with i + j >= 7, plus the high
//                                            words of the i + j == 6 products; q is at most 2 below floor(a*w/r))
//     t  = (a * w - q * r) mod 2^256        (lower halves only; t < 3r)
// i.e. 92 wide multiply-accumulates + 7 high + 16 low products instead of the 128 + 8 of a Montgomery product.
// w is a plain integer; a and the result share whatever representation a has (for a = x*R: result = (x*w)*R).
#pragma once
#include "fp.cuh"

namespace b2 {{

// t = a * w mod r, reduced to [0, r).  a < 2^256, w < r, wp = floor(w * 2^256 / r).
__device__ __forceinline__ Fr fr_mul_shoup(const Fr& a, const Fr& w, const Fr& wp) {{
    uint32_t qe[9], qo[9], q[8], re[8], ro[8];   // qe/qo index = position - 7
    Fr r;
{code}
    r.v[0] = re[0];
    // r < 3p: subtract 2p if possible, then p if possible
    {{
        uint32_t t[8], borrow;
        asm("sub.cc.u32 %0, %9, %17;\\n\\t"
            "subc.cc.u32 %1, %10, %18;\\n\\t"
            "subc.cc.u32 %2, %11, %19;\\n\\t"
            "subc.cc.u32 %3, %12, %20;\\n\\t"
            "subc.cc.u32 %4, %13, %21;\\n\\t"
            "subc.cc.u32 %5, %14, %22;\\n\\t"
            "subc.cc.u32 %6, %15, %23;\\n\\t"
            "subc.cc.u32 %7, %16, %24;\\n\\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
              "=r"(borrow)
            : "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]),
              {", ".join(f'"n"(0x{((2 * r) >> 32 * i) & 0xffffffff:08x}u)' for i in range(8))});
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = borrow ? r.v[i] : t[i];
    }}
    fp_reduce_once<FrParams>(r.v);
    return r;
}}

// Lazy form for the Harvey-style NTT butterflies: the same quotient estimate and remainder WITHOUT the final
// conditional subtractions.  For ANY a < 2^256 (in particular a < 4r): the result is congruent to a * w mod r and
// lies in [0, 4r) -- t0 = a*w - floor(a*wp / 2^256) * r < r * (1 + a / 2^256) < 2r, and the estimate is at most 2 below
// the exact floor.  4r < 2^256 for BN254's r, so nothing wraps.
__device__ __forceinline__ Fr fr_mul_shoup_lazy(const Fr& a, const Fr& w, const Fr& wp) {{
    uint32_t qe[9], qo[9], q[8], re[8], ro[8];   // qe/qo index = position - 7
    Fr r;
{code}
    r.v[0] = re[0];
    return r;
}}

// Shoup companion of a twiddle given in Montgomery form wm = w * 2^256 mod r:
//   floor(w * 2^256 / r) = (w * 2^256 - wm) / r = wm * (-r^-1) mod 2^256   (the division is exact)
__device__ __forceinline__ Fr fr_shoup_companion(const Fr& wm) {{
    const uint32_t ninv[8] = {{0xefffffffu, 0xc2e1f593u, 0x4c6911b3u, 0x6586864bu,
                              0x99062391u, 0xe39a9828u, 0x0d8341b2u, 0x73f82f1du}};   // -r^-1 mod 2^256
    Fr o = Fr::zero();
#pragma unroll
    for (int i = 0; i < 8; i++) {{
        uint32_t carry = 0;
#pragma unroll
        for (int j = 0; j + i < 8; j++) {{
            const unsigned long long t = (unsigned long long)wm.v[i] * ninv[j] + o.v[i + j] + carry;
            o.v[i + j] = (uint32_t)t;
            carry = (uint32_t)(t >> 32);
        }}
    }}
    return o;
}}

}}  // namespace b2
'''
    return hdr


def selftest(g, n_random=20000):
    r = R_MOD
    pp = BETA - r
    random.seed(5)
    edge = [0, 1, 2, r - 1, r - 2, (1 << 224) - 1, (1 << 253), M32, (1 << 64) - 1, r >> 1,
            int('ffffffff00000000' * 4, 16) % r, int('00000000ffffffff' * 4, 16) % r]
    cases = [(a, w) for a in edge for w in edge] + [(random.randrange(r), random.randrange(r)) for _ in range(n_random)]
    for a, w in cases:
        wp = (w << 256) // r
        env = {}
        for i in range(8):
            env[f'a{i}'] = (a >> 32 * i) & M32
            env[f'w{i}'] = (w >> 32 * i) & M32
            env[f'wp{i}'] = (wp >> 32 * i) & M32
            env[f'pp{i}'] = (pp >> 32 * i) & M32
        emulate(g, env)
        env['r0'] = env['re0']
        res = sum(env[f'r{i}'] << 32 * i for i in range(8))
        q = sum(env[f'q{i}'] << 32 * i for i in range(8))
        assert 0 <= a * w // r - q <= 2
        assert res == a * w - q * r and res < 3 * r, (hex(a), hex(w))
    # lazy form: unreduced inputs (anything below 2^256; the butterflies keep values below 4r)
    lazy_edge = [4 * r - 1, 4 * r - 2, 3 * r, 2 * r, 2 * r - 1, BETA - 1, BETA - 2, (1 << 255), 3 * r + 12345]
    lazy_cases = [(a, w) for a in lazy_edge for w in edge] + \
                 [(random.randrange(4 * r), random.randrange(r)) for _ in range(n_random)] + \
                 [(random.randrange(BETA), random.randrange(r)) for _ in range(n_random // 4)]
    for a, w in lazy_cases:
        wp = (w << 256) // r
        env = {}
        for i in range(8):
            env[f'a{i}'] = (a >> 32 * i) & M32
            env[f'w{i}'] = (w >> 32 * i) & M32
            env[f'wp{i}'] = (wp >> 32 * i) & M32
            env[f'pp{i}'] = (pp >> 32 * i) & M32
        emulate(g, env)
        env['r0'] = env['re0']
        res = sum(env[f'r{i}'] << 32 * i for i in range(8))
        q = sum(env[f'q{i}'] << 32 * i for i in range(8))
        assert 0 <= a * w // r - q <= 2, (hex(a), hex(w))
        assert res == a * w - q * r and res < 4 * r and res < BETA, (hex(a), hex(w))
    return len(cases) + len(lazy_cases)


if __name__ == '__main__':
    import os
    g = build()
    print(dict(count(g)), 'chains:', len(g.chains))
    print('emulation ok on', selftest(g), 'cases')
    if '--write' in sys.argv:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'halo2_gpu_specific_b200', 'csrc',
                            'fp_shoup.cuh')
        open(path, 'w').write(render(g))
        print('wrote', path)
