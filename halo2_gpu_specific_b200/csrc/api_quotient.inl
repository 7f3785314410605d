// api_quotient.inl -- host side of the quotient-evaluation engine (included by api.cu).
//
// b2_quotient_program_create lowers the reference's Calculation list
// (halo2_proofs/src/plonk/evaluation.rs:46-112) to the four-instruction form that
// quotient_eval_kernel interprets:
//   * Store(x) and column / constant / challenge operands are inlined into their users
//     (the reference wraps every column query in a Store, evaluation.rs:675-710);
//   * LcChallenge, LcTheta / Horner and AddChallenge become ADD / MUL pairs whose challenge
//     operand is an entry of the per-proof challenge table (powers are computed on the host
//     at evaluation time, evaluation.rs:205-211);
//   * instructions whose value never reaches `result` are dropped;
//   * intermediates get shared-memory slots from their live ranges (a slot is recycled after
//     the last read of its value, and a destination may reuse the slot of its own operand);
//   * when the live width costs more resident CTAs than a second slot class does (10 or more slots), the values
//     with the longest live ranges -- sub-expressions a circuit shares between gates that sit far
//     apart in its gate list -- move to a second slot class in global memory (one coalesced,
//     L2-resident 32-byte round trip per use instead of 32 B x 128 threads of shared memory for
//     most of the program); slot indices below n_slots_shared are shared memory, the rest global.

namespace {

constexpr size_t Q_SMEM_LIMIT = 200 * 1024;
// Shared-memory slots a program may use before long-lived values go to the global class: a slot is 4 KB per CTA
// (128 threads x 32 B) and the kernel's registers (70 - 72) allow 7 CTAs per SM, so up to 7 slots (7 x (28 + 1) KB of the
// 228 KB) cost no occupancy; 17 slots are 3 CTAs per SM and 27 % of the kernel's throughput.
constexpr uint32_t Q_SHARED_TARGET = 7;
// ... but the global class has a price too (a persistent one-wave grid, a class test per slot access, the round trips):
// measured (profiles/r2_quotient_slot_classes_ll*.json) a program of 8 live values is 5 % SLOWER as 7 + 1 than with all 8
// in shared memory at 6 CTAs per SM, one of 10 is 2 % faster as 7 + 3 (all shared: 5 CTAs), one of 18 is 1.35x faster,
// one of 34 2.9x.  So the second class is used from 10 live values on.
constexpr uint32_t Q_HYBRID_FROM = 10;
constexpr uint32_t Q_GLOBAL_MIN_LIVE = 24;   // instructions between definition and last use below which a value stays shared

struct QProgram {
    // lowered program
    std::vector<QInstr> instr;
    uint32_t result = 0;
    uint32_t n_slots = 0, n_mul = 0, n_addsub = 0, n_fused = 0;   // n_mul counts both products of a fused instruction
    uint32_t n_slots_shared = 0;       // slots [0, n_slots_shared) live in shared memory, [n_slots_shared, n_slots) in global
    bool uses_x = false;
    std::vector<int32_t> rotations;
    std::vector<uint64_t> constants;                         // 4 limbs each
    std::vector<std::pair<uint32_t, uint32_t>> derived;      // extra challenge entries: (challenge, power)
    uint32_t n_fixed = 0, n_advice = 0, n_instance = 0, n_aux = 0, n_challenges = 0;
    // device copies, per device, created on first use
    struct Dev {
        QInstr* prog = nullptr;
        Fr* constants = nullptr;
    };
    std::map<int, Dev> dev;
    std::mutex mu;
};

std::mutex g_qprog_mu;
std::map<b2_handle_t, QProgram*> g_qprog;

struct QLower {
    const b2_quotient_program_desc& d;
    QProgram& p;
    std::vector<QInstr> raw;           // dst = virtual register id
    std::vector<uint32_t> loc;         // operand word per calc (QK_SLOT index = virtual register)
    uint32_t n_vreg = 0;
    std::string err;

    QLower(const b2_quotient_program_desc& desc, QProgram& prog) : d(desc), p(prog) {}

    bool src(const b2_qsrc& s, uint32_t calc_index, uint32_t* out) {
        char buf[160];
        auto bad = [&](const char* what, uint32_t limit) {
            snprintf(buf, sizeof buf, "calc %u: %s index %u out of range (%u)", calc_index, what, s.index, limit);
            err = buf;
            return false;
        };
        uint32_t col_base = 0;
        switch (s.kind) {
        case B2_Q_CONSTANT:
            if (s.index >= d.n_constants) return bad("constant", d.n_constants);
            *out = q_operand(QK_CONST, s.index, 0);
            return true;
        case B2_Q_INTERMEDIATE:
            if (s.index >= calc_index) return bad("intermediate (must refer to an earlier calculation)", calc_index);
            *out = loc[s.index];
            return true;
        case B2_Q_AUX:
            col_base += d.n_instance;
            if (s.index >= d.n_aux) return bad("aux column", d.n_aux);
            /* fall through */
        case B2_Q_INSTANCE:
            col_base += d.n_advice;
            if (s.kind == B2_Q_INSTANCE && s.index >= d.n_instance) return bad("instance column", d.n_instance);
            /* fall through */
        case B2_Q_ADVICE:
            col_base += d.n_fixed;
            if (s.kind == B2_Q_ADVICE && s.index >= d.n_advice) return bad("advice column", d.n_advice);
            /* fall through */
        case B2_Q_FIXED:
            if (s.kind == B2_Q_FIXED && s.index >= d.n_fixed) return bad("fixed column", d.n_fixed);
            if (s.rotation >= d.n_rotations) {
                snprintf(buf, sizeof buf, "calc %u: rotation index %u out of range (%u)", calc_index, s.rotation,
                         d.n_rotations);
                err = buf;
                return false;
            }
            if (col_base + s.index >= (1u << 20)) return bad("column", 1u << 20);
            *out = q_operand(QK_COLUMN, col_base + s.index, s.rotation);
            return true;
        case B2_Q_CHALLENGE:
            if (s.index >= d.n_challenges) return bad("challenge", d.n_challenges);
            *out = q_operand(QK_CHAL, s.index, 0);
            return true;
        case B2_Q_COSET_X:
            p.uses_x = true;
            *out = q_operand(QK_COSET_X, 0, 0);
            return true;
        default:
            snprintf(buf, sizeof buf, "calc %u: unknown source kind %u", calc_index, s.kind);
            err = buf;
            return false;
        }
    }

    bool challenge(uint32_t ch, uint32_t power, uint32_t calc_index, uint32_t* out) {
        if (ch >= d.n_challenges) {
            char buf[128];
            snprintf(buf, sizeof buf, "calc %u: challenge index %u out of range (%u)", calc_index, ch, d.n_challenges);
            err = buf;
            return false;
        }
        if (power <= 1) {   // evaluation.rs:208-211: x.pow(p) only when p > 1
            *out = q_operand(QK_CHAL, ch, 0);
            return true;
        }
        for (size_t i = 0; i < p.derived.size(); i++)
            if (p.derived[i].first == ch && p.derived[i].second == power) {
                *out = q_operand(QK_CHAL, d.n_challenges + (uint32_t)i, 0);
                return true;
            }
        p.derived.push_back({ch, power});
        *out = q_operand(QK_CHAL, d.n_challenges + (uint32_t)p.derived.size() - 1, 0);
        return true;
    }

    uint32_t emit(uint32_t op, uint32_t a, uint32_t b) {
        QInstr in;
        in.op_dst = op | (n_vreg << 8);
        in.a = a;
        in.b = b;
        in.pad = 0;
        raw.push_back(in);
        return q_operand(QK_SLOT, n_vreg++, 0);
    }

    bool run() {
        if (d.n_rotations > 256) { err = "more than 256 distinct rotations"; return false; }
        if (d.n_constants >= (1u << 20) || d.n_calcs >= (1u << 19)) { err = "program too large"; return false; }
        loc.resize(d.n_calcs);
        for (uint32_t i = 0; i < d.n_calcs; i++) {
            const b2_qcalc& c = d.calcs[i];
            uint32_t a = 0, b = 0, ch = 0;
            if (!src(c.a, i, &a)) return false;
            const bool binary = c.op == B2_QOP_ADD || c.op == B2_QOP_SUB || c.op == B2_QOP_MUL ||
                                c.op == B2_QOP_LC_CHALLENGE || c.op == B2_QOP_MUL_CH_ADD;
            if (binary && !src(c.b, i, &b)) return false;
            switch (c.op) {
            case B2_QOP_ADD: loc[i] = emit(Q_ADD, a, b); break;
            case B2_QOP_SUB: loc[i] = emit(Q_SUB, a, b); break;
            case B2_QOP_MUL: loc[i] = emit(Q_MUL, a, b); break;
            case B2_QOP_NEGATE: loc[i] = emit(Q_NEG, a, 0); break;
            case B2_QOP_LC_CHALLENGE:
                if (!challenge(c.challenge, c.power, i, &ch)) return false;
                loc[i] = emit(Q_MUL, emit(Q_ADD, a, ch), b);
                break;
            case B2_QOP_MUL_CH_ADD:
                if (!challenge(c.challenge, 1, i, &ch)) return false;
                loc[i] = emit(Q_ADD, emit(Q_MUL, a, ch), b);
                break;
            case B2_QOP_ADD_CHALLENGE:
                if (!challenge(c.challenge, 1, i, &ch)) return false;
                loc[i] = emit(Q_ADD, a, ch);
                break;
            case B2_QOP_STORE: loc[i] = a; break;
            default: {
                char buf[96];
                snprintf(buf, sizeof buf, "calc %u: unknown op %u", i, c.op);
                err = buf;
                return false;
            }
            }
        }
        uint32_t res = 0;
        if (!src(d.result, d.n_calcs, &res)) return false;

        // From here on a node has up to four operands (the fused two-product form).
        struct Node { uint32_t op; uint32_t w[4]; };
        auto nops = [](uint32_t op) -> int { return (op == Q_NEG || op == Q_COPY) ? 1 : (op >= Q_MUL2ADD ? 4 : 2); };
        auto is_v = [](uint32_t w) { return (w >> 28) == QK_SLOT; };
        std::vector<Node> node(n_vreg);
        for (uint32_t v = 0; v < n_vreg; v++) node[v] = Node{raw[v].op_dst & 0xffu, {raw[v].a, raw[v].b, 0, 0}};

        // Fusion of a * b +- c * d.  Uses are counted over the instructions the result depends on; a product read only by
        // one ADD / SUB whose other operand is such a product too disappears into that instruction (evaluate_h is sums
        // of products: gate polynomials, the permutation and lookup terms).  B2_Q_NO_FUSE=1 keeps the plain form (A/B).
        static const bool no_fuse = getenv("B2_Q_NO_FUSE") != nullptr;
        if (!no_fuse) {
            std::vector<char> reach(n_vreg, 0);
            std::vector<uint32_t> work;
            if (is_v(res)) { reach[res & 0xfffffu] = 1; work.push_back(res & 0xfffffu); }
            std::vector<uint32_t> uses(n_vreg, 0);
            if (is_v(res)) uses[res & 0xfffffu]++;
            while (!work.empty()) {
                const uint32_t v = work.back();
                work.pop_back();
                for (int k = 0; k < nops(node[v].op); k++) {
                    const uint32_t w = node[v].w[k];
                    if (!is_v(w)) continue;
                    uses[w & 0xfffffu]++;
                    if (!reach[w & 0xfffffu]) { reach[w & 0xfffffu] = 1; work.push_back(w & 0xfffffu); }
                }
            }
            for (uint32_t v = 0; v < n_vreg; v++) {
                if (!reach[v] || (node[v].op != Q_ADD && node[v].op != Q_SUB)) continue;
                const uint32_t wa = node[v].w[0], wb = node[v].w[1];
                if (!is_v(wa) || !is_v(wb)) continue;
                const uint32_t va = wa & 0xfffffu, vb = wb & 0xfffffu;
                if (va == vb || node[va].op != Q_MUL || node[vb].op != Q_MUL || uses[va] != 1 || uses[vb] != 1) continue;
                node[v] = Node{node[v].op == Q_ADD ? (uint32_t)Q_MUL2ADD : (uint32_t)Q_MUL2SUB,
                               {node[va].w[0], node[va].w[1], node[vb].w[0], node[vb].w[1]}};
            }
        }

        // Scheduling.  The reference computes every Calculation first and folds the value parts
        // afterwards (evaluation.rs:877-907), which keeps one value per gate alive until the fold.  The
        // raw list is a DAG in SSA form (virtual register v is defined by node[v]), so it is re-emitted
        // in depth-first post-order from the result: each term is computed right before it is folded
        // and the live width drops from O(#gates) to the depth of one expression (plus shared
        // subexpressions).  The larger operand subtree goes first (Sethi-Ullman).  Instructions the
        // result does not depend on are never visited (dead-code elimination).
        std::vector<uint32_t> weight(n_vreg, 1);
        auto wt = [&](uint32_t w) -> uint32_t { return is_v(w) ? weight[w & 0xfffffu] : 0u; };
        for (uint32_t v = 0; v < n_vreg; v++) {
            uint64_t t = 1;
            for (int k = 0; k < nops(node[v].op); k++) t += wt(node[v].w[k]);
            weight[v] = (uint32_t)std::min<uint64_t>(t, 1u << 30);
        }
        std::vector<char> state(n_vreg, 0);   // 0 unvisited, 1 operands pushed, 2 emitted
        std::vector<uint32_t> kept;           // virtual registers in emission order
        std::vector<uint32_t> stack;
        if (is_v(res)) stack.push_back(res & 0xfffffu);
        while (!stack.empty()) {
            const uint32_t v = stack.back();
            if (state[v] == 2) { stack.pop_back(); continue; }
            if (state[v] == 1) {
                kept.push_back(v);
                state[v] = 2;
                stack.pop_back();
                continue;
            }
            state[v] = 1;
            // operands by decreasing weight; pushed in reverse, so the heaviest is processed first
            uint32_t ord[4];
            const int n = nops(node[v].op);
            for (int k = 0; k < n; k++) ord[k] = node[v].w[k];
            std::stable_sort(ord, ord + n, [&](uint32_t x, uint32_t y) { return wt(x) > wt(y); });
            for (int k = n - 1; k >= 0; k--)
                if (is_v(ord[k]) && state[ord[k] & 0xfffffu] == 0) stack.push_back(ord[k] & 0xfffffu);
        }
        // last use of every live virtual register
        std::vector<uint32_t> last(n_vreg, 0);
        auto use = [&](uint32_t w, uint32_t at) { if (is_v(w)) last[w & 0xfffffu] = at; };
        for (uint32_t j = 0; j < kept.size(); j++)
            for (int k = 0; k < nops(node[kept[j]].op); k++) use(node[kept[j]].w[k], j);
        use(res, (uint32_t)kept.size());
        // slot allocation, two classes: shared memory, and (hybrid programs only) global memory for long-lived values.
        // allocate(glob, emit): runs the allocator over the schedule with the values flagged in `glob` in the global
        // class; returns the two widths and, with emit, appends the instructions (global slot g is written as
        // Q_GLOBAL_FLAG | g until the shared width is known, then renumbered to n_shared + g).
        constexpr uint32_t Q_GLOBAL_FLAG = 1u << 19;
        std::vector<uint32_t> slot_of(n_vreg, 0xffffffffu);
        auto remap = [&](uint32_t w) {
            return is_v(w) ? q_operand(QK_SLOT, slot_of[w & 0xfffffu], 0) : w;
        };
        auto allocate = [&](const std::vector<char>& glob, bool emit, uint32_t* n_sh, uint32_t* n_gl) {
            std::vector<uint32_t> lastc(last), free_sh, free_gl;
            std::fill(slot_of.begin(), slot_of.end(), 0xffffffffu);
            uint32_t ns = 0, ng = 0;
            auto release = [&](uint32_t w, uint32_t at) {
                if (!is_v(w)) return;
                const uint32_t v = w & 0xfffffu;
                if (lastc[v] == at && slot_of[v] != 0xffffffffu) {
                    ((slot_of[v] & Q_GLOBAL_FLAG) ? free_gl : free_sh).push_back(slot_of[v]);
                    lastc[v] = 0xffffffffu;   // released once, even when several operands name it
                }
            };
            for (uint32_t j = 0; j < kept.size(); j++) {
                const uint32_t dst = kept[j];
                const Node& nd = node[dst];
                const int n = nops(nd.op);
                uint32_t m[4] = {0, 0, 0, 0};
                for (int k = 0; k < n; k++) m[k] = remap(nd.w[k]);
                for (int k = 0; k < n; k++) release(nd.w[k], j);
                std::vector<uint32_t>& fr = glob[dst] ? free_gl : free_sh;
                uint32_t s;
                if (!fr.empty()) {
                    s = fr.back();
                    fr.pop_back();
                } else {
                    s = glob[dst] ? (Q_GLOBAL_FLAG | ng++) : ns++;
                }
                slot_of[dst] = s;
                if (!emit) continue;
                QInstr in;
                in.op_dst = nd.op | (s << 8);
                in.a = m[0];
                in.b = m[1];
                in.pad = m[2];
                p.instr.push_back(in);
                if (n == 4) {
                    QInstr ext;
                    ext.op_dst = Q_EXT;
                    ext.a = m[3];
                    ext.b = 0;
                    ext.pad = 0;
                    p.instr.push_back(ext);
                    p.n_mul += 2;
                    p.n_fused++;
                } else if (nd.op == Q_MUL) {
                    p.n_mul++;
                } else if (nd.op == Q_ADD || nd.op == Q_SUB || nd.op == Q_NEG) {
                    p.n_addsub++;
                }
            }
            *n_sh = ns;
            *n_gl = ng;
        };
        std::vector<char> glob(n_vreg, 0);
        uint32_t n_sh = 0, n_gl = 0;
        allocate(glob, false, &n_sh, &n_gl);
        // B2_Q_HYBRID=0 keeps every slot in shared memory (A/B runs); default: hybrid when the width costs occupancy
        const char* hybrid_env = getenv("B2_Q_HYBRID");     // read per program: tests lower the same circuit both ways
        const bool hybrid = !(hybrid_env && atoi(hybrid_env) == 0);
        // test knobs (host-only): a smaller target / shorter minimum live range push small programs through the
        // two-class allocator (tests/test_prover_fuzz.py)
        uint32_t shared_target = Q_SHARED_TARGET, hybrid_from = Q_HYBRID_FROM, min_live = Q_GLOBAL_MIN_LIVE;
        if (const char* e = getenv("B2_Q_SHARED_TARGET")) {
            shared_target = (uint32_t)std::max(1, atoi(e));
            hybrid_from = shared_target + 1;
        }
        if (const char* e = getenv("B2_Q_GLOBAL_MIN_LIVE")) min_live = (uint32_t)std::max(1, atoi(e));
        if (hybrid && n_sh >= hybrid_from) {
            // candidates by decreasing live length; values read again within Q_GLOBAL_MIN_LIVE instructions stay shared
            std::vector<uint32_t> def_pos(n_vreg, 0);
            for (uint32_t j = 0; j < kept.size(); j++) def_pos[kept[j]] = j;
            std::vector<uint32_t> cand;
            for (uint32_t v : kept)
                if (last[v] > def_pos[v] && last[v] - def_pos[v] >= min_live) cand.push_back(v);
            std::stable_sort(cand.begin(), cand.end(),
                             [&](uint32_t x, uint32_t y) { return last[x] - def_pos[x] > last[y] - def_pos[y]; });
            for (size_t i = 0; i < cand.size() && n_sh > shared_target; i++) {
                glob[cand[i]] = 1;
                allocate(glob, false, &n_sh, &n_gl);
            }
        }
        allocate(glob, true, &n_sh, &n_gl);
        if (n_sh >= Q_GLOBAL_FLAG || n_gl >= Q_GLOBAL_FLAG) { err = "too many live intermediates"; return false; }
        if (n_sh == 0) n_sh = 1;
        // global slot g -> n_sh + g, in destinations and operands
        auto fix_slot = [&](uint32_t s) { return (s & Q_GLOBAL_FLAG) ? n_sh + (s & (Q_GLOBAL_FLAG - 1u)) : s; };
        auto fix_operand = [&](uint32_t w) {
            return (w >> 28) == QK_SLOT ? q_operand(QK_SLOT, fix_slot(w & 0xfffffu), 0) : w;
        };
        for (QInstr& in : p.instr) {
            const uint32_t op = in.op_dst & 0xffu;
            if (op == Q_EXT) {
                in.a = fix_operand(in.a);
                continue;
            }
            in.op_dst = op | (fix_slot(in.op_dst >> 8) << 8);
            in.a = fix_operand(in.a);
            if (op != Q_NEG && op != Q_COPY) in.b = fix_operand(in.b);
            if (op >= Q_MUL2ADD) in.pad = fix_operand(in.pad);
        }
        for (uint32_t& so : slot_of)
            if (so != 0xffffffffu) so = fix_slot(so);
        p.result = remap(res);
        p.n_slots_shared = n_sh;
        p.n_slots = n_sh + n_gl;
        return true;
    }
};

int qprog_lookup(b2_handle_t h, QProgram** out) {
    std::lock_guard<std::mutex> lk(g_qprog_mu);
    auto it = g_qprog.find(h);
    if (it == g_qprog.end()) return fail(B2_ERR_HANDLE, "unknown quotient program handle %llu", (unsigned long long)h);
    *out = it->second;
    return B2_OK;
}

}  // namespace

extern "C" {

int b2_quotient_program_create(const b2_quotient_program_desc* desc, b2_handle_t* out) {
    if (!desc || !out) return fail(B2_ERR_ARG, "quotient_program_create: null pointer");
    if ((desc->n_calcs && !desc->calcs) || (desc->n_constants && !desc->constants) ||
        (desc->n_rotations && !desc->rotations))
        return fail(B2_ERR_ARG, "quotient_program_create: null table");
    QProgram* p = new QProgram();
    QLower low(*desc, *p);
    if (!low.run()) {
        delete p;
        return fail(B2_ERR_ARG, "quotient_program_create: %s", low.err.c_str());
    }
    p->rotations.assign(desc->rotations, desc->rotations + desc->n_rotations);
    p->constants.resize((size_t)desc->n_constants * 4);
    if (desc->n_constants) memcpy(p->constants.data(), desc->constants, (size_t)desc->n_constants * 32);
    p->n_fixed = desc->n_fixed;
    p->n_advice = desc->n_advice;
    p->n_instance = desc->n_instance;
    p->n_aux = desc->n_aux;
    p->n_challenges = desc->n_challenges;
    std::lock_guard<std::mutex> lk(g_qprog_mu);
    b2_handle_t h = g_next_handle++;
    g_qprog[h] = p;
    *out = h;
    return B2_OK;
}

int b2_quotient_program_free(b2_handle_t program) {
    QProgram* p = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_qprog_mu);
        auto it = g_qprog.find(program);
        if (it == g_qprog.end()) return fail(B2_ERR_HANDLE, "unknown quotient program handle");
        p = it->second;
        g_qprog.erase(it);
    }
    for (auto& kv : p->dev) {
        cudaSetDevice(kv.first);
        if (kv.second.prog) cudaFree(kv.second.prog);
        if (kv.second.constants) cudaFree(kv.second.constants);
    }
    delete p;
    return B2_OK;
}

int b2_quotient_program_info(b2_handle_t program, uint32_t* n_instr, uint32_t* n_slots, uint32_t* n_mul,
                             uint32_t* n_addsub) {
    QProgram* p;
    int rc = qprog_lookup(program, &p);
    if (rc) return rc;
    if (n_instr) *n_instr = (uint32_t)p->instr.size();
    if (n_slots) *n_slots = p->n_slots;
    if (n_mul) *n_mul = p->n_mul;
    if (n_addsub) *n_addsub = p->n_addsub;
    return B2_OK;
}

int b2_quotient_program_slot_classes(b2_handle_t program, uint32_t* n_shared, uint32_t* n_global) {
    QProgram* p;
    int rc = qprog_lookup(program, &p);
    if (rc) return rc;
    if (n_shared) *n_shared = p->n_slots_shared;
    if (n_global) *n_global = p->n_slots - p->n_slots_shared;
    return B2_OK;
}

int b2_quotient_program_dump(b2_handle_t program, uint32_t* instr_words, size_t instr_capacity, uint32_t* result_word,
                             uint32_t* derived_pairs, size_t derived_capacity, uint32_t* n_derived) {
    QProgram* p;
    int rc = qprog_lookup(program, &p);
    if (rc) return rc;
    if (instr_capacity < p->instr.size() * 4 || derived_capacity < p->derived.size() * 2)
        return fail(B2_ERR_ARG, "quotient_program_dump: buffers too small");
    for (size_t i = 0; i < p->instr.size(); i++) {
        instr_words[4 * i] = p->instr[i].op_dst;
        instr_words[4 * i + 1] = p->instr[i].a;
        instr_words[4 * i + 2] = p->instr[i].b;
        instr_words[4 * i + 3] = p->instr[i].pad;
    }
    for (size_t i = 0; i < p->derived.size(); i++) {
        derived_pairs[2 * i] = p->derived[i].first;
        derived_pairs[2 * i + 1] = p->derived[i].second;
    }
    if (result_word) *result_word = p->result;
    if (n_derived) *n_derived = (uint32_t)p->derived.size();
    return B2_OK;
}

int b2_quotient_eval(b2_handle_t program, const b2_quotient_args* args) {
    if (!args || !args->out) return fail(B2_ERR_ARG, "quotient_eval: null pointer");
    QProgram* p;
    int rc = qprog_lookup(program, &p);
    if (rc) return rc;
    if (args->log_rows < 1 || args->log_rows > 28) return fail(B2_ERR_ARG, "quotient_eval: log_rows out of [1, 28]");
    if ((p->n_fixed && !args->fixed) || (p->n_advice && !args->advice) || (p->n_instance && !args->instance) ||
        (p->n_aux && !args->aux) || (p->n_challenges && !args->challenges))
        return fail(B2_ERR_ARG, "quotient_eval: a column / challenge table the program uses is NULL");
    if (p->uses_x && (!args->x0 || !args->x_step)) return fail(B2_ERR_ARG, "quotient_eval: program reads COSET_X but x0 / x_step is NULL");
    if (args->scale && (args->scale_len == 0 || (args->scale_len & (args->scale_len - 1))))
        return fail(B2_ERR_ARG, "quotient_eval: scale_len must be a power of two");
    if (args->rot_scale == 0 || args->out_stride == 0) return fail(B2_ERR_ARG, "quotient_eval: rot_scale / out_stride must be >= 1");
    const unsigned long long rows = 1ull << args->log_rows;
    const unsigned long long row_begin = args->row_count ? args->row_begin : 0;
    const unsigned long long row_count = args->row_count ? args->row_count : rows;
    if (row_begin >= rows || row_count > rows - row_begin) return fail(B2_ERR_ARG, "quotient_eval: row range outside the domain");

    LaneLock ll;
    if ((rc = ll.acquire((cudaStream_t)args->stream))) return rc;
    Lane* ctx = ll.lane;
    cudaStream_t st = args->stream ? (cudaStream_t)args->stream : ctx->stream;
    if ((rc = ll.order_after_busy(st))) return rc;
    const int dev = ctx->dev->dev;

    // program + constants on this device (once)
    QProgram::Dev pd;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        auto it = p->dev.find(dev);
        if (it == p->dev.end()) {
            QProgram::Dev nd;
            CK(cudaMalloc(&nd.prog, std::max<size_t>(1, p->instr.size()) * sizeof(QInstr)));
            CK(cudaMalloc(&nd.constants, std::max<size_t>(1, p->constants.size() / 4) * 32));
            if (!p->instr.empty())
                CK(cudaMemcpy(nd.prog, p->instr.data(), p->instr.size() * sizeof(QInstr), cudaMemcpyHostToDevice));
            if (!p->constants.empty())
                CK(cudaMemcpy(nd.constants, p->constants.data(), p->constants.size() * 8, cudaMemcpyHostToDevice));
            // small pageable uploads may return before their DMA has landed, and the lane streams do not wait for
            // the legacy stream
            CK(cudaStreamSynchronize(cudaStreamLegacy));
            static bool attr_set[MAX_DEV] = {};
            if (!attr_set[dev]) {
                CK(cudaFuncSetAttribute(quotient_eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)Q_SMEM_LIMIT));
                CK(cudaFuncSetAttribute(quotient_eval_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)Q_SMEM_LIMIT));
                attr_set[dev] = true;
            }
            p->dev[dev] = nd;
            pd = nd;
        } else {
            pd = it->second;
        }
    }

    // per-call tables, packed into one staging block: column pointers | rot_off | challenges (+ powers) | scale
    const uint32_t n_cols = p->n_fixed + p->n_advice + p->n_instance + p->n_aux;
    const uint32_t n_ch = p->n_challenges + (uint32_t)p->derived.size();
    const size_t off_cols = 0;
    const size_t off_rot = off_cols + (size_t)std::max(1u, n_cols) * 8;
    const size_t off_ch = (off_rot + (size_t)std::max<size_t>(1, p->rotations.size()) * 4 + 31) & ~(size_t)31;
    const size_t off_scale = off_ch + (size_t)std::max(1u, n_ch) * 32;
    const size_t total = off_scale + (size_t)(args->scale ? args->scale_len : 1) * 32;
    std::vector<unsigned char> stage(total, 0);
    {
        const void** cols = reinterpret_cast<const void**>(stage.data() + off_cols);
        uint32_t c = 0;
        for (uint32_t i = 0; i < p->n_fixed; i++) cols[c++] = args->fixed[i];
        for (uint32_t i = 0; i < p->n_advice; i++) cols[c++] = args->advice[i];
        for (uint32_t i = 0; i < p->n_instance; i++) cols[c++] = args->instance[i];
        for (uint32_t i = 0; i < p->n_aux; i++) cols[c++] = args->aux[i];
        for (uint32_t i = 0; i < c; i++)
            if (!cols[i]) return fail(B2_ERR_ARG, "quotient_eval: column %u is NULL", i);
        uint32_t* rot = reinterpret_cast<uint32_t*>(stage.data() + off_rot);
        for (size_t i = 0; i < p->rotations.size(); i++) {
            // get_rotation_idx (evaluation.rs:40-42): (idx + rot * rot_scale).rem_euclid(size)
            long long r = ((long long)p->rotations[i] * (long long)args->rot_scale) % (long long)rows;
            if (r < 0) r += (long long)rows;
            rot[i] = (uint32_t)r;
        }
        uint64_t* ch = reinterpret_cast<uint64_t*>(stage.data() + off_ch);
        if (p->n_challenges) memcpy(ch, args->challenges, (size_t)p->n_challenges * 32);
        for (size_t i = 0; i < p->derived.size(); i++)
            hr_pow(ch + 4 * (p->n_challenges + i), ch + 4 * (size_t)p->derived[i].first, p->derived[i].second);
        if (args->scale) memcpy(stage.data() + off_scale, args->scale, (size_t)args->scale_len * 32);
    }
    const size_t xlo_n = (size_t)1 << Q_XLO_BITS;
    const size_t xhi_n = std::max<unsigned long long>(1, rows >> Q_XLO_BITS);
    if ((rc = ctx->qtab.reserve(total + (xlo_n + xhi_n) * 32 + 64))) return rc;
    char* dbase = ctx->qtab.as<char>();
    CK(cudaMemcpyAsync(dbase, stage.data(), total, cudaMemcpyHostToDevice, st));
    Fr* d_xlo = nullptr;
    Fr* d_xhi = nullptr;
    if (p->uses_x) {
        d_xlo = reinterpret_cast<Fr*>(dbase + ((total + 31) & ~(size_t)31));
        d_xhi = d_xlo + xlo_n;
        const Fr x0 = fr_from_bytes(args->x0), step = fr_from_bytes(args->x_step);
        LAUNCH(*ctx, ntt_pow_table_kernel, (unsigned)((xlo_n + 127) / 128), 128, 0, st, d_xlo, step, 1ull,
               (uint32_t)xlo_n, 0, step);
        LAUNCH(*ctx, ntt_pow_table_kernel, (unsigned)((xhi_n + 127) / 128), 128, 0, st, d_xhi, step,
               (unsigned long long)xlo_n, (uint32_t)xhi_n, 1, x0);
    }

    QArgs a;
    memset(&a, 0, sizeof a);
    a.prog = reinterpret_cast<const uint4*>(pd.prog);
    a.n_instr = (uint32_t)p->instr.size();
    a.result = p->result;
    a.constants = pd.constants;
    a.challenges = reinterpret_cast<const Fr*>(dbase + off_ch);
    a.columns = reinterpret_cast<const uint4* const*>(dbase + off_cols);
    a.rot_off = reinterpret_cast<const uint32_t*>(dbase + off_rot);
    a.x_lo = d_xlo;
    a.x_hi = d_xhi;
    a.scale = args->scale ? reinterpret_cast<const Fr*>(dbase + off_scale) : nullptr;
    a.scale_mask = args->scale ? args->scale_len - 1 : 0;
    a.n_slots = p->n_slots;
    a.rows = rows;
    a.out = reinterpret_cast<uint4*>(args->out);
    a.out_stride = args->out_stride;
    a.out_offset = args->out_offset;
    a.row_begin = row_begin;
    a.row_count = row_count;
    a.n_slots_shared = p->n_slots_shared;
    const size_t slot_bytes = (size_t)Q_THREADS * 32;
    const size_t smem = (size_t)p->n_slots_shared * slot_bytes;
    const unsigned long long blocks_all = (row_count + Q_THREADS - 1) / Q_THREADS;
    const char* force_spill = getenv("B2_Q_FORCE_SPILL");   // tests: exercise the global-slot variant
    CK(cudaEventRecord(ctx->ev[12], st));
    if (smem > Q_SMEM_LIMIT || (force_spill && atoi(force_spill) == 1)) {
        // live width beyond shared memory: every slot in a global scratch (L2-resident), grid-stride over rows
        const unsigned blocks = (unsigned)std::min<unsigned long long>(blocks_all, (unsigned long long)ctx->sms * 8);
        if ((rc = ctx->qspill.reserve((size_t)blocks * p->n_slots * slot_bytes))) return rc;
        a.slot_spill = ctx->qspill.as<uint4>();
        LAUNCH(*ctx, quotient_eval_kernel<false>, blocks, Q_THREADS, 0, st, a);
    } else if (p->n_slots == p->n_slots_shared) {
        LAUNCH(*ctx, quotient_eval_kernel<true>, (unsigned)blocks_all, Q_THREADS, smem, st, a);
    } else {
        // two slot classes: the long-lived values of this program sit in a per-CTA global scratch, so the grid is one
        // wave of resident CTAs striding over the rows (the scratch is CTAs x global slots x 4 KB, L2-resident)
        auto* hybrid_kernel = quotient_eval_kernel<true, true>;
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hybrid_kernel, Q_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        const unsigned blocks =
            (unsigned)std::min<unsigned long long>(blocks_all, (unsigned long long)ctx->sms * (unsigned)per_sm);
        const size_t n_global = p->n_slots - p->n_slots_shared;
        if ((rc = ctx->qspill.reserve((size_t)blocks * n_global * slot_bytes))) return rc;
        a.slot_spill = ctx->qspill.as<uint4>();
        LAUNCH(*ctx, hybrid_kernel, blocks, Q_THREADS, smem, st, a);
    }
    CK(cudaEventRecord(ctx->ev[13], st));
    if (args->stream) return ll.mark_busy(st);
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev[12], ctx->ev[13]));
    ctx->last_kernel_ms = ms;
    g_last.kernel_ms = ms;
    g_last.total_ms = ms;
    return B2_OK;
}

}  // extern "C"
