// radix.cpp — radix-integer operators as sequences of batched PBS levels (see radix.h).
#include "radix.h"

#include <limits.h>
#include <stdlib.h>

#include <algorithm>
#include <functional>

namespace fsc {

// =======================================================================================
// symbolic block arithmetic
// =======================================================================================
bool noise_in_variance_units() {
    // default since round 2: variance units (measured on the GPU at sum c^2 = 25: tests/test_gpu_pbs.py::test_noise_budget_in_variance_units);
    // FSC_RADIX_NOISE=linear restores the sum |c| <= 5 rule tfhe's NoiseLevel applies
    static const bool v = [] { const char* e = getenv("FSC_RADIX_NOISE"); return !(e && e[0] == 'l'); }();
    return v;
}

static void merge_term(std::vector<std::pair<SlotP, int32_t>>& t, const SlotP& s, int32_t c) {
    if (c == 0) return;
    for (auto& e : t)
        if (e.first.get() == s.get()) { e.second += c; return; }
    t.emplace_back(s, c);
}
static void drop_zero_terms(Block& b) {
    b.terms.erase(std::remove_if(b.terms.begin(), b.terms.end(), [](const std::pair<SlotP, int32_t>& e) { return e.second == 0; }),
                  b.terms.end());
}

// variance of a linear combination of independent slots, from the merged coefficients
static int32_t variance_of(const Block& b) {
    int64_t v = 0;
    for (const auto& e : b.terms) v += (int64_t)e.second * e.second * e.first->nv;
    return (int32_t)std::min<int64_t>(v, INT32_MAX);
}

Block operator+(const Block& a, const Block& b) {
    Block r = a;
    for (const auto& e : b.terms) merge_term(r.terms, e.first, e.second);
    drop_zero_terms(r);
    r.cst += b.cst; r.deg += b.deg; r.nl += b.nl;
    r.nv = variance_of(r);
    return r;
}
Block operator*(const Block& a, int c) {
    if (c < 0) throw RadixError("negative block scaling");
    Block r = a;
    for (auto& e : r.terms) e.second *= c;
    drop_zero_terms(r);
    r.cst *= c; r.deg *= c; r.nl *= c;
    r.nv = variance_of(r);
    return r;
}
Block complement(const Block& a, int top) {
    if (a.deg > top) throw RadixError("complement: block may exceed its bound");
    Block r = a;
    for (auto& e : r.terms) e.second = -e.second;
    r.cst = top - a.cst; r.deg = top;
    return r;
}
Block add_const(const Block& a, int c) {
    Block r = a;
    r.cst += c; r.deg += c;
    return r;
}

std::vector<uint8_t> digits_from_u64(uint64_t v, int n_blocks) {
    std::vector<uint8_t> d(n_blocks, 0);
    for (int i = 0; i < n_blocks && i < 32; ++i) d[i] = (v >> (2 * i)) & 3;
    return d;
}
std::vector<uint8_t> digits_from_bytes_le(const uint8_t* bytes, size_t n_bytes, int n_blocks) {
    std::vector<uint8_t> d(n_blocks, 0);
    for (int i = 0; i < n_blocks; ++i) {
        const size_t byte = (size_t)i / 4;
        if (byte < n_bytes) d[i] = (bytes[byte] >> (2 * (i % 4))) & 3;
    }
    return d;
}

// =======================================================================================
// LUT helpers
// =======================================================================================
static LutTable make_lut(const std::function<int(int)>& f) {
    LutTable t;
    for (int v = 0; v < kSpace; ++v) t[v] = (uint8_t)(f(v) & (kSpace - 1));
    return t;
}
static LutTable make_bilut(const std::function<int(int, int)>& f) {      // input = hi * 4 + lo
    return make_lut([&](int v) { return f(v >> 2, v & 3); });
}
static LutTable make_sel_lut(const std::function<int(int, int)>& f) {    // input = x * 2 + bit
    return make_lut([&](int v) { return f(v >> 1, v & 1); });
}
static const LutTable& lut_identity() { static LutTable t = make_lut([](int v) { return v; }); return t; }
static const LutTable& lut_msg() { static LutTable t = make_lut([](int v) { return v & 3; }); return t; }
static const LutTable& lut_carry() { static LutTable t = make_lut([](int v) { return v >> 2; }); return t; }

// =======================================================================================
// level machinery
// =======================================================================================
std::vector<Block> Evaluator::level(const std::vector<Req>& reqs) {
    std::vector<Block> out(reqs.size());
    std::vector<LevelReq> dev;
    dev.reserve(reqs.size());
    for (size_t i = 0; i < reqs.size(); ++i) {
        const Req& r = reqs[i];
        if (r.in.deg >= kSpace) throw RadixError("level: block value may overflow the 4-bit plaintext space");
        int deg = 0;
        for (int v = 0; v <= std::min(r.in.deg, kSpace - 1); ++v) deg = std::max<int>(deg, r.lut[v]);
        if (r.in.trivial()) {
            if (r.in.cst < 0 || r.in.cst >= kSpace) throw RadixError("level: trivial block out of range");
            out[i] = Block::constant(r.lut[r.in.cst]);
            continue;
        }
        if (!within_noise_budget(r.in)) throw RadixError("level: noise budget exceeded before a bootstrap");
        LevelReq q;
        for (const auto& e : r.in.terms) q.terms.emplace_back(e.first->idx, e.second);
        q.cst = r.in.cst;
        q.lut = be_->lut_id(r.lut);
        q.dst = be_->alloc_slot();
        out[i] = Block::from_slot(std::make_shared<SlotRef>(be_, q.dst), deg, 1);
        dev.push_back(std::move(q));
    }
    if (!dev.empty()) {
        be_->run_level(dev);
        be_->pbs_count += dev.size();
        be_->level_count += 1;
    }
    return out;
}

void Evaluator::materialize(Radix& r) {
    std::vector<LinReq> lin;
    for (auto& b : r) {
        const bool single = b.terms.size() == 1 && b.terms[0].second == 1 && b.cst == 0;
        if (single) continue;
        LinReq q;
        for (const auto& e : b.terms) q.terms.emplace_back(e.first->idx, e.second);
        q.cst = b.cst;
        q.dst = be_->alloc_slot();
        Block nb = Block::from_slot(std::make_shared<SlotRef>(be_, q.dst), b.deg, b.nl);
        nb.terms[0].first->nv = std::max(b.nv, 1);      // the materialised slot carries the combination's variance
        nb.nv = b.nv;
        lin.push_back(std::move(q));
        // keep the sources alive until run_linear has been enqueued
        b.terms.swap(nb.terms);
        std::swap(b.cst, nb.cst);
        keep_.push_back(std::move(nb));
    }
    if (!lin.empty()) be_->run_linear(lin);
    keep_.clear();
}

static bool is_clean(const Block& b) {
    return b.deg <= 3 && (b.trivial() || (b.nl <= 1 && b.terms.size() == 1 && b.terms[0].second == 1 && b.cst == 0));
}

// identity bootstrap of every block that is not a single fresh slot
void Evaluator::clean(Radix& r) {
    std::vector<Req> reqs;
    std::vector<size_t> where;
    for (size_t i = 0; i < r.size(); ++i)
        if (!is_clean(r[i])) {
            if (r[i].deg > 3) throw RadixError("clean: block carries are not empty");
            reqs.push_back({r[i], lut_identity()});
            where.push_back(i);
        }
    if (reqs.empty()) return;
    std::vector<Block> o = level(reqs);
    for (size_t k = 0; k < where.size(); ++k) r[where[k]] = o[k];
}

Radix Evaluator::trivial_big(const std::vector<uint8_t>& blocks) {
    Radix r;
    for (uint8_t d : blocks) r.push_back(Block::constant(d & 3));
    return r;
}

Radix Evaluator::cast(const Radix& a, int n_blocks) {
    Radix r(a.begin(), a.begin() + std::min<size_t>(a.size(), n_blocks));
    while ((int)r.size() < n_blocks) r.push_back(Block::constant(0));
    return r;
}

// =======================================================================================
// carry propagation (parallel prefix over generate / propagate states)
// =======================================================================================
// Number of ranks a level of this evaluator may be cut over (1 without level sharding).
int Evaluator::scan_world() const {
    const Exchange& x = be_->exchange;
    return x.enabled() ? x.world : 1;
}

// Radix-3 Hillis-Steele scan.  States are re-encoded as the digits of a binary adder, e = 0 (no carry), 1 (propagates),
// 2 (generates): for three consecutive segments a (most significant), b, c the sum S = 4 e_a + 2 e_b + e_c <= 14 carries
// out of bit 2 exactly when the joined segment generates (S >= 8) and is all ones exactly when it propagates (S == 7),
// so ONE lookup joins three segments.  The noise budget (sum of coefficients <= 5 on fresh blocks) forbids the factor 4
// as a linear coefficient; every position therefore keeps two fresh encodings, Y = 2 e (used with coefficient 2 in its
// own next join and with coefficient 1 as the middle segment of another) and Z = e (least significant segment, and the
// final carry-in): S = 2 Y_a + Y_b + Z_c, noise level 4.  Twice the bootstraps of the radix-2 scan per level, log3 n
// levels instead of log2 n; propagate() takes this path only while a level stays at one ciphertext per SM.
Radix Evaluator::propagate_radix3(const std::vector<Block>& msg, std::vector<Block>& Y, std::vector<Block>& Z, Block* carry_out) {
    // msg: clean messages of the n blocks; Y = 2 e, Z = e of the single-block states (propagate()'s first level)
    // Variance-unit bookkeeping (the default): 4 e_a + 2 e_b + e_c on fresh blocks costs sum c^2 = 21 <= 25, so ONE encoding per
    // position is enough (Y is the linear 2 Z, no second bootstrap): the radix-3 scan then costs one bootstrap per position and
    // level like the radix-2 scan and is used at every width.  Linear bookkeeping (FSC_RADIX_NOISE=linear) forbids the factor 4 and
    // keeps the two fresh encodings.
    const bool single = noise_in_variance_units();
    const int n = (int)msg.size();
    const int n_state = (int)Z.size();
    auto join = [](int S) { return S >= 8 ? 2 : (S == 7 ? 1 : 0); };
    static const LutTable l_j1 = make_lut([join](int S) { return S <= 14 ? join(S) : 0; });
    static const LutTable l_j2 = make_lut([join](int S) { return S <= 14 ? 2 * join(S) : 0; });
    static const LutTable l_final = make_bilut([](int e, int m) { return (m + (e == 2)) & 3; });
    static const LutTable l_inc = make_lut([](int v) { return (v + 1) & 3; });
    static const LutTable l_is2 = make_lut([](int v) { return v == 2; });
    if (single) {
        Y.resize(n_state);
        for (int i = 0; i < n_state; ++i) Y[i] = Z[i] * 2;
    }
    for (int d = 1; d < n_state; d *= 3) {
        const bool last = 3 * (long)d >= n_state;
        std::vector<Req> todo;
        std::vector<std::pair<int, int>> where;
        std::vector<Block> nY = Y, nZ = Z;
        for (int i = d; i < n_state; ++i) {
            Block S = Y[i] * 2 + Y[i - d];
            if (i - 2 * d >= 0) S = S + Z[i - 2 * d];
            // Z: least significant segment of a later join, or the final carry-in; Y: any other role in the next level
            todo.push_back({S, l_j1}); where.emplace_back(1, i);
            if (!last && !single) { todo.push_back({S, l_j2}); where.emplace_back(0, i); }
        }
        std::vector<Block> o = level(todo);
        for (size_t k = 0; k < o.size(); ++k) (where[k].first ? nZ : nY)[where[k].second] = o[k];
        if (single) for (int i = d; i < n_state; ++i) nY[i] = nZ[i] * 2;
        Y.swap(nY); Z.swap(nZ);
    }
    // final level: add the incoming carry (prefix of the blocks below) to every message
    Radix res(n);
    std::vector<Req> todo;
    std::vector<int> where;
    if (n > 0) res[0] = msg[0];
    for (int i = 1; i < n; ++i) {
        const Block& e = Z[i - 1];
        if (e.trivial()) {
            if (e.cst != 2) { res[i] = msg[i]; continue; }
            todo.push_back({msg[i], l_inc});
        } else {
            todo.push_back({e * 4 + msg[i], l_final});
        }
        where.push_back(i);
    }
    if (carry_out) { todo.push_back({Z[n - 1], l_is2}); where.push_back(-1); }
    std::vector<Block> o = level(todo);
    for (size_t k = 0; k < o.size(); ++k) {
        if (where[k] < 0) *carry_out = o[k];
        else res[where[k]] = o[k];
    }
    return res;
}

// state: 0 = no carry out, 1 = generates a carry, 2 = propagates an incoming carry
Radix Evaluator::propagate(const std::vector<Block>& sums, Block* carry_out) {
    const int n = (int)sums.size();
    static const LutTable l_state = make_lut([](int v) { return v >= 4 ? 1 : (v == 3 ? 2 : 0); });
    static const LutTable l_comb = make_bilut([](int hi, int lo) { return hi == 2 ? lo : hi; });
    static const LutTable l_final = make_bilut([](int st, int m) { return (m + (st == 1)) & 3; });
    static const LutTable l_inc = make_lut([](int v) { return (v + 1) & 3; });
    static const LutTable l_is1 = make_lut([](int v) { return v == 1; });
    for (int i = 0; i < n; ++i) {
        // without a carry out the top block yields a message only: anything that fits the lookup is fine there
        const int limit = (i == n - 1 && !carry_out) ? (i == 0 ? 15 : 14) : (i == 0 ? 7 : 6);
        if (sums[i].deg > limit) throw RadixError("propagate: block sum too large for a single carry bit");
    }

    // level 1: message and state of every block
    const int n_state = carry_out ? n : n - 1;
    // blocks that are already clean and cannot carry need no bootstrap at all
    // Radix-3 scan for levels that stay within one ciphertext per SM even at two bootstraps per block (2 + log3 n levels
    // instead of 2 + log2 n): the 16- and 32-block operators on one GPU, everything up to 512 blocks on eight.  A narrow
    // level costs the same whatever its width, so depth is what counts there.
    // With variance-unit bookkeeping the radix-3 scan needs one bootstrap per position (propagate_radix3) and is used at every width.
    const bool single = noise_in_variance_units();
    const bool radix3 = n_state > 2 && (single || 2 * (size_t)n_state <= kBlocksPerGpuLevel * (size_t)scan_world());
    static const LutTable l_s1 = make_lut([](int v) { return v >= 4 ? 2 : (v == 3 ? 1 : 0); });       // block sum -> e
    static const LutTable l_s2 = make_lut([](int v) { return v >= 4 ? 4 : (v == 3 ? 2 : 0); });       // block sum -> 2 e
    std::vector<Block> msg(n), st(std::max(n_state, 0)), st2(radix3 && !single ? n_state : 0);
    {
        std::vector<Req> todo;
        std::vector<std::pair<int, int>> where;     // (kind, index)
        for (int i = 0; i < n; ++i) {
            if (is_clean(sums[i])) msg[i] = sums[i];
            else { todo.push_back({sums[i], lut_msg()}); where.emplace_back(0, i); }
        }
        for (int i = 0; i < n_state; ++i) {
            if (sums[i].deg <= 2) { st[i] = Block::constant(0); if (radix3 && !single) st2[i] = Block::constant(0); }
            else if (!radix3) { todo.push_back({sums[i], l_state}); where.emplace_back(1, i); }
            else {
                todo.push_back({sums[i], l_s1}); where.emplace_back(1, i);
                if (!single) { todo.push_back({sums[i], l_s2}); where.emplace_back(2, i); }
            }
        }
        std::vector<Block> o = level(todo);
        for (size_t k = 0; k < o.size(); ++k)
            (where[k].first == 0 ? msg[where[k].second] : where[k].first == 1 ? st[where[k].second] : st2[where[k].second]) = o[k];
    }
    if (radix3) return propagate_radix3(msg, st2, st, carry_out);
    // Hillis-Steele inclusive scan, most significant state dominates unless it propagates
    for (int d = 1; d < n_state; d <<= 1) {
        std::vector<Req> todo;
        std::vector<int> where;
        std::vector<Block> nxt = st;
        for (int i = d; i < n_state; ++i) {
            const Block& hi = st[i];
            const Block& lo = st[i - d];
            if (hi.trivial()) { nxt[i] = (hi.cst == 2) ? lo : hi; continue; }
            if (lo.trivial()) {
                if (lo.cst == 2) { nxt[i] = hi; continue; }
                const int lc = lo.cst;
                todo.push_back({hi, make_lut([lc](int v) { return v == 2 ? lc : v; })});
            } else {
                todo.push_back({hi * 4 + lo, l_comb});
            }
            where.push_back(i);
        }
        std::vector<Block> o = level(todo);
        for (size_t k = 0; k < o.size(); ++k) nxt[where[k]] = o[k];
        st.swap(nxt);
    }
    // final level: add the incoming carry to every message
    Radix res(n);
    {
        std::vector<Req> todo;
        std::vector<int> where;
        if (n > 0) res[0] = msg[0];
        for (int i = 1; i < n; ++i) {
            const Block& s = st[i - 1];
            if (s.trivial()) {
                if (s.cst != 1) { res[i] = msg[i]; continue; }
                todo.push_back({msg[i], l_inc});
            } else {
                todo.push_back({s * 4 + msg[i], l_final});
            }
            where.push_back(i);
        }
        if (carry_out) {
            todo.push_back({st[n - 1], l_is1});
            where.push_back(-1);
        }
        std::vector<Block> o = level(todo);
        for (size_t k = 0; k < o.size(); ++k) {
            if (where[k] < 0) *carry_out = o[k];
            else res[where[k]] = o[k];
        }
    }
    return res;
}

Radix Evaluator::add(const Radix& a, const Radix& b, Block* carry_out) {
    const size_t n = std::max(a.size(), b.size());
    std::vector<Block> sums(n);
    for (size_t i = 0; i < n; ++i) {
        const Block x = i < a.size() ? a[i] : Block::constant(0);
        const Block y = i < b.size() ? b[i] : Block::constant(0);
        sums[i] = x + y;
    }
    return propagate(sums, carry_out);
}

Radix Evaluator::sub(const Radix& a, const Radix& b, Block* not_borrow) {
    const size_t n = std::max(a.size(), b.size());
    std::vector<Block> sums(n);
    for (size_t i = 0; i < n; ++i) {
        const Block x = i < a.size() ? a[i] : Block::constant(0);
        const Block y = i < b.size() ? b[i] : Block::constant(0);
        sums[i] = x + complement(y, 3);
    }
    if (n) sums[0] = add_const(sums[0], 1);
    return propagate(sums, not_borrow);
}

Radix Evaluator::scalar_add(const Radix& a, const std::vector<uint8_t>& c) {
    std::vector<Block> sums(a.size());
    for (size_t i = 0; i < a.size(); ++i) sums[i] = add_const(a[i], i < c.size() ? (c[i] & 3) : 0);
    return propagate(sums, nullptr);
}

// =======================================================================================
// column sums (carry-save reduction, then one carry propagation)
// =======================================================================================
Radix Evaluator::sum_columns(std::vector<std::vector<Block>>& cols) {
    const int n = (int)cols.size();
    // A column is ready for the carry propagation when its terms add up to at most 6 (7 in column 0, which receives no
    // carry) within the noise budget of one lookup - however many terms that is: the propagation's first level reads the
    // sum directly.  (Asking for at most two terms costs whole levels of one or two bootstraps at the end of narrow
    // products.)
    auto needs_round = [&]() {
        for (int c = 0; c < n; ++c) {
            int deg = 0;
            Block acc = Block::constant(0);
            for (const auto& b : cols[c]) { deg += b.deg; acc.nl += b.nl; acc.nv += b.nv; }
            // the most significant column wraps: no state is derived from it, only its message (any sum below 15 + carry-in)
            const int limit = c == n - 1 ? (c == 0 ? 15 : 14) : (c == 0 ? 7 : 6);
            if (deg > limit || !within_noise_budget(acc)) return true;
        }
        return false;
    };
    while (needs_round()) {
        // Columns with more than two terms are reduced.  A two-term column that is about to receive a carry from a
        // reduced neighbour is reduced in the same round: otherwise it would hold three terms in the next round,
        // pass a carry on, and so forth - one extra PBS level per column of the ripple.
        std::vector<char> mark(n, 0);
        for (int c = 0; c < n; ++c) {
            int deg = 0;
            for (const auto& b : cols[c]) deg += b.deg;
            const bool sends_carry = c > 0 && mark[c - 1] == 2;
            if (cols[c].size() > 2 || (cols[c].size() == 2 && sends_carry)) mark[c] = deg >= 4 ? 2 : 1;   // 2: may emit a carry
        }
        std::vector<std::vector<Block>> nxt(n);
        std::vector<Req> reqs;
        std::vector<int> dest;
        for (int c = 0; c < n; ++c) {
            auto& col = cols[c];
            if (!mark[c]) { for (auto& b : col) nxt[c].push_back(b); continue; }
            // trivial constants first so that they ride along in a chunk for free
            std::stable_sort(col.begin(), col.end(), [](const Block& x, const Block& y) { return x.trivial() > y.trivial(); });
            size_t i = 0;
            while (i < col.size()) {
                Block s = col[i];
                size_t j = i + 1;
                while (j < col.size() && s.deg + col[j].deg < kSpace && within_noise_budget(s + col[j])) { s = s + col[j]; ++j; }
                if (j == i + 1) { nxt[c].push_back(col[i]); }
                else {
                    reqs.push_back({s, lut_msg()}); dest.push_back(c);
                    if (s.deg >= 4 && c + 1 < n) { reqs.push_back({s, lut_carry()}); dest.push_back(c + 1); }
                }
                i = j;
            }
        }
        std::vector<Block> o = level(reqs);
        for (size_t k = 0; k < o.size(); ++k)
            if (!(o[k].trivial() && o[k].cst == 0)) nxt[dest[k]].push_back(o[k]);
        cols.swap(nxt);
    }
    std::vector<Block> sums(n);
    bool all_clean = true;
    for (int c = 0; c < n; ++c) {
        Block s = Block::constant(0);
        for (const auto& b : cols[c]) s = s + b;
        sums[c] = s;
        all_clean = all_clean && is_clean(s);
    }
    if (all_clean) return sums;
    return propagate(sums, nullptr);
}

Radix Evaluator::sum(const std::vector<Radix>& operands, int n_blocks) {
    std::vector<std::vector<Block>> cols(n_blocks);
    for (const auto& r : operands)
        for (int i = 0; i < n_blocks && i < (int)r.size(); ++i)
            if (!(r[i].trivial() && r[i].cst == 0)) cols[i].push_back(r[i]);
    return sum_columns(cols);
}

// =======================================================================================
// multiplication
// =======================================================================================
Radix Evaluator::mul(const Radix& a_in, const Radix& b_in, int out_blocks) {
    return mul_add(a_in, b_in, nullptr, out_blocks);
}

// a * b + addend with the addend's blocks riding in the product's column sum: one carry propagation for both
// (the reference's `k + (e * d)`, src/schnorr.rs:274)
Radix Evaluator::mul_add(const Radix& a_in, const Radix& b_in, const Radix* addend, int out_blocks) {
    Radix a = a_in, b = b_in;
    clean(a); clean(b);
    const int n = out_blocks < 0 ? (int)a.size() : out_blocks;
    static const LutTable l_lo = make_bilut([](int x, int y) { return (x * y) & 3; });
    static const LutTable l_hi = make_bilut([](int x, int y) { return (x * y) >> 2; });
    std::vector<Req> reqs;
    std::vector<int> dest;
    std::vector<std::vector<Block>> cols(n);
    for (int i = 0; i < (int)a.size(); ++i) {
        if (a[i].trivial() && a[i].cst == 0) continue;
        for (int j = 0; j < (int)b.size() && i + j < n; ++j) {
            if (b[j].trivial() && b[j].cst == 0) continue;
            const bool hi = i + j + 1 < n;
            if (a[i].trivial() && b[j].trivial()) {
                const int p = a[i].cst * b[j].cst;
                if (p & 3) cols[i + j].push_back(Block::constant(p & 3));
                if (hi && (p >> 2)) cols[i + j + 1].push_back(Block::constant(p >> 2));
            } else if (a[i].trivial() || b[j].trivial()) {
                const int c = a[i].trivial() ? a[i].cst : b[j].cst;
                const Block& x = a[i].trivial() ? b[j] : a[i];
                if (c == 1) { cols[i + j].push_back(x); continue; }
                reqs.push_back({x, make_lut([c](int v) { return (v * c) & 3; })}); dest.push_back(i + j);
                if (hi) { reqs.push_back({x, make_lut([c](int v) { return ((v & 3) * c) >> 2; })}); dest.push_back(i + j + 1); }
            } else {
                const Block in = a[i] * 4 + b[j];
                reqs.push_back({in, l_lo}); dest.push_back(i + j);
                if (hi) { reqs.push_back({in, l_hi}); dest.push_back(i + j + 1); }
            }
        }
    }
    std::vector<Block> o = level(reqs);
    for (size_t k = 0; k < o.size(); ++k) cols[dest[k]].push_back(o[k]);
    if (addend) {
        Radix ad = *addend;
        clean(ad);
        for (int i = 0; i < (int)ad.size() && i < n; ++i)
            if (!(ad[i].trivial() && ad[i].cst == 0)) cols[i].push_back(ad[i]);
    }
    return sum_columns(cols);
}

Radix Evaluator::scalar_mul(const Radix& a_in, const std::vector<uint8_t>& c, int out_blocks) {
    return scalar_mul_add(a_in, c, nullptr, out_blocks);
}

// a * c + addend in one column sum and one carry propagation (addend: clean blocks, may be null)
Radix Evaluator::scalar_mul_add(const Radix& a_in, const std::vector<uint8_t>& c, const Radix* addend, int out_blocks) {
    Radix a = a_in;
    clean(a);
    const int n = out_blocks < 0 ? (int)a.size() : out_blocks;
    bool need[4] = {false, false, false, false};
    for (uint8_t d : c) need[d & 3] = true;
    std::vector<Block> lo[4], hi[4];
    std::vector<Req> reqs;
    std::vector<std::pair<int, int>> where;       // (d * 2 + is_hi, i)
    for (int d = 2; d <= 3; ++d) {
        if (!need[d]) continue;
        lo[d].resize(a.size()); hi[d].resize(a.size());
        for (int i = 0; i < (int)a.size(); ++i) {
            reqs.push_back({a[i], make_lut([d](int v) { return ((v & 3) * d) & 3; })}); where.emplace_back(d * 2, i);
            reqs.push_back({a[i], make_lut([d](int v) { return ((v & 3) * d) >> 2; })}); where.emplace_back(d * 2 + 1, i);
        }
    }
    std::vector<Block> o = level(reqs);
    for (size_t k = 0; k < o.size(); ++k) ((where[k].first & 1) ? hi : lo)[where[k].first >> 1][where[k].second] = o[k];
    std::vector<std::vector<Block>> cols(n);
    auto push = [&](int col, const Block& b) {
        if (col < n && !(b.trivial() && b.cst == 0)) cols[col].push_back(b);
    };
    for (int j = 0; j < (int)c.size() && j < n; ++j) {
        const int d = c[j] & 3;
        if (!d) continue;
        for (int i = 0; i < (int)a.size() && i + j < n; ++i) {
            if (d == 1) push(i + j, a[i]);
            else { push(i + j, lo[d][i]); push(i + j + 1, hi[d][i]); }
        }
    }
    if (addend) {
        Radix ad = *addend;
        clean(ad);
        for (int i = 0; i < (int)ad.size(); ++i) push(i, ad[i]);
    }
    return sum_columns(cols);
}

// =======================================================================================
// shifts, masks, bitwise
// =======================================================================================
Radix Evaluator::scalar_shr(const Radix& a_in, unsigned bits) {
    const int n = (int)a_in.size();
    const int q = bits / 2, r = bits % 2;
    Radix out(n, Block::constant(0));
    if (q >= n) return out;
    if (r == 0) {
        for (int i = 0; i + q < n; ++i) out[i] = a_in[i + q];
        return out;
    }
    Radix a = a_in;
    clean(a);
    static const LutTable l_two = make_bilut([](int hi, int lo) { return (lo >> 1) | ((hi & 1) << 1); });
    static const LutTable l_one = make_lut([](int v) { return (v & 3) >> 1; });
    std::vector<Req> reqs;
    for (int i = 0; i + q < n; ++i) {
        if (i + q + 1 < n) reqs.push_back({a[i + q + 1] * 4 + a[i + q], l_two});
        else reqs.push_back({a[i + q], l_one});
    }
    std::vector<Block> o = level(reqs);
    for (size_t k = 0; k < o.size(); ++k) out[k] = o[k];
    return out;
}

Radix Evaluator::scalar_shl(const Radix& a_in, unsigned bits) {
    const int n = (int)a_in.size();
    const int q = bits / 2, r = bits % 2;
    Radix out(n, Block::constant(0));
    if (q >= n) return out;
    if (r == 0) {
        for (int i = q; i < n; ++i) out[i] = a_in[i - q];
        return out;
    }
    Radix a = a_in;
    clean(a);
    static const LutTable l_two = make_bilut([](int cur, int below) { return ((cur << 1) & 3) | (below >> 1); });
    static const LutTable l_one = make_lut([](int v) { return (v << 1) & 3; });
    std::vector<Req> reqs;
    for (int i = q; i < n; ++i) {
        if (i - q - 1 >= 0) reqs.push_back({a[i - q] * 4 + a[i - q - 1], l_two});
        else reqs.push_back({a[i - q], l_one});
    }
    std::vector<Block> o = level(reqs);
    for (size_t k = 0; k < o.size(); ++k) out[q + k] = o[k];
    return out;
}

Radix Evaluator::scalar_and(const Radix& a, const std::vector<uint8_t>& mask) {
    const int n = (int)a.size();
    Radix out(n, Block::constant(0));
    std::vector<Req> reqs;
    std::vector<int> where;
    for (int i = 0; i < n; ++i) {
        const int m = i < (int)mask.size() ? (mask[i] & 3) : 0;
        if (m == 0) continue;
        if (m == 3 && a[i].deg <= 3) { out[i] = a[i]; continue; }
        reqs.push_back({a[i], make_lut([m](int v) { return v & m; })});
        where.push_back(i);
    }
    std::vector<Block> o = level(reqs);
    for (size_t k = 0; k < o.size(); ++k) out[where[k]] = o[k];
    return out;
}

Radix Evaluator::bitop(const Radix& a_in, const Radix& b_in, int op) {
    Radix a = a_in, b = b_in;
    clean(a); clean(b);
    const int n = (int)std::max(a.size(), b.size());
    a = cast(a, n); b = cast(b, n);
    const LutTable l = make_bilut([op](int x, int y) { return op == 0 ? (x & y) : op == 1 ? (x | y) : (x ^ y); });
    std::vector<Req> reqs;
    for (int i = 0; i < n; ++i) reqs.push_back({a[i] * 4 + b[i], l});
    return level(reqs);
}

// =======================================================================================
// shift by an encrypted amount (barrel shifter); width must be a power of two bits
// =======================================================================================
static Radix barrel(Evaluator& ev, const Radix& a_in, const Radix& amount_in, bool left) {
    Radix a = a_in, amount = amount_in;
    ev.clean(a); ev.clean(amount);
    const int n = (int)a.size();
    int width_bits = 2 * n, nbits = 0;
    while ((1 << nbits) < width_bits) ++nbits;
    if ((1 << nbits) != width_bits) throw RadixError("encrypted shift needs a power-of-two bit width");
    // bits of the amount (amount mod width, like the reference backend)
    std::vector<Block> bit(nbits);
    {
        std::vector<Evaluator::Req> reqs;
        for (int k = 0; k < nbits; ++k) {
            const Block src = (k / 2) < (int)amount.size() ? amount[k / 2] : Block::constant(0);
            const int sh = k % 2;
            reqs.push_back({src, make_lut([sh](int v) { return (v >> sh) & 1; })});
        }
        bit = ev.level(reqs);
    }
    static const LutTable l_keep = make_sel_lut([](int x, int s) { return s ? 0 : (x & 3); });
    static const LutTable l_take = make_sel_lut([](int x, int s) { return s ? (x & 3) : 0; });
    static const LutTable l_r_self = make_sel_lut([](int x, int s) { return s ? ((x & 3) >> 1) : (x & 3); });
    static const LutTable l_r_next = make_sel_lut([](int y, int s) { return s ? ((y & 1) << 1) : 0; });
    static const LutTable l_l_self = make_sel_lut([](int x, int s) { return s ? ((x << 1) & 3) : (x & 3); });
    static const LutTable l_l_prev = make_sel_lut([](int y, int s) { return s ? ((y & 3) >> 1) : 0; });
    for (int k = 0; k < nbits; ++k) {
        std::vector<Evaluator::Req> reqs;
        std::vector<std::pair<int, int>> where;
        const int dist = k == 0 ? 1 : (1 << (k - 1));      // neighbour distance in blocks
        for (int i = 0; i < n; ++i) {
            const int j = left ? i - dist : i + dist;
            const bool has = j >= 0 && j < n;
            if (k == 0) {
                reqs.push_back({a[i] * 2 + bit[k], left ? l_l_self : l_r_self}); where.emplace_back(i, 0);
                if (has) { reqs.push_back({a[j] * 2 + bit[k], left ? l_l_prev : l_r_next}); where.emplace_back(i, 1); }
            } else {
                reqs.push_back({a[i] * 2 + bit[k], l_keep}); where.emplace_back(i, 0);
                if (has) { reqs.push_back({a[j] * 2 + bit[k], l_take}); where.emplace_back(i, 1); }
            }
        }
        std::vector<Block> o = ev.level(reqs);
        Radix nxt(n, Block::constant(0));
        for (size_t t = 0; t < o.size(); ++t) {
            Block& d = nxt[where[t].first];
            d = d + o[t];
            d.deg = 3;                                       // the two parts occupy disjoint bits / cases
        }
        a.swap(nxt);
    }
    return a;
}

Radix Evaluator::shr(const Radix& a, const Radix& amount) { return barrel(*this, a, amount, false); }
Radix Evaluator::shl(const Radix& a, const Radix& amount) { return barrel(*this, a, amount, true); }

// =======================================================================================
// comparisons and selection
// =======================================================================================
// per-block ordering code: 0 equal, 1 a < b, 2 a > b; reduced most-significant-first
Block Evaluator::lt(const Radix& a, const Radix& b) {
    static const LutTable l_is0 = make_lut([](int v) { return v == 0; });
    return level({{order_code(a, b), l_is0}})[0];
}

// Ordering code of two radix integers: 0 a < b, 1 equal, 2 a > b.  Reduced most-significant-first by a radix-3 tree:
// the codes are the digits of a binary adder again (propagate_radix3: "equal" is the transparent digit 1), so one
// lookup on S = 4 e_a + 2 e_b + e_c joins three segments; in a tree every node has exactly one role in the next level,
// so it is emitted once, as 2 e (roles a, b) or e (role c, or the root) - same bootstrap count as a binary tree,
// log3 n levels.
Block Evaluator::order_code(const Radix& a_in, const Radix& b_in) {
    Radix a = a_in, b = b_in;
    clean(a); clean(b);
    const int n = (int)std::max(a.size(), b.size());
    if (n == 0) return Block::constant(1);
    a = cast(a, n); b = cast(b, n);
    auto cmp = [](int x, int y) { return x < y ? 0 : (x > y ? 2 : 1); };
    static const LutTable l_cmp1 = make_bilut([cmp](int x, int y) { return cmp(x, y); });
    static const LutTable l_cmp2 = make_bilut([cmp](int x, int y) { return 2 * cmp(x, y); });
    auto join3 = [](int S) { return S >= 8 ? 2 : (S == 7 ? 1 : 0); };          // S = 4 e_a + 2 e_b + e_c
    auto join2 = [](int S) { return S >= 4 ? 2 : (S == 3 ? 1 : 0); };          // S = 2 e_b + e_c
    static const LutTable l_j31 = make_lut([join3](int S) { return join3(S); });
    static const LutTable l_j32 = make_lut([join3](int S) { return 2 * join3(S); });
    static const LutTable l_j21 = make_lut([join2](int S) { return S <= 6 ? join2(S) : 0; });
    static const LutTable l_j22 = make_lut([join2](int S) { return S <= 6 ? 2 * join2(S) : 0; });
    // sizes of the tree levels; a node that is alone in its group passes through unchanged, so its encoding is chosen
    // for the first level in which it has company
    std::vector<int> sizes{n};
    while (sizes.back() > 1) sizes.push_back((sizes.back() + 2) / 3);
    auto doubled = [&](size_t k, int j) {      // does node j of level k have to be emitted as 2 e ?
        while (k + 1 < sizes.size() && j == sizes[k] - 1 && sizes[k] % 3 == 1) { ++k; j = sizes[k] - 1; }
        if (sizes[k] == 1) return false;        // the root: plain code
        return j % 3 != 0;
    };
    std::vector<Req> reqs;
    for (int i = 0; i < n; ++i) reqs.push_back({a[i] * 4 + b[i], doubled(0, i) ? l_cmp2 : l_cmp1});
    std::vector<Block> st = level(reqs);
    for (size_t k = 0; k + 1 < sizes.size(); ++k) {
        reqs.clear();
        std::vector<Block> nxt((size_t)sizes[k + 1]);
        std::vector<int> where;
        for (int g = 0; g < sizes[k + 1]; ++g) {
            const int lo = 3 * g, cnt = std::min(3, sizes[k] - lo);
            if (cnt == 1) { nxt[g] = st[lo]; continue; }
            const bool dbl = doubled(k + 1, g);
            if (cnt == 2) reqs.push_back({st[lo + 1] + st[lo], dbl ? l_j22 : l_j21});
            else reqs.push_back({st[lo + 2] * 2 + st[lo + 1] + st[lo], dbl ? l_j32 : l_j31});
            where.push_back(g);
        }
        std::vector<Block> o = level(reqs);
        for (size_t q = 0; q < o.size(); ++q) nxt[where[q]] = o[q];
        st.swap(nxt);
    }
    return st[0];
}

// out = (code == 0, i.e. a < b) ? if_lt : otherwise, the ordering code used directly as the selector (one level less than
// lt() + select(): min / max of the reference's perf_test.rs:44)
Radix Evaluator::select_by_order(const Block& code_in, const Radix& lt_in, const Radix& ge_in) {
    Radix t = lt_in, f = ge_in, c{code_in};
    clean(t); clean(f); clean(c);
    const Block& code = c[0];
    const int n = (int)std::max(t.size(), f.size());
    t = cast(t, n); f = cast(f, n);
    static const LutTable l_lt = make_bilut([](int cd, int x) { return cd == 0 ? x : 0; });
    static const LutTable l_ge = make_bilut([](int cd, int x) { return cd == 0 ? 0 : x; });
    std::vector<Req> reqs;
    for (int i = 0; i < n; ++i) {
        reqs.push_back({code * 4 + t[i], l_lt});
        reqs.push_back({code * 4 + f[i], l_ge});
    }
    std::vector<Block> o = level(reqs);
    Radix out(n);
    for (int i = 0; i < n; ++i) {
        out[i] = o[2 * i] + o[2 * i + 1];
        out[i].deg = 3;
    }
    return out;
}

Block Evaluator::eq(const Radix& a_in, const Radix& b_in) {
    Radix a = a_in, b = b_in;
    clean(a); clean(b);
    const int n = (int)std::max(a.size(), b.size());
    if (n == 0) return Block::constant(1);
    a = cast(a, n); b = cast(b, n);
    static const LutTable l_eq = make_bilut([](int x, int y) { return x == y; });
    std::vector<Req> reqs;
    for (int i = 0; i < n; ++i) reqs.push_back({a[i] * 4 + b[i], l_eq});
    std::vector<Block> bits = level(reqs);
    while (bits.size() > 1) {
        reqs.clear();
        std::vector<Block> nxt;
        std::vector<size_t> where;
        for (size_t i = 0; i < bits.size(); i += kMaxNoise) {
            const size_t e = std::min(bits.size(), i + kMaxNoise);
            if (e - i == 1) { nxt.push_back(bits[i]); continue; }
            Block s2 = Block::constant(0);
            for (size_t j = i; j < e; ++j) s2 = s2 + bits[j];
            const int cnt = (int)(e - i);
            reqs.push_back({s2, make_lut([cnt](int v) { return v == cnt; })});
            where.push_back(nxt.size());
            nxt.push_back(Block());
        }
        std::vector<Block> o = level(reqs);
        for (size_t k = 0; k < o.size(); ++k) nxt[where[k]] = o[k];
        bits.swap(nxt);
    }
    return bits[0];
}

Radix Evaluator::select(const Block& cond_in, const Radix& t_in, const Radix& f_in) {
    Radix t = t_in, f = f_in, c{cond_in};
    clean(t); clean(f); clean(c);
    const Block& cond = c[0];
    const int n = (int)std::max(t.size(), f.size());
    t = cast(t, n); f = cast(f, n);
    static const LutTable l_true = make_sel_lut([](int x, int s) { return s ? (x & 3) : 0; });
    static const LutTable l_false = make_sel_lut([](int x, int s) { return s ? 0 : (x & 3); });
    std::vector<Req> reqs;
    for (int i = 0; i < n; ++i) {
        reqs.push_back({t[i] * 2 + cond, l_true});
        reqs.push_back({f[i] * 2 + cond, l_false});
    }
    std::vector<Block> o = level(reqs);
    Radix out(n);
    for (int i = 0; i < n; ++i) {
        out[i] = o[2 * i] + o[2 * i + 1];
        out[i].deg = 3;
    }
    return out;
}

Radix Evaluator::min(const Radix& a, const Radix& b) { return select_by_order(order_code(a, b), a, b); }
Radix Evaluator::max(const Radix& a, const Radix& b) { return select_by_order(order_code(a, b), b, a); }

// =======================================================================================
// division by a plaintext constant (Granlund-Montgomery: multiply-high by a magic number)
// =======================================================================================
namespace {
struct Nat {                                   // little-endian bits, tiny and slow by design (host, once per operator)
    std::vector<uint8_t> bit;
    explicit Nat(size_t n = 0) : bit(n, 0) {}
    static Nat from_digits(const std::vector<uint8_t>& d) {
        Nat r(d.size() * 2);
        for (size_t i = 0; i < d.size(); ++i) { r.bit[2 * i] = d[i] & 1; r.bit[2 * i + 1] = (d[i] >> 1) & 1; }
        return r;
    }
    std::vector<uint8_t> to_digits(size_t n_blocks) const {
        std::vector<uint8_t> d(n_blocks, 0);
        for (size_t i = 0; i < n_blocks; ++i) {
            const int b0 = 2 * i < bit.size() ? bit[2 * i] : 0, b1 = 2 * i + 1 < bit.size() ? bit[2 * i + 1] : 0;
            d[i] = (uint8_t)(b0 | (b1 << 1));
        }
        return d;
    }
    int top() const { for (int i = (int)bit.size() - 1; i >= 0; --i) if (bit[i]) return i; return -1; }
    bool zero() const { return top() < 0; }
};
int cmp(const Nat& a, const Nat& b) {
    const int ta = a.top(), tb = b.top();
    if (ta != tb) return ta < tb ? -1 : 1;
    for (int i = ta; i >= 0; --i) if (a.bit[i] != b.bit[i]) return a.bit[i] < b.bit[i] ? -1 : 1;
    return 0;
}
void sub_in_place(Nat& a, const Nat& b) {     // a >= b
    int borrow = 0;
    for (size_t i = 0; i < a.bit.size(); ++i) {
        int v = a.bit[i] - (i < b.bit.size() ? b.bit[i] : 0) - borrow;
        borrow = v < 0; a.bit[i] = (uint8_t)(v & 1);
    }
}
Nat divide(const Nat& num, const Nat& den, Nat* rem = nullptr) {  // floor(num / den), schoolbook binary long division
    Nat q(num.bit.size()), r(num.bit.size() + 1);
    for (int i = (int)num.bit.size() - 1; i >= 0; --i) {
        for (int j = (int)r.bit.size() - 1; j > 0; --j) r.bit[j] = r.bit[j - 1];
        r.bit[0] = num.bit[i];
        if (cmp(r, den) >= 0) { sub_in_place(r, den); q.bit[i] = 1; }
    }
    if (rem) *rem = r;
    return q;
}
Nat multiply(const Nat& a, const Nat& b) {
    Nat r(a.bit.size() + b.bit.size() + 1);
    for (size_t i = 0; i < a.bit.size(); ++i) {
        if (!a.bit[i]) continue;
        int carry = 0;
        for (size_t j = 0; j < b.bit.size() || carry; ++j) {
            const int v = r.bit[i + j] + (j < b.bit.size() ? b.bit[j] : 0) + carry;
            r.bit[i + j] = (uint8_t)(v & 1); carry = v >> 1;
        }
    }
    return r;
}
Nat nat_add(const Nat& a, const Nat& b) {
    Nat r(std::max(a.bit.size(), b.bit.size()) + 1);
    int carry = 0;
    for (size_t i = 0; i < r.bit.size(); ++i) {
        const int v = (i < a.bit.size() ? a.bit[i] : 0) + (i < b.bit.size() ? b.bit[i] : 0) + carry;
        r.bit[i] = (uint8_t)(v & 1); carry = v >> 1;
    }
    return r;
}
Nat shift_right(const Nat& a, size_t k) {
    Nat r(a.bit.size() > k ? a.bit.size() - k : 0);
    for (size_t i = 0; i < r.bit.size(); ++i) r.bit[i] = a.bit[i + k];
    return r;
}
void increment(Nat& m) {
    size_t i = 0;
    while (i < m.bit.size() && m.bit[i]) { m.bit[i] = 0; ++i; }
    if (i == m.bit.size()) m.bit.push_back(1); else m.bit[i] = 1;
}
}  // namespace

Radix Evaluator::scalar_div(const Radix& a, const std::vector<uint8_t>& d_digits, Radix* rem) {
    const int n = (int)a.size();
    const int W = 2 * n;
    const Nat d = Nat::from_digits(d_digits);
    if (d.zero()) throw RadixError("division by zero");
    const int top = d.top();
    if (top >= W) {                              // divisor wider than the dividend type: quotient 0
        if (rem) *rem = a;
        return trivial_big(std::vector<uint8_t>(n, 0));
    }
    bool pow2 = true;
    for (int i = 0; i < top; ++i) pow2 = pow2 && !d.bit[i];
    Radix q;
    if (pow2) {
        q = scalar_shr(a, (unsigned)top);
    } else {
        const int l = top + 1;                   // ceil(log2 d) for non powers of two
        // Round-up method: M_s = ceil(2^(W+s) / d) divides exactly for every W-bit dividend as soon as its excess
        // e = M_s d - 2^(W+s) is at most 2^s.  If such an M_s still fits W bits, floor(a / d) = mulhi(a, M_s) >> s with
        // no fix-up at all (d = 5: M = 0xCC...CD, s = 2): one product and a shift instead of product, subtraction,
        // shift, addition, shift - 10 PBS levels instead of 21 for the reference's `x / 5` (perf_test.rs:54).
        int s_fit = -1;
        Nat m_fit;
        for (int sft = 0; sft <= l && s_fit < 0; ++sft) {
            Nat num((size_t)W + sft + 1);
            num.bit[W + sft] = 1;
            Nat r;
            Nat m = divide(num, d, &r);
            Nat e(d.bit.size() + 1);                 // excess of the rounded-up multiplier: d - r (0 if d divides)
            if (!r.zero()) { e = Nat(d.bit.size() + 1); for (size_t i = 0; i < d.bit.size(); ++i) e.bit[i] = d.bit[i]; sub_in_place(e, r); increment(m); }
            Nat lim((size_t)sft + 1);
            lim.bit[sft] = 1;                        // 2^s
            if (cmp(e, lim) <= 0 && m.top() < W) { s_fit = sft; m_fit = m; }
        }
        if (s_fit >= 0) {
            Radix wide = scalar_mul(cast(a, 2 * n), m_fit.to_digits(n), 2 * n);      // full 2W-bit product
            Radix t1(wide.begin() + n, wide.end());                               // mulhi
            q = scalar_shr(t1, (unsigned)s_fit);
        } else {
            // the multiplier needs W + 1 bits, M = 2^W + m': floor(a / d) = (mulhi(a, m') + a) >> l, the sum taken one
            // block wider instead of the register-width dance (a - t) / 2 + t
            Nat num((size_t)W + l + 1);
            {
                Nat t((size_t)l + 1);
                t.bit[l] = 1;
                sub_in_place(t, d);                  // 2^l - d
                for (int i = 0; i <= l; ++i) if (t.bit[i]) num.bit[i + W] = 1;
            }
            Nat m = divide(num, d);                  // m' = floor(2^W (2^l - d) / d) + 1, fits W bits
            increment(m);
            Radix wide = scalar_mul(cast(a, 2 * n), m.to_digits(n), 2 * n);
            Radix t1(wide.begin() + n, wide.end());
            Radix sum = add(cast(t1, n + 1), cast(a, n + 1));
            q = cast(scalar_shr(sum, (unsigned)l), n);
        }
    }
    if (rem) {
        Radix qd = scalar_mul(q, d.to_digits(n), n);
        *rem = sub(a, qd);
    }
    return q;
}

// a mod d.  Moduli of the form 2^k - c with a small c (the secp256k1 group order of the reference's scalar.rs:8 is
// 2^256 - c, c < 2^129) are reduced by folding, hi 2^k + lo = hi c + lo (mod d): each fold is one multiplication by the
// constant c with the low part riding in the same column sum, the value shrinks by about k - bits(c) bits per fold, and
// once it is provably below 2 d one conditional subtraction finishes.  A 514-bit value takes three folds (258-, 132- and
// 6-bit high parts) instead of the two 257-block products of the general quotient-and-multiply-back path.
Radix Evaluator::scalar_rem(const Radix& a, const std::vector<uint8_t>& d_digits) {
    const Nat d = Nat::from_digits(d_digits);
    if (d.zero()) throw RadixError("division by zero");
    const int k = d.top() + 1;
    Nat c((size_t)k + 1);
    c.bit[k] = 1;
    sub_in_place(c, d);                              // 2^k - d
    const int cbits = c.top() + 1;
    bool foldable = (k % 2 == 0) && cbits > 0 && cbits <= k / 2 + 1 && 2 * (int)a.size() > k;
    if (foldable) {      // dry run on the bounds: the folds must bring the value below 2 d within a few rounds
        Nat ub0(2 * a.size());
        for (auto& b : ub0.bit) b = 1;
        Nat ones((size_t)k);
        for (auto& b : ones.bit) b = 1;
        const Nat dd = nat_add(d, d);
        int rounds = 0;
        while (cmp(ub0, dd) >= 0 && rounds <= 8) { ub0 = nat_add(ones, multiply(shift_right(ub0, (size_t)k), c)); ++rounds; }
        foldable = cmp(ub0, dd) < 0;
    }
    if (!foldable) {
        Radix r;
        scalar_div(a, d_digits, &r);
        return r;
    }
    const int kb = k / 2;                            // blocks below 2^k
    const std::vector<uint8_t> cdig = c.to_digits((size_t)(cbits + 1) / 2);
    Nat two_d = nat_add(d, d);
    Nat ub(2 * a.size());
    for (auto& b : ub.bit) b = 1;                    // upper bound of the running value
    Nat low_ones((size_t)k);
    for (auto& b : low_ones.bit) b = 1;              // 2^k - 1
    Radix r = a;
    for (int guard = 0; cmp(ub, two_d) >= 0; ++guard) {
        if (guard > 16) throw RadixError("scalar_rem: fold did not converge");
        const Nat hi_ub = shift_right(ub, (size_t)k);
        Nat nub = nat_add(low_ones, multiply(hi_ub, c));
        const int out_blocks = (nub.top() + 2) / 2;
        Radix lo(r.begin(), r.begin() + std::min<size_t>(kb, r.size()));
        Radix hi(r.size() > (size_t)kb ? r.begin() + kb : r.end(), r.end());
        hi.resize(std::min<size_t>(hi.size(), (size_t)(hi_ub.top() + 2) / 2), Block::constant(0));   // blocks above the bound are zero
        r = scalar_mul_add(hi, cdig, &lo, out_blocks);
        ub = nub;
    }
    if (cmp(ub, d) >= 0) {                           // r < 2 d: subtract d where that does not borrow
        const int w = (int)r.size();
        Block not_borrow;
        Radix t = sub(r, trivial_big(d.to_digits((size_t)w)), &not_borrow);
        r = select(not_borrow, t, r);
    }
    return cast(r, (int)a.size());
}

}  // namespace fsc
