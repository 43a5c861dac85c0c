// pbs_emu2.cpp — executes fhe_sign_b200/csrc/pbs_core2.cuh (the single-routine "stream" blind rotation) lane by
// lane on the CPU.  TEST INFRASTRUCTURE: the no-GPU proof of the table, transpose, bit-reversal-absorbing product
// and twist logic of pbs_stream_kernel.cu; each block below is one region between two warp syncs of the kernel.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../fhe_sign_b200/csrc/pbs_core2.cuh"

using namespace fsc;

namespace {
struct Tables {
    cplx t[4][16][32];          // [pass][ci][lane]
    cplx tw64[32][32], tw32[32][32];
    Tables() {
        for (int q = 0; q < 4; ++q)
            for (int ci = 0; ci < 16; ++ci)
                for (int l = 0; l < 32; ++l) t[q][ci][l] = pass_const(ci, pass_g(q, l));
        for (int pos = 0; pos < 32; ++pos)
            for (int l = 0; l < 32; ++l) { tw64[pos][l] = twist_const<uint64_t>(pos, l); tw32[pos][l] = twist_const<uint32_t>(pos, l); }
    }
};
const Tables& tables() { static Tables T; return T; }

int g_uniform = 0;      // 1: the uniform passes (q = 0, 2) through pass32_uniform (what pbs_stream_tx_kernel<1..2> runs)
void run_pass(int q, cplx (*v)[32]) {
    const Tables& T = tables();
    for (int l = 0; l < 32; ++l) {
        const StridedConsts sp{&T.t[q][0][l], 32};
        if (g_uniform && q == 0) pass32_uniform<32>(v[l], sp);
        else if (g_uniform && q == 2) pass32_uniform<0>(v[l], sp);
        else if (g_uniform && q == 1) {      // level 1 in the tangent form: entry 0 as (cos, tan), what the kernel's table copy makes of it
            cplx t[16];
            for (int ci = 0; ci < 16; ++ci) t[ci] = T.t[q][ci][l];
            t[0].y = t[0].y / t[0].x;
            pass32<true>(v[l], StridedConsts{t, 1});
        }
        else pass32(v[l], sp);
    }
}
void transpose(cplx (*v)[32], bool inverse) {
    std::vector<double> xb(kXBufDoubles);
    for (int comp = 0; comp < 2; ++comp) {
        for (int l = 0; l < 32; ++l) xp_store(l, xb.data(), v[l], comp);
        for (int l = 0; l < 32; ++l) xp_load(inverse ? (32 - l) & 31 : l, xb.data(), v[l], comp);
    }
}
void fft_fwd(cplx (*v)[32]) { run_pass(0, v); transpose(v, false); run_pass(1, v); }
// v: slot k1 natural (after the product); leaves the untwisted result (slot pos <-> j2 = tail_j2(pos))
void fft_inv(cplx (*v)[32]) { run_pass(2, v); transpose(v, true); run_pass(3, v); }

template <int R0>
void mac_all(cplx (*X)[32], const cplx (*other)[32], const cplx* g, int own, int oth) {
    for (int l = 0; l < 32; ++l) {
        cplx o[4], gw[4], go[4];
        for (int rr = 0; rr < 4; ++rr) {
            const int pos = R0 + rr;
            // the partner's slot s holds frequency brev5(s): frequency freq_at(pos) sits in its slot brev5(freq_at(pos))
            o[rr] = other[l][brev5(freq_at(pos))];
            gw[rr] = g[(pos * 4 + own) * 32 + l];
            go[rr] = g[(pos * 4 + oth) * 32 + l];
        }
        mac_chunk<R0>(X[l], o, gw, go);
    }
}

template <typename AccT>
void blind_rotate(int n, int base_log, const cplx* bsk_f, const uint64_t* ct, const uint64_t* lut, uint64_t* out) {
    const Tables& T = tables();
    std::vector<pair_t<AccT>> acc(2 * 1024);
    static cplx v[2][32][32], w[2][32][32];
    const int b = modswitch(ct[n]);
    for (int idx = 0; idx < 1024; ++idx) { acc[idx].x = 0; acc[idx].y = 0; acc[1024 + idx] = lut_pair<AccT>(lut, idx, b); }
    for (int i = 0; i < n; ++i) {
        const int a = modswitch(ct[i]);
        for (int p = 0; p < 2; ++p) {
            for (int l = 0; l < 32; ++l) cmux_head<AccT>(l, acc.data() + p * 1024, a, base_log, v[p][l]);
            fft_fwd(v[p]);
        }
        memcpy(w, v, sizeof(v));
        const cplx* g = bsk_f + (size_t)i * 32 * 4 * 32;
        for (int p = 0; p < 2; ++p) {      // warp p: X_p <- X_p G[p][p] + X_{1-p} G[1-p][p]   (g = 2 row + col)
            const int own = 3 * p, oth = 2 - p;
            mac_all<0>(v[p], w[1 - p], g, own, oth);  mac_all<4>(v[p], w[1 - p], g, own, oth);
            mac_all<8>(v[p], w[1 - p], g, own, oth);  mac_all<12>(v[p], w[1 - p], g, own, oth);
            mac_all<16>(v[p], w[1 - p], g, own, oth); mac_all<20>(v[p], w[1 - p], g, own, oth);
            mac_all<24>(v[p], w[1 - p], g, own, oth); mac_all<28>(v[p], w[1 - p], g, own, oth);
        }
        for (int p = 0; p < 2; ++p) {
            fft_inv(v[p]);
            const cplx* tw = sizeof(AccT) == 8 ? &T.tw64[0][0] : &T.tw32[0][0];
            for (int l = 0; l < 32; ++l) stream_tail<AccT>(l, acc.data() + p * 1024, tw, v[p][l]);
        }
    }
    for (int j = 0; j <= kN; ++j) out[j] = extract_word<AccT>(acc.data(), acc.data() + 1024, j);
}
}  // namespace

extern "C" {
void emu2_set_uniform(int on) { g_uniform = on; }
// standard-domain BSK [n][2][1][2][2048] -> Fourier layout [n][32 position][4 g][32 lane]
void emu2_convert_bsk(int n, const uint64_t* bsk, double* out_f) {
    cplx* o = reinterpret_cast<cplx*>(out_f);
    static cplx v[32][32];
    for (int i = 0; i < n; ++i)
        for (int g = 0; g < 4; ++g) {
            const uint64_t* src = bsk + ((size_t)i * 4 + g) * kN;
            for (int l = 0; l < 32; ++l)
                for (int j2 = 0; j2 < 32; ++j2) {
                    v[l][j2].x = (double)(int64_t)src[l + 32 * j2];
                    v[l][j2].y = (double)(int64_t)src[l + 32 * j2 + 1024];
                }
            fft_fwd(v);
            for (int l = 0; l < 32; ++l)
                for (int s = 0; s < 32; ++s) o[(((size_t)i * 32 + slot_position(s)) * 4 + g) * 32 + l] = v[l][s];
        }
}

void emu2_blind_rotate(int acc_bits, int n, int base_log, const double* bsk_f, const uint64_t* cts, int count,
                       const uint64_t* lut, uint64_t* out) {
    const cplx* f = reinterpret_cast<const cplx*>(bsk_f);
    for (int c = 0; c < count; ++c) {
        if (acc_bits == 64) blind_rotate<uint64_t>(n, base_log, f, cts + (size_t)c * (n + 1), lut, out + (size_t)c * (kN + 1));
        else blind_rotate<uint32_t>(n, base_log, f, cts + (size_t)c * (n + 1), lut, out + (size_t)c * (kN + 1));
    }
}

// c = a (torus) * b (small ints) through the stream FFT; result rounded like the kernel's tail
void emu2_negacyclic_mul(const uint64_t* a, const int64_t* b, uint64_t* c) {
    const Tables& T = tables();
    static cplx va[32][32], vb[32][32];
    for (int l = 0; l < 32; ++l)
        for (int j2 = 0; j2 < 32; ++j2) {
            va[l][j2].x = (double)(int64_t)a[l + 32 * j2]; va[l][j2].y = (double)(int64_t)a[l + 32 * j2 + 1024];
            vb[l][j2].x = (double)b[l + 32 * j2];          vb[l][j2].y = (double)b[l + 32 * j2 + 1024];
        }
    fft_fwd(va);
    fft_fwd(vb);
    static cplx prod[32][32];
    for (int l = 0; l < 32; ++l)
        for (int s = 0; s < 32; ++s) {      // slot s holds k1 = brev5(s): move to slot k1
            const cplx x = va[l][s], y = vb[l][s];
            prod[l][brev5(s)].x = x.x * y.x - x.y * y.y; prod[l][brev5(s)].y = x.x * y.y + x.y * y.x;
        }
    fft_inv(prod);
    std::vector<pair_t<uint64_t>> acc(1024);
    for (auto& e : acc) { e.x = 0; e.y = 0; }
    for (int l = 0; l < 32; ++l) stream_tail<uint64_t>(l, acc.data(), &T.tw64[0][0], prod[l]);
    for (int idx = 0; idx < 1024; ++idx) { c[idx] = acc[idx].x; c[idx + 1024] = acc[idx].y; }
}

// smallest |cos| over the tangent-form constants (levels 2..5) of the four tables
double emu2_min_cos() {
    double m = 1.0;
    for (int q = 0; q < 4; ++q)
        for (int ci = 1; ci < 16; ++ci)
            for (int l = 0; l < 32; ++l) { const double c = fabs(twiddle4096(node_exponent(ci, pass_g(q, l))).x); if (c < m) m = c; }
    return m;
}
}
