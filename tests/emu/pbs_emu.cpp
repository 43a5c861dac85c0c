// pbs_emu.cpp — executes the ring kernel's per-lane core (fhe_sign_b200/csrc/pbs_core.cuh, and pass32 /
// pass32_inv_gs of pbs_core2.cuh) lane by lane on the CPU.  TEST INFRASTRUCTURE: proves the index/twiddle/rounding
// logic of pbs_ring_kernel without a GPU.  Each "phase" below is one region between two __syncwarp() in
// pbs_kernel.cu.  `form` selects the arithmetic the kernel configuration uses:
//   0  plain (re, im) constants, Cooley-Tukey forward / Gentleman-Sande inverse (64-bit accumulator, pair kernel)
//   1  half-step configuration (32-bit accumulator): forward passes in the tangent form (pass32), inverse pass 2
//      rebuilt from the same (cos, tan) table (pass32_inv_gs), inverse pass 1 plain
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../fhe_sign_b200/csrc/pbs_core2.cuh"

using namespace fsc;

namespace {
struct Lane {
    cplx s2[16];      // pass-2 constants of the lane: plain (form 0) or pass32 form (form 1)
    cplx X0[32], X1[32];
};
cplx g_p1[16], g_wt0[16];

void init_lanes(std::vector<Lane>& L, int form) {
    for (int l = 0; l < 32; ++l) {
        if (form == 0) lane_consts(4 * l + 1, L[l].s2);
        else for (int ci = 0; ci < 16; ++ci) L[l].s2[ci] = pass_const(ci, 4 * l + 1);
    }
}
void fft_fwd_warp(Lane* L, cplx (*v)[32], cplx* xbuf, int form) {
    for (int l = 0; l < 32; ++l) { if (form == 0) dft32_fwd(v[l], PtrConsts{g_p1}); else pass32(v[l], PtrConsts{g_wt0}); }
    for (int l = 0; l < 32; ++l) xpose_store_fwd(l, xbuf, v[l]);
    for (int l = 0; l < 32; ++l) xpose_load_fwd(l, xbuf, v[l]);
    for (int l = 0; l < 32; ++l) { if (form == 0) dft32_fwd(v[l], RegConsts(L[l].s2)); else pass32(v[l], RegConsts(L[l].s2)); }
}
void fft_inv_warp(Lane* L, cplx (*v)[32], cplx* xbuf, int form) {
    for (int l = 0; l < 32; ++l) { if (form == 0) dft32_inv(v[l], RegConsts(L[l].s2)); else pass32_inv_gs(v[l], RegConsts(L[l].s2)); }
    for (int l = 0; l < 32; ++l) xpose_store_inv(l, xbuf, v[l]);
    for (int l = 0; l < 32; ++l) xpose_load_inv(l, xbuf, v[l]);
    for (int l = 0; l < 32; ++l) dft32_inv(v[l], PtrConsts{g_p1});
}

template <typename AccT>
void blind_rotate(int form, int n, int base_log, const cplx* bsk_f, const uint64_t* ct, const uint64_t* lut, uint64_t* out) {
    std::vector<pair_t<AccT>> acc(2 * 1024);
    std::vector<cplx> xbuf(1024);
    std::vector<Lane> L(32);
    static cplx v[32][32];
    init_lanes(L, form);
    const int b = modswitch(ct[n]);
    for (int idx = 0; idx < 1024; ++idx) { acc[idx].x = 0; acc[idx].y = 0; acc[1024 + idx] = lut_pair<AccT>(lut, idx, b); }
    for (int i = 0; i < n; ++i) {
        const int a = modswitch(ct[i]);
        for (int p = 0; p < 2; ++p) {
            for (int l = 0; l < 32; ++l) cmux_head<AccT>(l, acc.data() + p * 1024, a, base_log, v[l]);
            fft_fwd_warp(L.data(), v, xbuf.data(), form);
            for (int l = 0; l < 32; ++l) memcpy(p ? L[l].X1 : L[l].X0, v[l], sizeof(v[l]));
        }
        const cplx* g = bsk_f + (size_t)i * 32 * 4 * 32;
        for (int l = 0; l < 32; ++l)
            for (int r = 0; r < 32; ++r) {
                G4 q; q.g00 = g[(r * 4 + 0) * 32 + l]; q.g01 = g[(r * 4 + 1) * 32 + l];
                q.g10 = g[(r * 4 + 2) * 32 + l]; q.g11 = g[(r * 4 + 3) * 32 + l];
                mac_one(L[l].X0[r], L[l].X1[r], q);
            }
        for (int p = 0; p < 2; ++p) {
            for (int l = 0; l < 32; ++l) memcpy(v[l], p ? L[l].X1 : L[l].X0, sizeof(v[l]));
            fft_inv_warp(L.data(), v, xbuf.data(), form);
            for (int l = 0; l < 32; ++l) cmux_tail<AccT>(l, acc.data() + p * 1024, v[l]);
        }
    }
    for (int j = 0; j <= kN; ++j) out[j] = extract_word<AccT>(acc.data(), acc.data() + 1024, j);
}
}  // namespace

extern "C" {
void emu_init() {
    lane_consts(kP1G, g_p1);
    for (int ci = 0; ci < 16; ++ci) g_wt0[ci] = pass_const(ci, kP1G);
}

// standard-domain BSK [n][2][1][2][2048] -> Fourier layout [n][32 r][4 g][32 lane]  (bsk_convert_kernel: plain form)
void emu_convert_bsk(int n, const uint64_t* bsk, double* out_f) {
    cplx* o = reinterpret_cast<cplx*>(out_f);
    std::vector<Lane> L(32);
    std::vector<cplx> xbuf(1024);
    static cplx v[32][32];
    init_lanes(L, 0);
    for (int i = 0; i < n; ++i)
        for (int g = 0; g < 4; ++g) {
            const uint64_t* src = bsk + ((size_t)i * 4 + g) * kN;
            for (int l = 0; l < 32; ++l)
                for (int j2 = 0; j2 < 32; ++j2) {
                    v[l][j2].x = (double)(int64_t)src[l + 32 * j2];
                    v[l][j2].y = (double)(int64_t)src[l + 32 * j2 + 1024];
                }
            fft_fwd_warp(L.data(), v, xbuf.data(), 0);
            for (int l = 0; l < 32; ++l)
                for (int r = 0; r < 32; ++r) o[(((size_t)i * 32 + r) * 4 + g) * 32 + l] = v[l][r];
        }
}

void emu_blind_rotate(int acc_bits, int form, int n, int base_log, const double* bsk_f, const uint64_t* cts, int count,
                      const uint64_t* lut, uint64_t* out) {
    const cplx* f = reinterpret_cast<const cplx*>(bsk_f);
    for (int c = 0; c < count; ++c) {
        if (acc_bits == 64) blind_rotate<uint64_t>(form, n, base_log, f, cts + (size_t)c * (n + 1), lut, out + (size_t)c * (kN + 1));
        else blind_rotate<uint32_t>(form, n, base_log, f, cts + (size_t)c * (n + 1), lut, out + (size_t)c * (kN + 1));
    }
}

// c = a (torus) * b (small ints) through the warp FFT; result rounded like the kernel's tail
void emu_negacyclic_mul(int form, const uint64_t* a, const int64_t* b, uint64_t* c) {
    std::vector<Lane> L(32);
    std::vector<cplx> xbuf(1024);
    static cplx va[32][32], vb[32][32];
    init_lanes(L, form);
    for (int l = 0; l < 32; ++l)
        for (int j2 = 0; j2 < 32; ++j2) {
            va[l][j2].x = (double)(int64_t)a[l + 32 * j2]; va[l][j2].y = (double)(int64_t)a[l + 32 * j2 + 1024];
            vb[l][j2].x = (double)b[l + 32 * j2];          vb[l][j2].y = (double)b[l + 32 * j2 + 1024];
        }
    fft_fwd_warp(L.data(), va, xbuf.data(), form);
    fft_fwd_warp(L.data(), vb, xbuf.data(), form);
    for (int l = 0; l < 32; ++l)
        for (int r = 0; r < 32; ++r) {
            cplx x = va[l][r], y = vb[l][r];
            va[l][r].x = x.x * y.x - x.y * y.y; va[l][r].y = x.x * y.y + x.y * y.x;
        }
    fft_inv_warp(L.data(), va, xbuf.data(), form);
    for (int l = 0; l < 32; ++l)
        for (int j2 = 0; j2 < 32; ++j2) {
            c[l + 32 * j2] = to_acc<uint64_t>(va[l][j2].x);
            c[l + 32 * j2 + 1024] = to_acc<uint64_t>(va[l][j2].y);
        }
}
}
