// pbs_emu.cpp — executes fhe_sign_b200/csrc/pbs_core.cuh lane by lane on the CPU.
// TEST INFRASTRUCTURE: proves the index/twiddle/rounding logic of the warp-resident blind rotation
// without a GPU.  Each "phase" below is one region between two __syncwarp() in pbs_kernel.cu.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../fhe_sign_b200/csrc/pbs_core.cuh"

using namespace fsc;

namespace {
struct Lane {
    cplx s2[16];
    cplx X0[32], X1[32];
};
cplx g_uni[kUniSize];

void fft_fwd_warp(Lane* L, cplx (*v)[32], cplx* xbuf) {
    PtrConsts c1{g_uni};
    for (int l = 0; l < 32; ++l) dft32_fwd_tan<kP1Center, kP1MinLevel>(v[l], c1);
    for (int l = 0; l < 32; ++l) xpose_store_fwd(l, xbuf, v[l]);
    for (int l = 0; l < 32; ++l) xpose_load_fwd(l, xbuf, v[l]);
    for (int l = 0; l < 32; ++l) dft32_fwd_tan<kP2Center, kP2MinLevel>(v[l], RegConsts(L[l].s2));
}
// output already scaled by the table at tw_base (kUniTw64 / kUniTw32)
void fft_inv_warp(Lane* L, cplx (*v)[32], cplx* xbuf, int tw_base) {
    PtrConsts c1{g_uni};
    for (int l = 0; l < 32; ++l) dft32_inv_tan<kP2Center, kP2MinLevel>(v[l], RegConsts(L[l].s2));
    for (int l = 0; l < 32; ++l) xpose_store_inv(l, xbuf, v[l]);
    for (int l = 0; l < 32; ++l) xpose_load_inv(l, xbuf, v[l]);
    for (int l = 0; l < 32; ++l) idft32_dit_twist(v[l], c1, tw_base);
}

template <typename AccT>
void blind_rotate(int n, int base_log, const cplx* bsk_f, const uint64_t* ct, const uint64_t* lut, uint64_t* out) {
    std::vector<pair_t<AccT>> acc(2 * 1024);
    std::vector<cplx> xbuf(1024);
    std::vector<Lane> L(32);
    static cplx v[32][32];
    for (int l = 0; l < 32; ++l) lane_consts_tan(4 * l + 1, kP2Center, kP2MinLevel, L[l].s2);
    const int b = modswitch(ct[n]);
    for (int idx = 0; idx < 1024; ++idx) { acc[idx].x = 0; acc[idx].y = 0; acc[1024 + idx] = lut_pair<AccT>(lut, idx, b); }
    for (int i = 0; i < n; ++i) {
        const int a = modswitch(ct[i]);
        if (a == 0) continue;
        for (int p = 0; p < 2; ++p) {
            for (int l = 0; l < 32; ++l) cmux_head<AccT>(l, acc.data() + p * 1024, a, base_log, v[l]);
            fft_fwd_warp(L.data(), v, xbuf.data());
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
            fft_inv_warp(L.data(), v, xbuf.data(), uni_tw<AccT>::base);
            for (int l = 0; l < 32; ++l) cmux_tail_scaled<AccT>(l, acc.data() + p * 1024, v[l]);
        }
    }
    for (int j = 0; j <= kN; ++j) out[j] = extract_word<AccT>(acc.data(), acc.data() + 1024, j);
}
}  // namespace

extern "C" {
void emu_init() { fill_uniform_table(g_uni); }

// standard-domain BSK [n][2][1][2][2048] -> Fourier layout [n][32 r][4 g][32 lane]
void emu_convert_bsk(int n, const uint64_t* bsk, double* out_f) {
    cplx* o = reinterpret_cast<cplx*>(out_f);
    std::vector<Lane> L(32);
    std::vector<cplx> xbuf(1024);
    static cplx v[32][32];
    for (int l = 0; l < 32; ++l) lane_consts_tan(4 * l + 1, kP2Center, kP2MinLevel, L[l].s2);
    for (int i = 0; i < n; ++i)
        for (int g = 0; g < 4; ++g) {
            const uint64_t* src = bsk + ((size_t)i * 4 + g) * kN;
            for (int l = 0; l < 32; ++l)
                for (int j2 = 0; j2 < 32; ++j2) {
                    v[l][j2].x = (double)(int64_t)src[l + 32 * j2];
                    v[l][j2].y = (double)(int64_t)src[l + 32 * j2 + 1024];
                }
            fft_fwd_warp(L.data(), v, xbuf.data());
            for (int l = 0; l < 32; ++l)
                for (int r = 0; r < 32; ++r) o[(((size_t)i * 32 + r) * 4 + g) * 32 + l] = v[l][r];
        }
}

void emu_blind_rotate(int acc_bits, int n, int base_log, const double* bsk_f, const uint64_t* cts, int count,
                      const uint64_t* lut, uint64_t* out) {
    const cplx* f = reinterpret_cast<const cplx*>(bsk_f);
    for (int c = 0; c < count; ++c) {
        if (acc_bits == 64) blind_rotate<uint64_t>(n, base_log, f, cts + (size_t)c * (n + 1), lut, out + (size_t)c * (kN + 1));
        else blind_rotate<uint32_t>(n, base_log, f, cts + (size_t)c * (n + 1), lut, out + (size_t)c * (kN + 1));
    }
}

// c = a (torus) * b (small ints) through the warp FFT; result rounded like the kernel's tail
void emu_negacyclic_mul(const uint64_t* a, const int64_t* b, uint64_t* c) {
    std::vector<Lane> L(32);
    std::vector<cplx> xbuf(1024);
    static cplx va[32][32], vb[32][32];
    for (int l = 0; l < 32; ++l) lane_consts_tan(4 * l + 1, kP2Center, kP2MinLevel, L[l].s2);
    for (int l = 0; l < 32; ++l)
        for (int j2 = 0; j2 < 32; ++j2) {
            va[l][j2].x = (double)(int64_t)a[l + 32 * j2]; va[l][j2].y = (double)(int64_t)a[l + 32 * j2 + 1024];
            vb[l][j2].x = (double)b[l + 32 * j2];          vb[l][j2].y = (double)b[l + 32 * j2 + 1024];
        }
    fft_fwd_warp(L.data(), va, xbuf.data());
    fft_fwd_warp(L.data(), vb, xbuf.data());
    for (int l = 0; l < 32; ++l)
        for (int r = 0; r < 32; ++r) {
            cplx x = va[l][r], y = vb[l][r];
            va[l][r].x = x.x * y.x - x.y * y.y; va[l][r].y = x.x * y.y + x.y * y.x;
        }
    fft_inv_warp(L.data(), va, xbuf.data(), kUniTw64);
    for (int l = 0; l < 32; ++l)
        for (int j2 = 0; j2 < 32; ++j2) {
            c[l + 32 * j2] = to_acc_scaled<uint64_t>(va[l][j2].x);
            c[l + 32 * j2 + 1024] = to_acc_scaled<uint64_t>(va[l][j2].y);
        }
}
}
