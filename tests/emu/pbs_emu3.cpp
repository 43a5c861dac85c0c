// pbs_emu3.cpp — executes fhe_sign_b200/csrc/pbs_core3.cuh (the four-warps-per-ciphertext "split" blind rotation)
// lane by lane on the CPU.  TEST INFRASTRUCTURE: the no-GPU proof of the half-pass, exchange-buffer and product
// logic of pbs_split_kernel.cu; each loop nest below is one region between two block barriers of the kernel.
// The Fourier key comes from emu2_convert_bsk (same layout as the stream kernel).
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../fhe_sign_b200/csrc/pbs_core3.cuh"

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

template <typename AccT>
void blind_rotate(int px, int n, int base_log, const cplx* bsk_f, const uint64_t* ct, const uint64_t* lut, uint64_t* out) {
    const Tables& T = tables();
    std::vector<pair_t<AccT>> acc(2 * 1024);
    std::vector<cplx> E(2 * kSplitECplx), Tb(2 * kSplitTCplx), X(2 * kSplitXCplx), X2(2 * kSplitX2Cplx);
    static cplx w[2][2][32][16];          // registers of warp (p, h), lane
    const int b = modswitch(ct[n]);
    for (int idx = 0; idx < 1024; ++idx) { acc[idx].x = 0; acc[idx].y = 0; acc[1024 + idx] = lut_pair<AccT>(lut, idx, b); }
    const cplx* tw = sizeof(AccT) == 8 ? &T.tw64[0][0] : &T.tw32[0][0];
#define ALL for (int p = 0; p < 2; ++p) for (int h = 0; h < 2; ++h) for (int l = 0; l < 32; ++l)
    for (int i = 0; i < n; ++i) {
        const int a = modswitch(ct[i]);
        ALL split_head<AccT>(l, h, acc.data() + p * 1024, a, base_log, E.data() + p * kSplitECplx);
        // barrier
        if (px == 1) {      // the kernel's default form: uniform passes with the root parameter at compile time
            ALL { split_level1_u<32>(h, SplitLoadS{E.data() + p * kSplitECplx + l, 32}, StridedConsts{&T.t[0][0][l], 32}, w[p][h][l]);
                  split_levels25(h, StridedConsts{&T.t[0][0][l], 32}, w[p][h][l]); }
        } else {
            ALL split_pass(h, SplitLoadS{E.data() + p * kSplitECplx + l, 32}, StridedConsts{&T.t[0][0][l], 32}, w[p][h][l]);
        }
        ALL split_xp_store(l, h, Tb.data() + p * kSplitTCplx, w[p][h][l]);
        // barrier
        ALL split_pass(h, SplitLoadT{Tb.data() + p * kSplitTCplx + l * kSplitTRow}, StridedConsts{&T.t[1][0][l], 32}, w[p][h][l]);
        ALL split_spec_store(l, h, E.data() + p * kSplitECplx, w[p][h][l]);
        // barrier
        const cplx* g = bsk_f + (size_t)i * 32 * 4 * 32;
        if (px == 2) {  // run-time-h form of the split product (what the kernel runs): every level-1 output through X2
            ALL {
                SplitLoadProduct ld{E.data() + p * kSplitECplx + l, E.data() + (1 - p) * kSplitECplx + l, g + l, g + 16 * 4 * 32 + l, 3 * p, 2 - p};
                split_product_send2(l, h, ld, StridedConsts{&T.t[2][0][l], 32}, X2.data() + p * kSplitX2Cplx);
            }
            // barrier
            ALL {
                split_product_recv2(l, h, X2.data() + p * kSplitX2Cplx, w[p][h][l]);
                split_levels25(h, StridedConsts{&T.t[2][0][l], 32}, w[p][h][l]);
            }
        } else if (!px) {      // product by both warps of a polynomial, feeding level 1 directly
            ALL {
                SplitLoadProduct ld{E.data() + p * kSplitECplx + l, E.data() + (1 - p) * kSplitECplx + l, g + l, g + 16 * 4 * 32 + l, 3 * p, 2 - p};
                split_pass(h, ld, StridedConsts{&T.t[2][0][l], 32}, w[p][h][l]);
            }
        } else {        // product split between the two warps, level-1 outputs of the other half exchanged
            ALL {
                SplitLoadProduct ld{E.data() + p * kSplitECplx + l, E.data() + (1 - p) * kSplitECplx + l, g + l, g + 16 * 4 * 32 + l, 3 * p, 2 - p};
                split_product_send<true>(l, h, ld, StridedConsts{&T.t[2][0][l], 32}, X.data() + p * kSplitXCplx, w[p][h][l]);
            }
            // barrier
            ALL {
                split_product_recv(l, h, X.data() + p * kSplitXCplx, w[p][h][l]);
                split_levels25_u<0>(h, StridedConsts{&T.t[2][0][l], 32}, w[p][h][l]);
            }
        }
        ALL split_xp_store(l, h, Tb.data() + p * kSplitTCplx, w[p][h][l]);
        // barrier
        ALL split_pass(h, SplitLoadT{Tb.data() + p * kSplitTCplx + ((32 - l) & 31) * kSplitTRow}, StridedConsts{&T.t[3][0][l], 32}, w[p][h][l]);
        ALL split_tail<AccT>(l, h, acc.data() + p * 1024, tw, w[p][h][l]);
        // barrier
    }
#undef ALL
    for (int j = 0; j <= kN; ++j) out[j] = extract_word<AccT>(acc.data(), acc.data() + 1024, j);
}
}  // namespace

extern "C" {
void emu3_blind_rotate(int acc_bits, int px, int n, int base_log, const double* bsk_f, const uint64_t* cts, int count,
                       const uint64_t* lut, uint64_t* out) {
    const cplx* f = reinterpret_cast<const cplx*>(bsk_f);
    for (int c = 0; c < count; ++c) {
        if (acc_bits == 64) blind_rotate<uint64_t>(px, n, base_log, f, cts + (size_t)c * (n + 1), lut, out + (size_t)c * (kN + 1));
        else blind_rotate<uint32_t>(px, n, base_log, f, cts + (size_t)c * (n + 1), lut, out + (size_t)c * (kN + 1));
    }
}
}
