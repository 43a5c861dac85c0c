// pbs_emu4.cpp — executes fhe_sign_b200/csrc/pbs_core4.cuh (the "quad" blind rotation: four warps per ciphertext, whole
// level-1 butterflies + an 8-value join, accumulator and spectra in tensor memory) lane by lane on the CPU.
// TEST INFRASTRUCTURE: the no-GPU proof of the slot ownership, join, transpose swizzle, tensor-memory layouts and product
// indexing of pbs_quad_kernel.cu; each loop nest below is one region between two barriers of the kernel, the arrays named
// ACC / SPEC / JOIN stand for the tensor-memory columns of one lane quarter.  32-bit accumulator (the kernel's only form).
// The Fourier key comes from emu2_convert_bsk (same layout as the stream kernel).
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../fhe_sign_b200/csrc/pbs_core4.cuh"

using namespace fsc;

namespace {
struct Tables {
    cplx t[4][16][32];          // [pass][ci][lane]
    cplx tw32[32][32];
    Tables() {
        for (int q = 0; q < 4; ++q)
            for (int ci = 0; ci < 16; ++ci)
                for (int l = 0; l < 32; ++l) t[q][ci][l] = pass_const(ci, pass_g(q, l));
        for (int pos = 0; pos < 32; ++pos)
            for (int l = 0; l < 32; ++l) tw32[pos][l] = twist_const<uint32_t>(pos, l);
    }
};
const Tables& tables() { static Tables T; return T; }

typedef uint32_t AccT;
struct OwnFromAcc {      // own-index pair of j2 = 16 b + 8 h + u out of the lane's 64 accumulator columns
    const uint32_t* cols; int h;
    pair_t<AccT> operator()(int b, int u) const {
        const int c = quad_acc_col(16 * b + 8 * h + u);
        pair_t<AccT> z; z.x = cols[c]; z.y = cols[c + 1];
        return z;
    }
};

void blind_rotate(int n, int base_log, const cplx* bsk_f, const uint64_t* ct, const uint64_t* lut, uint64_t* out) {
    const Tables& T = tables();
    std::vector<pair_t<AccT>> scratch(2 * 1024);
    static uint32_t ACC[2][32][64];             // [p][lane][column]
    static cplx SPEC[2][32][32];                // [p][lane][old slot]
    static cplx JOIN[2][2][32][8];              // [p][h][lane][u]: written by warp (p, h), read by warp (p, 1 - h)
    std::vector<cplx> Tb(2 * kQuadTCplx);
    static cplx v[2][2][32][16];                // registers of warp (p, h), lane
    const int b = modswitch(ct[n]);
#define ALL for (int p = 0; p < 2; ++p) for (int h = 0; h < 2; ++h) for (int l = 0; l < 32; ++l)
    ALL for (int k = 0; k < 16; ++k) {           // warp (p, h) initialises the pairs of parity h
        const int j2 = 2 * k + h, idx = l + 32 * j2;
        pair_t<AccT> z; z.x = 0; z.y = 0;
        if (p) z = lut_pair<AccT>(lut, idx, b);
        scratch[p * 1024 + idx] = z;
        ACC[p][l][32 * h + 2 * k] = z.x; ACC[p][l][32 * h + 2 * k + 1] = z.y;
    }
    auto join = [&]() {      // st | pair barrier | ld
        ALL for (int u = 0; u < 8; ++u) JOIN[p][h][l][u] = v[p][h][l][8 * (1 - h) + u];
        ALL for (int u = 0; u < 8; ++u) v[p][h][l][8 * (1 - h) + u] = JOIN[p][1 - h][l][u];
    };
    for (int i = 0; i < n; ++i) {
        const int a = modswitch(ct[i]);
        ALL quad_head<AccT>(l, h, scratch.data() + p * 1024, OwnFromAcc{ACC[p][l], h}, a, base_log, v[p][h][l]);
        ALL quad_level1(StridedConsts{&T.t[0][0][l], 32}, v[p][h][l]);
        join();
        ALL split_levels25(h, StridedConsts{&T.t[0][0][l], 32}, v[p][h][l]);
        ALL quad_xp_store(l, h, Tb.data() + p * kQuadTCplx, v[p][h][l]);
        // pair barrier
        ALL quad_xp_load(l, h, Tb.data() + p * kQuadTCplx, v[p][h][l]);
        ALL quad_level1(StridedConsts{&T.t[1][0][l], 32}, v[p][h][l]);
        join();
        ALL split_levels25(h, StridedConsts{&T.t[1][0][l], 32}, v[p][h][l]);
        ALL for (int jj = 0; jj < 16; ++jj) SPEC[p][l][16 * h + jj] = v[p][h][l][jj];
        // quad barrier
        const cplx* g = bsk_f + (size_t)i * 32 * 4 * 32;
        ALL {
            const QuadKey key{g + l, g + 16 * 4 * 32 + l, 3 * p, 2 - p};
            for (int u = 0; u < 8; ++u)
                for (int bb = 0; bb < 2; ++bb) {
                    const int o = quad_old_slot(u, h) + bb;
                    cplx gw, go;
                    key.load(u, bb, h, gw, go);
                    v[p][h][l][8 * bb + u] = quad_mac(SPEC[p][l][o], SPEC[1 - p][l][o], gw, go);
                }
        }
        ALL quad_level1(StridedConsts{&T.t[2][0][l], 32}, v[p][h][l]);
        join();      // quad barrier in the kernel: every warp has read the spectra
        ALL split_levels25(h, StridedConsts{&T.t[2][0][l], 32}, v[p][h][l]);
        ALL quad_xp_store(l, h, Tb.data() + p * kQuadTCplx, v[p][h][l]);
        // pair barrier
        ALL quad_xp_load((32 - l) & 31, h, Tb.data() + p * kQuadTCplx, v[p][h][l]);
        ALL quad_level1(StridedConsts{&T.t[3][0][l], 32}, v[p][h][l]);
        join();
        ALL split_levels25(h, StridedConsts{&T.t[3][0][l], 32}, v[p][h][l]);
        ALL {
            uint32_t d[32], R[32];
            quad_tail_delta(l, h, &T.tw32[0][0], v[p][h][l], d);
            for (int c = 0; c < 32; ++c) R[c] = ACC[p][l][32 * h + c];
            if (h) quad_tail_add<1>(d, R); else quad_tail_add<0>(d, R);
            for (int c = 0; c < 32; ++c) ACC[p][l][32 * h + c] = R[c];
            for (int k = 0; k < 16; ++k) {      // scratch copy for the rotated reads of the next head (the transpose buffer is free here)
                pair_t<AccT> z; z.x = R[2 * k]; z.y = R[2 * k + 1];
                scratch[p * 1024 + l + 32 * (2 * k + h)] = z;
            }
        }
        // pair barrier
    }
#undef ALL
    for (int j = 0; j <= kN; ++j) out[j] = extract_word<AccT>(scratch.data(), scratch.data() + 1024, j);
}
}  // namespace

extern "C" {
void emu4_blind_rotate(int n, int base_log, const double* bsk_f, const uint64_t* cts, int count, const uint64_t* lut, uint64_t* out) {
    const cplx* f = reinterpret_cast<const cplx*>(bsk_f);
    for (int c = 0; c < count; ++c) blind_rotate(n, base_log, f, cts + (size_t)c * (n + 1), lut, out + (size_t)c * (kN + 1));
}
// conflict check of the swizzled transpose buffer: worst number of lanes of a quarter warp that hit the same 16-byte bank group
int emu4_transpose_conflicts() {
    int worst = 0;
    for (int mode = 0; mode < 3; ++mode)          // 0: store (row fixed, column = lane), 1: forward load (row = lane), 2: inverse load
        for (int fixed = 0; fixed < 32; ++fixed)
            for (int q = 0; q < 4; ++q) {
                int hits[8] = {0};
                for (int i = 0; i < 8; ++i) {
                    const int lane = 8 * q + i;
                    const int idx = mode == 0 ? quad_t_index(fixed, lane) : quad_t_index(mode == 1 ? lane : (32 - lane) & 31, fixed);
                    ++hits[idx & 7];
                }
                for (int k = 0; k < 8; ++k) worst = hits[k] > worst ? hits[k] : worst;
            }
    return worst;
}
}
