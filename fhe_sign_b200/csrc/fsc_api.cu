// fsc_api.cu — context, device buffers and the C ABI of include/fhe_sign_cuda.h.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "fsc_internal.h"
#include "engine.h"
#include "ctx.h"
#include "radix_cuda.h"

namespace fsc {

static thread_local std::string g_create_error;

void ensure_dynamic_smem(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> done;      // (kernel, device) -> bytes opted in
    int dev = 0;
    FSC_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = done[{kernel, dev}];
    if (have >= bytes) return;
    FSC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
}

// ---- LUT polynomial construction (integer host code; what tfhe's generate_lookup_table does) ----
void build_lut_poly(const fsc_params& p, const uint64_t* table, uint64_t* poly) {
    const uint32_t N = p.poly_size, space = p.message_modulus * p.carry_modulus, box = N / space, half = box / 2;
    const uint64_t delta = ((uint64_t)1 << 63) / space;
    std::vector<uint64_t> tmp(N);
    for (uint32_t i = 0; i < space; ++i)
        for (uint32_t j = 0; j < box; ++j) tmp[i * box + j] = table[i] * delta;
    for (uint32_t j = 0; j < N; ++j) {
        const uint32_t s = j + half;
        poly[j] = s < N ? tmp[s] : (uint64_t)0 - tmp[s - N];
    }
}

Engine::Engine(const fsc_params& prm, int device, uintptr_t ext_stream) : p(prm), dev(device) {
    if (p.acc_bits == 0) p.acc_bits = 32;      // measured sigma identical to the 64-bit accumulator, 1.5x the throughput
    if (p.glwe_dim != 1 || p.poly_size != 2048 || p.pbs_level != 1)
        throw Error(FSC_ERR_PARAMS, "kernels are specialised for glwe_dim=1, poly_size=2048, pbs_level=1");
    if (p.acc_bits != 64 && p.acc_bits != 32) throw Error(FSC_ERR_PARAMS, "acc_bits must be 32 or 64");
    if (pbs_variant_for((int)p.acc_bits) >= 5 && pbs_variant_for((int)p.acc_bits) <= 7 && p.acc_bits != 32) throw Error(FSC_ERR_PARAMS, "the solo, quad and duo kernels exist for the 32-bit accumulator only");
    if (p.lwe_dim == 0 || p.lwe_dim > 4096 || p.pbs_base_log < 8 || p.pbs_base_log > 30)
        throw Error(FSC_ERR_PARAMS, "lwe_dim / pbs_base_log out of range");
    // The 32-bit accumulator kernels round through F2I.S64.F64 (pbs_core.cuh to_torus32): the value being rounded is
    // 2^(pbs_base_log + 34.4) rms, so base_log 24 leaves 2^63 at 24 sigma; beyond that the conversion could saturate.
    if (p.acc_bits == 32 && p.pbs_base_log > 24)
        throw Error(FSC_ERR_PARAMS, "acc_bits = 32 supports pbs_base_log <= 24 (use the 64-bit accumulator above that)");
    if (p.ks_level == 0 || p.ks_level > 8 || p.ks_base_log == 0 || p.ks_base_log > 7)
        throw Error(FSC_ERR_PARAMS, "keyswitch decomposition out of range");
    if (p.message_modulus * p.carry_modulus == 0 || (p.poly_size % (p.message_modulus * p.carry_modulus)) != 0)
        throw Error(FSC_ERR_PARAMS, "message/carry modulus must divide poly_size");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw Error(FSC_ERR_CUDA, std::string("no CUDA device available (no CPU fallback exists): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) throw Error(FSC_ERR_BAD_ARG, "device index out of range");
    FSC_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    FSC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) throw Error(FSC_ERR_CUDA, "this library is built for sm_100a (B200) only");
    sm_count = prop.multiProcessorCount;
    if (ext_stream) { stream = reinterpret_cast<cudaStream_t>(ext_stream); own_stream = false; }
    else { FSC_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)); own_stream = true; }
    FSC_CUDA_CHECK(cudaEventCreate(&ev0));
    FSC_CUDA_CHECK(cudaEventCreate(&ev1));
    pbs_init_constants();
    const char* kv = getenv("FSC_KS_VARIANT");      // "simt" | "mma" (default)
    ks_variant = (kv && kv[0] == 's') ? 0 : (kv && kv[0] == 'm') ? 1 : 2;      // simt | mma | umma (default)
    if (const char* wk = getenv("FSC_PBS_WIDE")) wide_tx = wk[0] != 'r';         // "ring": wide batches stay on the ring kernel (comparison)
    if (const char* ns = getenv("FSC_PBS_SPLIT")) use_split = atoi(ns) != 0;      // 0: narrow levels stay on the stream kernel
    if (const char* hw = getenv("FSC_HOST_CHUNK_WAVES")) { host_chunk_waves = (size_t)atoi(hw); host_chunk_explicit = true; }      // 0: one chunk, copies not overlapped
}

void Engine::ensure_copy_streams() {
    if (copy_in) return;
    use();
    FSC_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking));
    FSC_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking));
    for (auto& l : lane) FSC_CUDA_CHECK(cudaStreamCreateWithFlags(&l, cudaStreamNonBlocking));
    FSC_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    FSC_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
}

cudaEvent_t Engine::chunk_event(size_t i) {
    while (chunk_events.size() <= i) {
        cudaEvent_t ev = nullptr;
        FSC_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        chunk_events.push_back(ev);
    }
    return chunk_events[i];
}

Engine::~Engine() {
    cudaSetDevice(dev);
    cudaStreamSynchronize(stream);
    if (bsk_f) cudaFree(bsk_f);
    if (bsk_s) cudaFree(bsk_s);
    if (ksk) cudaFree(ksk);
    if (ksk_limbs) cudaFree(ksk_limbs);
    if (ks_digits) cudaFree(ks_digits);
    if (scratch_small) cudaFree(scratch_small);
    if (scratch_idx) cudaFree(scratch_idx);
    if (scratch_big) cudaFree(scratch_big);
    if (pinned) cudaFreeHost(pinned);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (cudaEvent_t ev : chunk_events) cudaEventDestroy(ev);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (copy_in) { cudaStreamSynchronize(copy_in); cudaStreamDestroy(copy_in); }
    if (copy_out) { cudaStreamSynchronize(copy_out); cudaStreamDestroy(copy_out); }
    for (auto& l : lane) if (l) { cudaStreamSynchronize(l); cudaStreamDestroy(l); }
    if (own_stream && stream) cudaStreamDestroy(stream);
}

void Engine::use() { FSC_CUDA_CHECK(cudaSetDevice(dev)); }

namespace {
// device allocation released on scope exit unless handed over (key upload builds everything in locals first)
struct DevBuf {
    void* p = nullptr;
    DevBuf() {}
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    void alloc(size_t bytes) { FSC_CUDA_CHECK(cudaMalloc(&p, bytes)); }
    template <class T> T* as() const { return static_cast<T*>(p); }
    template <class T> T* release() { T* r = static_cast<T*>(p); p = nullptr; return r; }
};
}  // namespace

// The new key buffers are built in locals and swapped into the engine only after the final stream sync succeeded:
// a failure anywhere (allocation, copy, conversion launch) leaves the engine with NO keys (FSC_ERR_NO_KEYS on the
// next batch call) instead of half-converted ones, and leaks nothing.
void Engine::upload_keys(const uint64_t* bsk_std, size_t bsk_words, const uint64_t* ksk_h, size_t ksk_words) {
    use();
    const size_t n = p.lwe_dim, N = p.poly_size;
    const size_t want_bsk = n * 4 * N, want_ksk = (size_t)N * p.ks_level * (n + 1);
    FSC_REQUIRE(bsk_std && ksk_h, "null key pointer");
    FSC_REQUIRE(bsk_words == want_bsk, "bootstrapping key size does not match the parameter set");
    FSC_REQUIRE(ksk_words == want_ksk, "keyswitching key size does not match the parameter set");
    FSC_CUDA_CHECK(cudaStreamSynchronize(stream));      // nothing in flight may still read the old keys
    if (bsk_f) { cudaFree(bsk_f); bsk_f = nullptr; }
    if (bsk_s) { cudaFree(bsk_s); bsk_s = nullptr; }
    if (ksk) { cudaFree(ksk); ksk = nullptr; }
    if (ksk_limbs) { cudaFree(ksk_limbs); ksk_limbs = nullptr; }
    const size_t fourier_bytes = n * 32 * 4 * 32 * 16;
    const int variant = pbs_variant_for((int)p.acc_bits);
    DevBuf tmp, nf, ns, nk, nl;
    tmp.alloc(want_bsk * 8);
    nf.alloc(fourier_bytes);
    nk.alloc(want_ksk * 8);
    FSC_CUDA_CHECK(cudaMemcpyAsync(tmp.p, bsk_std, want_bsk * 8, cudaMemcpyHostToDevice, stream));
    FSC_CUDA_CHECK(cudaMemcpyAsync(nk.p, ksk_h, want_ksk * 8, cudaMemcpyHostToDevice, stream));
    // Fourier key: correctly rounded direct DFT in double-double arithmetic (bsk_exact.cu: removes the key's share of the
    // floating-point noise, measured 7-11 % of the output variance); FSC_BSK_CONVERT=fft keeps the kernels' own f64 FFT
    // (comparison).  Layouts: ring order for the ring / pair kernels, stream order for the stream / split / solo kernels.
    const bool stream_main = variant == 2 || variant == 4, both = variant == 3 || variant == 5 || variant == 6 || variant == 7 || variant == 8;
    if (both) ns.alloc(fourier_bytes);
    const char* conv = getenv("FSC_BSK_CONVERT");
    if (conv && conv[0] == 'f') {
        if (stream_main) launch_bsk_convert_stream(tmp.as<uint64_t>(), nf.p, (int)n, stream);
        else launch_bsk_convert(tmp.as<uint64_t>(), nf.p, (int)n, stream);
        ++launches;
        FSC_CUDA_CHECK(cudaGetLastError());
        if (both) {
            launch_bsk_convert_stream(tmp.as<uint64_t>(), ns.p, (int)n, stream);
            ++launches;
            FSC_CUDA_CHECK(cudaGetLastError());
        }
    } else {
        launch_bsk_convert_exact(tmp.as<uint64_t>(), stream_main ? nullptr : nf.p, stream_main ? nf.p : ns.p, (int)n, stream);
        ++launches;
        FSC_CUDA_CHECK(cudaGetLastError());
    }
    const size_t K = (size_t)N * p.ks_level;
    nl.alloc(ks_mma_limb_rows((int)n) * K);
    launch_ksk_limb_transpose(nk.as<uint64_t>(), nl.as<uint8_t>(), (int)K, (int)n, stream);
    ++launches;
    FSC_CUDA_CHECK(cudaGetLastError());
    FSC_CUDA_CHECK(cudaStreamSynchronize(stream));
    pbs_variant = variant;
    bsk_f = nf.release<void>();
    bsk_s = ns.release<void>();
    ksk = nk.release<uint64_t>();
    ksk_limbs = nl.release<uint8_t>();
}

void Engine::ensure_scratch(size_t count) {
    if (count <= scratch_cap) return;
    use();
    FSC_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (scratch_small) cudaFree(scratch_small);
    if (scratch_idx) cudaFree(scratch_idx);
    scratch_small = nullptr; scratch_idx = nullptr; scratch_cap = 0;
    size_t cap = 1024;
    while (cap < count) cap *= 2;
    FSC_CUDA_CHECK(cudaMalloc(&scratch_small, cap * (p.lwe_dim + 1) * 8));
    FSC_CUDA_CHECK(cudaMalloc(&scratch_idx, cap * sizeof(uint32_t)));
    scratch_cap = cap;
}

void Engine::ensure_pinned(size_t bytes) {
    if (bytes <= pinned_cap) return;
    use();
    FSC_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (pinned) cudaFreeHost(pinned);
    pinned = nullptr; pinned_cap = 0;
    size_t cap = 1 << 16;
    while (cap < bytes) cap *= 2;
    FSC_CUDA_CHECK(cudaMallocHost(&pinned, cap));
    pinned_cap = cap;
}

const uint32_t* Engine::stage_lut_idx(const uint32_t* lut_idx, size_t count, const Luts* luts) {
    if (!lut_idx) return nullptr;
    for (size_t i = 0; i < count; ++i) FSC_REQUIRE(lut_idx[i] < luts->n, "lut index out of range");
    ensure_scratch(count);
    // the pinned staging area is reused by the next call: wait for earlier copies out of it first
    FSC_CUDA_CHECK(cudaStreamSynchronize(stream));
    ensure_pinned(count * sizeof(uint32_t));
    memcpy(pinned, lut_idx, count * sizeof(uint32_t));
    FSC_CUDA_CHECK(cudaMemcpyAsync(scratch_idx, pinned, count * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    return scratch_idx;
}

void Engine::keyswitch(const uint64_t* in_big, uint64_t* out_small, size_t count) {
    use();
    if (!ksk) throw Error(FSC_ERR_NO_KEYS, "server keys not uploaded");
    if (ks_variant >= 1) {
        const size_t rows = ks_mma_digit_rows(count), K = (size_t)p.poly_size * p.ks_level;
        if (rows > ks_digits_cap) {
            FSC_CUDA_CHECK(cudaStreamSynchronize(stream));
            if (ks_digits) cudaFree(ks_digits);
            ks_digits = nullptr; ks_digits_cap = 0;
            FSC_CUDA_CHECK(cudaMalloc(&ks_digits, rows * K));
            ks_digits_cap = rows;
        }
        if (ks_variant == 2) {
            FSC_REQUIRE(p.ks_base_log <= 7 && p.ks_base_log * p.ks_level < 64, "keyswitch (tensor-core path): unsupported decomposition");
            launch_ks_decompose(in_big, ks_digits, (int)count, (int)rows, (int)p.poly_size, (int)p.ks_base_log, (int)p.ks_level, stream);
            launch_keyswitch_umma(ksk_limbs, ks_mma_limb_rows((int)p.lwe_dim), ks_digits, (int)rows, in_big, out_small, (int)count,
                                  (int)p.poly_size, (int)p.lwe_dim, (int)p.ks_level, stream);
        } else {
            launch_keyswitch_mma(ksk_limbs, ks_digits, in_big, out_small, (int)count, (int)p.poly_size, (int)p.lwe_dim,
                                 (int)p.ks_base_log, (int)p.ks_level, stream);
        }
        launches += 2;
    } else {
        launch_keyswitch(ksk, in_big, out_small, (int)count, (int)p.poly_size, (int)p.lwe_dim, (int)p.ks_base_log,
                         (int)p.ks_level, stream);
        ++launches;
    }
    FSC_CUDA_CHECK(cudaGetLastError());
}

void Engine::pbs(const uint64_t* in_small, const Luts* luts, const uint32_t* lut_idx_dev, uint64_t* out_big, size_t count,
                 const int32_t* out_idx_dev, const OutDest* dests) {
    use();
    if (!bsk_f) throw Error(FSC_ERR_NO_KEYS, "server keys not uploaded");
    const OutDest od = dests ? *dests : single_dest(out_big);
    // variant 3: levels of at most one ciphertext per SM on the split kernel (four warps per ciphertext: the latency-bound
    // case), up to two per SM on the stream kernel, wide batches on the ring kernel
    const bool mixed = pbs_variant == 3 || pbs_variant == 5 || pbs_variant == 6 || pbs_variant == 7 || (pbs_variant == 8 && wide_tx);      // narrow levels on the split / stream kernels, wide ones on ring (3), solo (5) or quad (6)
    if (pbs_variant == 4 || (mixed && use_split && (int)count <= sm_count))
        launch_pbs_split((int)p.acc_bits, pbs_variant == 4 ? bsk_f : bsk_s, in_small, (int)p.lwe_dim, (int)p.pbs_base_log, luts->d,
                         lut_idx_dev, od, out_idx_dev, (int)count, stream);
    else if (mixed && (int)count <= 2 * sm_count)
        launch_pbs_stream((int)p.acc_bits, bsk_s, in_small, (int)p.lwe_dim, (int)p.pbs_base_log, luts->d, lut_idx_dev, od,
                          out_idx_dev, (int)count, sm_count, stream);
    else if (pbs_variant == 5)
        launch_pbs_solo(bsk_s, in_small, (int)p.lwe_dim, (int)p.pbs_base_log, luts->d, lut_idx_dev, od, out_idx_dev, (int)count, stream);
    else if (pbs_variant == 7)
        launch_pbs_duo(bsk_s, in_small, (int)p.lwe_dim, (int)p.pbs_base_log, luts->d, lut_idx_dev, od, out_idx_dev, (int)count, stream);
    else if (pbs_variant == 6)
        launch_pbs_quad(bsk_s, in_small, (int)p.lwe_dim, (int)p.pbs_base_log, luts->d, lut_idx_dev, od, out_idx_dev, (int)count, stream);
    else if (pbs_variant == 2)
        launch_pbs_stream((int)p.acc_bits, bsk_f, in_small, (int)p.lwe_dim, (int)p.pbs_base_log, luts->d, lut_idx_dev, od,
                          out_idx_dev, (int)count, sm_count, stream);
    else if (pbs_variant == 3 && p.acc_bits == 32 && wide_tx)
        // wide batches, 32-bit accumulator: the stream formulation with tensor memory, straight-line trips, the two uniform
        // passes specialised (pbs_stream_tx_kernel<2>: 83.5 k PBS/s against 77.7 k for the ring kernel, profiles/README.md)
        launch_pbs_stream((int)p.acc_bits, bsk_s, in_small, (int)p.lwe_dim, (int)p.pbs_base_log, luts->d, lut_idx_dev, od,
                          out_idx_dev, (int)count, sm_count, stream);
    else if (pbs_variant == 8 && wide_tx && (int)count > 2 * sm_count)
        // wide batches at the reference's accumulator width: the straight-line stream kernel with four ciphertexts per SM
        launch_pbs_stream((int)p.acc_bits, bsk_s, in_small, (int)p.lwe_dim, (int)p.pbs_base_log, luts->d, lut_idx_dev, od,
                          out_idx_dev, (int)count, sm_count, stream);
    else
        launch_pbs(pbs_variant == 8 ? 1 : pbs_variant, (int)p.acc_bits, bsk_f, in_small, (int)p.lwe_dim, (int)p.pbs_base_log, luts->d, lut_idx_dev, od,
                   out_idx_dev, (int)count, sm_count, stream);
    ++launches;
    FSC_CUDA_CHECK(cudaGetLastError());
}

void Engine::ks_pbs(const uint64_t* in_big, const Luts* luts, const uint32_t* lut_idx_dev, uint64_t* out_big, size_t count) {
    ensure_scratch(count);
    keyswitch(in_big, scratch_small, count);
    pbs(scratch_small, luts, lut_idx_dev, out_big, count);
}

}  // namespace fsc

// =======================================================================================
// C ABI
// =======================================================================================
using fsc::Engine;
using fsc::Error;

struct fsc_lwe {
    uint32_t kind;
    size_t count, words;
    uint64_t* d;
};
struct fsc_luts : fsc::Luts {};

#define FSC_API_BEGIN(ctx)                                   \
    if (!(ctx)) return FSC_ERR_BAD_ARG;                      \
    try {
#define FSC_API_END(ctx)                                     \
    }                                                        \
    catch (const Error& e) { (ctx)->err = e.what(); return e.code; }                     \
    catch (const std::bad_alloc&) { (ctx)->err = "host allocation failed"; return FSC_ERR_OOM; } \
    catch (const std::exception& e) { (ctx)->err = e.what(); return FSC_ERR_INTERNAL; }   \
    catch (...) { (ctx)->err = "unknown error"; return FSC_ERR_INTERNAL; }               \
    return FSC_OK;

extern "C" {

fsc_status fsc_map_exception(const std::exception& e) {
    const Error* fe = dynamic_cast<const Error*>(&e);
    return fe ? fe->code : FSC_ERR_INTERNAL;
}

fsc_status fsc_ctx_create(const fsc_params* params, int32_t device, uintptr_t stream, fsc_ctx** out) {
    if (!params || !out) { fsc::g_create_error = "null argument"; return FSC_ERR_BAD_ARG; }
    *out = nullptr;
    fsc_ctx* c = nullptr;
    try {
        c = new fsc_ctx();
        c->eng = new Engine(*params, device, stream);
        c->params = c->eng->p;
        *out = c;
        return FSC_OK;
    } catch (const Error& e) {
        fsc::g_create_error = e.what(); delete c; return e.code;
    } catch (const std::exception& e) {
        fsc::g_create_error = e.what(); delete c; return FSC_ERR_INTERNAL;
    } catch (...) {
        fsc::g_create_error = "unknown error"; delete c; return FSC_ERR_INTERNAL;
    }
}

fsc_status fsc_ctx_destroy(fsc_ctx* ctx) {
    if (!ctx) return FSC_ERR_BAD_ARG;
    try { delete ctx->ev; delete ctx->rb; delete ctx->eng; } catch (...) {}
    delete ctx;
    return FSC_OK;
}

const char* fsc_last_error(const fsc_ctx* ctx) { return ctx ? ctx->err.c_str() : fsc::g_create_error.c_str(); }

fsc_status fsc_get_params(const fsc_ctx* ctx, fsc_params* out) {
    if (!ctx || !out) return FSC_ERR_BAD_ARG;
    *out = ctx->eng->p;
    return FSC_OK;
}

fsc_status fsc_sync(fsc_ctx* ctx) {
    FSC_API_BEGIN(ctx)
    ctx->eng->use();
    FSC_CUDA_CHECK(cudaStreamSynchronize(ctx->eng->stream));
    FSC_API_END(ctx)
}

fsc_status fsc_keys_upload(fsc_ctx* ctx, const uint64_t* bsk_std, size_t bsk_words, const uint64_t* ksk, size_t ksk_words) {
    FSC_API_BEGIN(ctx)
    ctx->eng->upload_keys(bsk_std, bsk_words, ksk, ksk_words);
    if (!ctx->rb) {
        ctx->rb = fsc::make_cuda_backend(ctx->eng);
        ctx->ev = new fsc::Evaluator(ctx->rb);
    }
    FSC_API_END(ctx)
}

fsc_status fsc_lwe_alloc(fsc_ctx* ctx, uint32_t kind, size_t count, fsc_lwe** out) {
    FSC_API_BEGIN(ctx)
    FSC_REQUIRE(out, "null out pointer");
    *out = nullptr;
    FSC_REQUIRE(kind == FSC_LWE_BIG || kind == FSC_LWE_SMALL, "unknown ciphertext kind");
    ctx->eng->use();
    fsc_lwe* a = new fsc_lwe();
    a->kind = kind; a->count = count;
    a->words = kind == FSC_LWE_BIG ? (size_t)ctx->eng->p.glwe_dim * ctx->eng->p.poly_size + 1 : (size_t)ctx->eng->p.lwe_dim + 1;
    a->d = nullptr;
    if (count) {
        cudaError_t e = cudaMalloc(&a->d, count * a->words * 8);
        if (e != cudaSuccess) { delete a; FSC_CUDA_CHECK(e); }
    }
    *out = a;
    FSC_API_END(ctx)
}

fsc_status fsc_lwe_free(fsc_ctx* ctx, fsc_lwe* a) {
    FSC_API_BEGIN(ctx)
    if (a) {
        ctx->eng->use();
        FSC_CUDA_CHECK(cudaStreamSynchronize(ctx->eng->stream));
        if (a->d) cudaFree(a->d);
        delete a;
    }
    FSC_API_END(ctx)
}

fsc_status fsc_lwe_upload(fsc_ctx* ctx, fsc_lwe* dst, size_t first, const uint64_t* host, size_t count) {
    FSC_API_BEGIN(ctx)
    FSC_REQUIRE(dst && (host || !count), "null argument");
    FSC_REQUIRE(first <= dst->count && count <= dst->count - first, "range exceeds the ciphertext array");
    ctx->eng->use();
    if (count) {
        FSC_CUDA_CHECK(cudaMemcpyAsync(dst->d + first * dst->words, host, count * dst->words * 8, cudaMemcpyHostToDevice, ctx->eng->stream));
        FSC_CUDA_CHECK(cudaStreamSynchronize(ctx->eng->stream));
    }
    FSC_API_END(ctx)
}

fsc_status fsc_lwe_download(fsc_ctx* ctx, const fsc_lwe* src, size_t first, uint64_t* host, size_t count) {
    FSC_API_BEGIN(ctx)
    FSC_REQUIRE(src && (host || !count), "null argument");
    FSC_REQUIRE(first <= src->count && count <= src->count - first, "range exceeds the ciphertext array");
    ctx->eng->use();
    if (count)
        FSC_CUDA_CHECK(cudaMemcpyAsync(host, src->d + first * src->words, count * src->words * 8, cudaMemcpyDeviceToHost, ctx->eng->stream));
    FSC_CUDA_CHECK(cudaStreamSynchronize(ctx->eng->stream));
    FSC_API_END(ctx)
}

fsc_status fsc_lwe_info(const fsc_lwe* a, uint32_t* kind, size_t* count, size_t* words, void** device_ptr) {
    if (!a) return FSC_ERR_BAD_ARG;
    if (kind) *kind = a->kind;
    if (count) *count = a->count;
    if (words) *words = a->words;
    if (device_ptr) *device_ptr = a->d;
    return FSC_OK;
}

static fsc_status luts_upload_impl(fsc_ctx* ctx, const uint64_t* polys, size_t n_luts, fsc_luts** out) {
    FSC_API_BEGIN(ctx)
    FSC_REQUIRE(polys && out && n_luts, "null argument or zero LUTs");
    *out = nullptr;
    ctx->eng->use();
    fsc_luts* l = new fsc_luts();
    l->n = n_luts; l->d = nullptr;
    const size_t bytes = n_luts * ctx->eng->p.poly_size * 8;
    cudaError_t e = cudaMalloc(&l->d, bytes);
    if (e != cudaSuccess) { delete l; FSC_CUDA_CHECK(e); }
    e = cudaMemcpyAsync(l->d, polys, bytes, cudaMemcpyHostToDevice, ctx->eng->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->eng->stream);
    if (e != cudaSuccess) { cudaFree(l->d); delete l; FSC_CUDA_CHECK(e); }
    *out = l;
    FSC_API_END(ctx)
}

fsc_status fsc_luts_upload(fsc_ctx* ctx, const uint64_t* polys, size_t n_luts, fsc_luts** out) {
    return luts_upload_impl(ctx, polys, n_luts, out);
}

fsc_status fsc_luts_from_tables(fsc_ctx* ctx, const uint64_t* tables, size_t n_luts, fsc_luts** out) {
    if (!ctx) return FSC_ERR_BAD_ARG;
    if (!tables || !out || !n_luts) { ctx->err = "null argument or zero LUTs"; return FSC_ERR_BAD_ARG; }
    std::vector<uint64_t> polys;
    try {
        const fsc_params& p = ctx->eng->p;
        const size_t space = p.message_modulus * p.carry_modulus;
        polys.resize(n_luts * p.poly_size);
        for (size_t i = 0; i < n_luts; ++i) fsc::build_lut_poly(p, tables + i * space, polys.data() + i * p.poly_size);
    } catch (...) { ctx->err = "host allocation failed"; return FSC_ERR_OOM; }
    return luts_upload_impl(ctx, polys.data(), n_luts, out);
}

fsc_status fsc_luts_free(fsc_ctx* ctx, fsc_luts* l) {
    FSC_API_BEGIN(ctx)
    if (l) {
        ctx->eng->use();
        FSC_CUDA_CHECK(cudaStreamSynchronize(ctx->eng->stream));
        if (l->d) cudaFree(l->d);
        delete l;
    }
    FSC_API_END(ctx)
}

static void check_range(const fsc_lwe* a, uint32_t kind, size_t first, size_t count, const char* what) {
    if (!a) throw Error(FSC_ERR_BAD_ARG, std::string(what) + ": null ciphertext array");
    if (a->kind != kind) throw Error(FSC_ERR_BAD_ARG, std::string(what) + ": wrong ciphertext kind");
    if (first > a->count || count > a->count - first) throw Error(FSC_ERR_BAD_ARG, std::string(what) + ": range exceeds the ciphertext array");
}

fsc_status fsc_keyswitch_batch(fsc_ctx* ctx, const fsc_lwe* in_big, size_t in_first, fsc_lwe* out_small, size_t out_first, size_t count) {
    FSC_API_BEGIN(ctx)
    check_range(in_big, FSC_LWE_BIG, in_first, count, "keyswitch input");
    check_range(out_small, FSC_LWE_SMALL, out_first, count, "keyswitch output");
    if (count) ctx->eng->keyswitch(in_big->d + in_first * in_big->words, out_small->d + out_first * out_small->words, count);
    FSC_API_END(ctx)
}

fsc_status fsc_pbs_batch(fsc_ctx* ctx, const fsc_lwe* in_small, size_t in_first, const fsc_luts* luts, const uint32_t* lut_idx,
                         fsc_lwe* out_big, size_t out_first, size_t count) {
    FSC_API_BEGIN(ctx)
    FSC_REQUIRE(luts, "null LUT handle");
    check_range(in_small, FSC_LWE_SMALL, in_first, count, "pbs input");
    check_range(out_big, FSC_LWE_BIG, out_first, count, "pbs output");
    if (count) {
        const uint32_t* di = ctx->eng->stage_lut_idx(lut_idx, count, luts);
        ctx->eng->pbs(in_small->d + in_first * in_small->words, luts, di, out_big->d + out_first * out_big->words, count);
    }
    FSC_API_END(ctx)
}

fsc_status fsc_ks_pbs_batch(fsc_ctx* ctx, const fsc_lwe* in_big, size_t in_first, const fsc_luts* luts, const uint32_t* lut_idx,
                            fsc_lwe* out_big, size_t out_first, size_t count) {
    FSC_API_BEGIN(ctx)
    FSC_REQUIRE(luts, "null LUT handle");
    check_range(in_big, FSC_LWE_BIG, in_first, count, "ks_pbs input");
    check_range(out_big, FSC_LWE_BIG, out_first, count, "ks_pbs output");
    if (count) {
        const uint32_t* di = ctx->eng->stage_lut_idx(lut_idx, count, luts);
        ctx->eng->ks_pbs(in_big->d + in_first * in_big->words, luts, di, out_big->d + out_first * out_big->words, count);
    }
    FSC_API_END(ctx)
}

fsc_status fsc_apply_lut_host(fsc_ctx* ctx, const uint64_t* in_host, const fsc_luts* luts, const uint32_t* lut_idx,
                              uint64_t* out_host, size_t count) {
    FSC_API_BEGIN(ctx)
    FSC_REQUIRE(luts && (in_host || !count) && (out_host || !count), "null argument");
    if (count) {
        Engine* e = ctx->eng;
        e->use();
        const size_t words = (size_t)e->p.glwe_dim * e->p.poly_size + 1;
        if (count > e->scratch_big_cap) {
            FSC_CUDA_CHECK(cudaStreamSynchronize(e->stream));
            if (e->scratch_big) cudaFree(e->scratch_big);
            e->scratch_big = nullptr; e->scratch_big_cap = 0;
            FSC_CUDA_CHECK(cudaMalloc(&e->scratch_big, 2 * count * words * 8));
            e->scratch_big_cap = count;
        }
        uint64_t* din = e->scratch_big;
        uint64_t* dout = e->scratch_big + e->scratch_big_cap * words;
        const uint32_t* di = e->stage_lut_idx(lut_idx, count, luts);
        // Copies overlap the bootstraps: the batch is cut at whole waves of the blind-rotation kernel (4 or 3
        // ciphertexts per SM), chunk c + 1 uploads and chunk c - 1 downloads on two copy streams while chunk c computes, so only
        // the first upload and the last download are exposed.
        const size_t wave = (size_t)e->sm_count * (e->p.acc_bits == 32 ? 4 : 3);
        const size_t chunk = e->host_chunk_waves ? e->host_chunk_waves * wave : count;
        if (count <= chunk + wave / 2) {
            FSC_CUDA_CHECK(cudaMemcpyAsync(din, in_host, count * words * 8, cudaMemcpyHostToDevice, e->stream));
            e->ks_pbs(din, luts, di, dout, count);
            FSC_CUDA_CHECK(cudaMemcpyAsync(out_host, dout, count * words * 8, cudaMemcpyDeviceToHost, e->stream));
        } else {
            e->ensure_copy_streams();
            e->ensure_scratch(count);
            const size_t nsmall = (size_t)e->p.lwe_dim + 1;
            FSC_CUDA_CHECK(cudaEventRecord(e->ev_fork, e->stream));          // staging buffers are free once earlier work is done
            FSC_CUDA_CHECK(cudaStreamWaitEvent(e->copy_in, e->ev_fork, 0));
            FSC_CUDA_CHECK(cudaStreamWaitEvent(e->copy_out, e->ev_fork, 0));
            // chunk plan.  Default: first wave | everything between, as ONE launch | last (partial) wave — every launch boundary is a
            // barrier of the whole GPU on the slowest SM of the wave before it (7 one-wave launches of a 4096-block batch cost 2.8 ms
            // against 3 launches), while the uploads stream back to back and only the first upload (one wave) and the last download
            // stay exposed.  FSC_HOST_CHUNK_WAVES=k: uniform chunks of k waves (comparison).
            std::vector<std::pair<size_t, size_t>> plan;
            if (!e->host_chunk_explicit && count > 2 * wave + wave / 2) {
                size_t last = (count - wave) % wave;
                if (last == 0) last = wave;
                plan.emplace_back(0, wave);
                plan.emplace_back(wave, count - wave - last);
                plan.emplace_back(count - last, last);
            } else {
                for (size_t off = 0; off < count; off += chunk) plan.emplace_back(off, std::min(chunk, count - off));
            }
            size_t k = 0;
            for (const auto& pc : plan) {
                const size_t off = pc.first, c = pc.second;
                ++k;
                cudaEvent_t up = e->chunk_event(3 * (k - 1)), done = e->chunk_event(3 * (k - 1) + 1), ksd = e->chunk_event(3 * (k - 1) + 2);
                FSC_CUDA_CHECK(cudaMemcpyAsync(din + off * words, in_host + off * words, c * words * 8, cudaMemcpyHostToDevice, e->copy_in));
                FSC_CUDA_CHECK(cudaEventRecord(up, e->copy_in));
                FSC_CUDA_CHECK(cudaStreamWaitEvent(e->stream, up, 0));
                e->keyswitch(din + off * words, e->scratch_small + off * nsmall, c);      // keyswitches stay in order on the context's stream (one digit scratch)
                FSC_CUDA_CHECK(cudaEventRecord(ksd, e->stream));
                // the blind rotations of consecutive chunks go to two alternating streams: they touch disjoint ciphertexts, and without
                // stream order between them the CTAs of chunk k + 1 start as the SMs of chunk k drain instead of after its slowest one
                cudaStream_t main_stream = e->stream;
                cudaStream_t ls = e->host_chunk_explicit ? main_stream : e->lane[k & 1];
                if (ls != main_stream) { FSC_CUDA_CHECK(cudaStreamWaitEvent(ls, ksd, 0)); e->stream = ls; }
                try {
                    e->pbs(e->scratch_small + off * nsmall, luts, di ? di + off : nullptr, dout + off * words, c);
                } catch (...) { e->stream = main_stream; throw; }
                e->stream = main_stream;
                FSC_CUDA_CHECK(cudaEventRecord(done, ls));
                FSC_CUDA_CHECK(cudaStreamWaitEvent(e->copy_out, done, 0));
                FSC_CUDA_CHECK(cudaMemcpyAsync(out_host + off * words, dout + off * words, c * words * 8, cudaMemcpyDeviceToHost, e->copy_out));
            }
            FSC_CUDA_CHECK(cudaEventRecord(e->ev_join, e->copy_out));
            FSC_CUDA_CHECK(cudaStreamWaitEvent(e->stream, e->ev_join, 0));      // later work on the context's stream stays ordered
        }
        FSC_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    }
    FSC_API_END(ctx)
}

fsc_status fsc_timer_start(fsc_ctx* ctx) {
    FSC_API_BEGIN(ctx)
    ctx->eng->use();
    FSC_CUDA_CHECK(cudaEventRecord(ctx->eng->ev0, ctx->eng->stream));
    FSC_API_END(ctx)
}

fsc_status fsc_timer_stop(fsc_ctx* ctx, float* ms) {
    FSC_API_BEGIN(ctx)
    FSC_REQUIRE(ms, "null argument");
    ctx->eng->use();
    FSC_CUDA_CHECK(cudaEventRecord(ctx->eng->ev1, ctx->eng->stream));
    FSC_CUDA_CHECK(cudaEventSynchronize(ctx->eng->ev1));
    FSC_CUDA_CHECK(cudaEventElapsedTime(ms, ctx->eng->ev0, ctx->eng->ev1));
    FSC_API_END(ctx)
}

fsc_status fsc_launch_count(const fsc_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return FSC_ERR_BAD_ARG;
    *out = ctx->eng->launches;
    return FSC_OK;
}

const char* fsc_pbs_kernel_name(const fsc_ctx* ctx) {
    if (!ctx) return "";
    const int v = ctx->eng->bsk_f ? ctx->eng->pbs_variant : fsc::pbs_variant_for((int)ctx->eng->p.acc_bits);
    if ((v == 3 && ctx->eng->p.acc_bits == 32 && ctx->eng->wide_tx) || (v == 8 && ctx->eng->wide_tx)) return "pbs_stream_tx_kernel";
    return v == 7 ? "pbs_duo_kernel" : v == 6 ? "pbs_quad_kernel" : v == 5 ? "pbs_solo_kernel" : v == 4 ? "pbs_split_kernel" : v == 2 ? "pbs_stream_kernel" : v == 0 ? "pbs_pair_kernel" : "pbs_ring_kernel";
}

fsc_status fsc_measure_fp64_peak(fsc_ctx* ctx, double* tflops) {
    FSC_API_BEGIN(ctx)
    FSC_REQUIRE(tflops, "null argument");
    Engine* e = ctx->eng;
    e->use();
    double* sink = nullptr;
    FSC_CUDA_CHECK(cudaMalloc(&sink, 8));
    float best = 1e30f;
    double fmas = 0;
    cudaError_t err = cudaSuccess;
    for (int rep = 0; rep < 4 && err == cudaSuccess; ++rep) {          // first repetition warms up
        err = cudaEventRecord(e->ev0, e->stream);
        fmas = fsc::launch_fp64_peak(sink, e->sm_count, 2048, e->stream); ++e->launches;
        if (err == cudaSuccess) err = cudaGetLastError();
        if (err == cudaSuccess) err = cudaEventRecord(e->ev1, e->stream);
        if (err == cudaSuccess) err = cudaEventSynchronize(e->ev1);
        float ms = 0;
        if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e->ev0, e->ev1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(sink);
    FSC_CUDA_CHECK(err);
    *tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
    FSC_API_END(ctx)
}

fsc_status fsc_debug_fourier_key(fsc_ctx* ctx, double* out, int32_t stream_order) {
    FSC_API_BEGIN(ctx)
    FSC_REQUIRE(out, "null argument");
    Engine* e = ctx->eng;
    e->use();
    if (!e->bsk_f) throw Error(FSC_ERR_NO_KEYS, "server keys not uploaded");
    const bool main_is_stream = e->pbs_variant == 2 || e->pbs_variant == 4;
    const void* src = (stream_order != 0) == main_is_stream ? e->bsk_f : (stream_order ? e->bsk_s : nullptr);
    FSC_REQUIRE(src, "this context does not hold the Fourier key in that layout");
    const size_t bytes = (size_t)e->p.lwe_dim * 32 * 4 * 32 * 16;
    FSC_CUDA_CHECK(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, e->stream));
    FSC_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    FSC_API_END(ctx)
}

fsc_status fsc_debug_negacyclic_mul(fsc_ctx* ctx, const uint64_t* a, const int64_t* b, uint64_t* c, size_t count) {
    FSC_API_BEGIN(ctx)
    FSC_REQUIRE(a && b && c && count, "null argument");
    Engine* e = ctx->eng;
    e->use();
    const size_t bytes = count * e->p.poly_size * 8;
    uint64_t* d = nullptr;
    FSC_CUDA_CHECK(cudaMalloc(&d, 3 * bytes));
    cudaError_t err = cudaMemcpyAsync(d, a, bytes, cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(d + bytes / 8, b, bytes, cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) {
        if (fsc::pbs_variant_for((int)e->p.acc_bits) == 2)
            fsc::launch_negacyclic_mul_stream(d, reinterpret_cast<const int64_t*>(d + bytes / 8), d + 2 * (bytes / 8), (int)count, e->stream);
        else
            fsc::launch_negacyclic_mul(d, reinterpret_cast<const int64_t*>(d + bytes / 8), d + 2 * (bytes / 8), (int)count, e->stream);
        ++e->launches;
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaMemcpyAsync(c, d + 2 * (bytes / 8), bytes, cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    cudaFree(d);
    FSC_CUDA_CHECK(err);
    FSC_API_END(ctx)
}

}  // extern "C"
