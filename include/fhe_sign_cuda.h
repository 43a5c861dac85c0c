/*
 * fhe_sign_cuda.h — C ABI of libfhe_sign_cuda.so, the B200 (sm_100a) server-side engine that
 * replaces the tfhe-rs CPU server key behind fhe-sign's BigUintFHE.
 *
 * Drop-in seam.  The reference installs a thread-local CPU server key with
 * `tfhe::set_server_key(server_keys)` (src/biguint.rs:278, src/schnorr.rs:472, src/perf_test.rs:24)
 * and every `FheUint32/FheUint64` operator (src/biguint.rs:110,116,135-143,221-248;
 * src/perf_test.rs:28-54) then runs keyswitch + programmable bootstrap on the CPU.  The reference has
 * no FFI of its own; these entry points are what a `fhe-sign-cuda` Rust crate binds with
 * `extern "C"` so that biguint.rs keeps its API.
 *
 * Conventions
 *  - every function returns an fsc_status (0 = OK); no C++ exception crosses this boundary;
 *    fsc_last_error() gives the message of the last failure on that context (or of the last failed
 *    fsc_ctx_create when ctx == NULL);
 *  - handles are opaque and owned by the library; host buffers passed in are read or written during
 *    the call only and are never freed by the library;
 *  - a context is not thread-safe (it mirrors the thread-local set_server_key): one per host thread;
 *    work is enqueued on the context's CUDA stream, in call order; fsc_sync() or any download waits;
 *  - ciphertexts are LWE vectors of u64 words, mask first, body last: "big" = k*N+1 = 2049 words
 *    (the radix blocks the reference's FheUint types are made of), "small" = n+1 words;
 *  - there is no CPU fallback: if no usable CUDA device exists fsc_ctx_create fails with FSC_ERR_CUDA.
 */
#ifndef FHE_SIGN_CUDA_H
#define FHE_SIGN_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t fsc_status;
enum {
    FSC_OK = 0,
    FSC_ERR_BAD_ARG = 1,     /* null pointer, size mismatch, out-of-range index                 */
    FSC_ERR_PARAMS = 2,      /* parameter set not supported by the kernels (k=1, N=2048, l=1)   */
    FSC_ERR_OOM = 3,         /* device or host allocation failed                                */
    FSC_ERR_CUDA = 4,        /* CUDA runtime error / no device                                  */
    FSC_ERR_NO_KEYS = 5,     /* server keys not uploaded yet                                    */
    FSC_ERR_COMM = 6,        /* level exchange failed: callback error or a peer missed the barrier */
    FSC_ERR_INTERNAL = 7
};

/* Replaces tfhe::ConfigBuilder::default().build() (src/biguint.rs:276, src/perf_test.rs:9,
 * src/schnorr.rs:441): the PARAM_MESSAGE_2_CARRY_2_KS_PBS constants as plain data.            */
typedef struct {
    uint32_t lwe_dim;          /* n                                                   */
    uint32_t glwe_dim;         /* k  (kernels require 1)                              */
    uint32_t poly_size;        /* N  (kernels require 2048)                           */
    uint32_t pbs_base_log;     /* log2 beta_pbs                                       */
    uint32_t pbs_level;        /* l_pbs (kernels require 1)                           */
    uint32_t ks_base_log;      /* log2 beta_ks                                        */
    uint32_t ks_level;         /* l_ks                                                */
    uint32_t message_modulus;  /* 4                                                   */
    uint32_t carry_modulus;    /* 4                                                   */
    uint32_t acc_bits;         /* blind-rotation accumulator width: 32 (default if 0: adds variance 2^-60 to a PBS
                                * variance of 2^-30, measured sigma identical, 1.9x the throughput; needs
                                * pbs_base_log <= 24: its kernels round through F2I.S64) or 64               */
} fsc_params;

typedef struct fsc_ctx fsc_ctx;          /* replaces the thread-local tfhe ServerKey              */
typedef struct fsc_lwe fsc_lwe;          /* device array of LWE ciphertexts                       */
typedef struct fsc_luts fsc_luts;        /* device array of LUT accumulator polynomials           */

/* ---- context & keys ------------------------------------------------------------------ */
/* `stream` is an existing cudaStream_t cast to uintptr_t (e.g. torch.cuda.current_stream().cuda_stream)
 * or 0 to let the context create its own non-blocking stream.                                    */
fsc_status fsc_ctx_create(const fsc_params *params, int32_t device, uintptr_t stream, fsc_ctx **out);
fsc_status fsc_ctx_destroy(fsc_ctx *ctx);
const char *fsc_last_error(const fsc_ctx *ctx);
fsc_status fsc_get_params(const fsc_ctx *ctx, fsc_params *out);
fsc_status fsc_sync(fsc_ctx *ctx);

/* Replaces tfhe::set_server_key (src/biguint.rs:278).  bsk_std: [n][k+1][l][k+1][N] standard-domain
 * GGSW rows; ksk: [k*N][l_ks][n+1].  The bootstrapping key is converted to the Fourier domain on
 * the device; the host buffers may be released when the call returns.                            */
fsc_status fsc_keys_upload(fsc_ctx *ctx, const uint64_t *bsk_std, size_t bsk_words,
                           const uint64_t *ksk, size_t ksk_words);

/* ---- ciphertext buffers --------------------------------------------------------------- */
enum { FSC_LWE_BIG = 0, FSC_LWE_SMALL = 1 };
fsc_status fsc_lwe_alloc(fsc_ctx *ctx, uint32_t kind, size_t count, fsc_lwe **out);
fsc_status fsc_lwe_free(fsc_ctx *ctx, fsc_lwe *a);
fsc_status fsc_lwe_upload(fsc_ctx *ctx, fsc_lwe *dst, size_t first, const uint64_t *host, size_t count);
fsc_status fsc_lwe_download(fsc_ctx *ctx, const fsc_lwe *src, size_t first, uint64_t *host, size_t count);
fsc_status fsc_lwe_info(const fsc_lwe *a, uint32_t *kind, size_t *count, size_t *words_per_ct, void **device_ptr);

/* ---- lookup tables --------------------------------------------------------------------- */
/* tables: n_luts x (message_modulus*carry_modulus) function values f(0..15); the library builds the
 * redundant, half-box-rotated accumulator polynomials (what tfhe's generate_lookup_table does).  */
fsc_status fsc_luts_from_tables(fsc_ctx *ctx, const uint64_t *tables, size_t n_luts, fsc_luts **out);
/* polys: n_luts x N torus coefficients, used as given.                                           */
fsc_status fsc_luts_upload(fsc_ctx *ctx, const uint64_t *polys, size_t n_luts, fsc_luts **out);
fsc_status fsc_luts_free(fsc_ctx *ctx, fsc_luts *l);

/* ---- the hot path ---------------------------------------------------------------------- */
/* lut_idx: host array of `count` indices into `luts`, or NULL for LUT 0 everywhere.
 * in/out ranges: ciphertexts [in_first, in_first+count) -> [out_first, out_first+count).        */
fsc_status fsc_keyswitch_batch(fsc_ctx *ctx, const fsc_lwe *in_big, size_t in_first,
                               fsc_lwe *out_small, size_t out_first, size_t count);
fsc_status fsc_pbs_batch(fsc_ctx *ctx, const fsc_lwe *in_small, size_t in_first, const fsc_luts *luts,
                         const uint32_t *lut_idx, fsc_lwe *out_big, size_t out_first, size_t count);
/* shortint apply_lookup_table over a batch: keyswitch then PBS then sample extraction.           */
fsc_status fsc_ks_pbs_batch(fsc_ctx *ctx, const fsc_lwe *in_big, size_t in_first, const fsc_luts *luts,
                            const uint32_t *lut_idx, fsc_lwe *out_big, size_t out_first, size_t count);

/* Host-buffer convenience (the end-to-end call bench.py times as `e2e`): upload `count` big
 * ciphertexts, keyswitch + PBS, download the results.  Synchronous: returns when out_big_host is
 * complete.  Batches of more than one and a half kernel waves are cut at whole waves; uploads and
 * downloads of neighbouring chunks run on two copy streams beside the bootstraps (pinned host
 * buffers make the copies asynchronous; pageable ones still work).                               */
fsc_status fsc_apply_lut_host(fsc_ctx *ctx, const uint64_t *in_big_host, const fsc_luts *luts,
                              const uint32_t *lut_idx, uint64_t *out_big_host, size_t count);

/* ---- radix integers: the FheUint8/32/64 operator surface of the reference ------------------ */
/* A radix value is a little-endian vector of big-LWE blocks carrying 2 message bits each
 * (FheUint8/32/64 = 4/16/32 blocks; a 256-bit integer = 128 blocks).  Values live on the device
 * between operators; every operator returns a NEW handle and never modifies its inputs (the
 * reference clones every operand: src/biguint.rs:135-137,221-222).  Wrapping semantics and the
 * shift-amount-modulo-width rule are those of the reference's tfhe types (src/biguint.rs:469-499). */
typedef struct fsc_radix fsc_radix;

enum {
    FSC_OP_ADD = 0,   /* +   src/biguint.rs:138,148,236,243,248 ; src/perf_test.rs:28 ; scalar: src/schnorr.rs:604 */
    FSC_OP_SUB = 1,
    FSC_OP_MUL = 2,   /* *   src/biguint.rs:223 ; src/perf_test.rs:32 ; scalar: src/schnorr.rs:588              */
    FSC_OP_MIN = 3,   /* FheOrd::min   src/perf_test.rs:44                                                      */
    FSC_OP_MAX = 4,
    FSC_OP_SHR = 5,   /* >>  encrypted amount: src/perf_test.rs:36 ; scalar amount: src/biguint.rs:110,141      */
    FSC_OP_SHL = 6,
    FSC_OP_AND = 7,   /* &   scalar mask: src/biguint.rs:116,143 ; src/perf_test.rs:48                          */
    FSC_OP_OR = 8,
    FSC_OP_XOR = 9,
    FSC_OP_LT = 10,   /* result: one block holding 0 / 1                                                         */
    FSC_OP_EQ = 11,
    FSC_OP_DIV = 12,  /* /   scalar divisor: src/perf_test.rs:54 (division by zero -> FSC_ERR_BAD_ARG)          */
    FSC_OP_REM = 13
};

/* FheUintN::try_encrypt happens on the client (src/biguint.rs:26); the server receives the blocks.
 * host_blocks: n_blocks * (k*N+1) words, least significant block first.                          */
fsc_status fsc_radix_from_lwe(fsc_ctx *ctx, const uint64_t *host_blocks, size_t n_blocks, fsc_radix **out);
/* Brings the blocks back for FheUintN::decrypt (src/biguint.rs:70); synchronises.                */
fsc_status fsc_radix_to_lwe(fsc_ctx *ctx, fsc_radix *r, uint64_t *host_blocks);
/* Noiseless encryption of a plaintext constant (little-endian bytes), e.g. the zero digits of Mul. */
fsc_status fsc_radix_trivial(fsc_ctx *ctx, const uint8_t *value_le, size_t n_bytes, size_t n_blocks, fsc_radix **out);
fsc_status fsc_radix_clone(fsc_ctx *ctx, const fsc_radix *a, fsc_radix **out);
fsc_status fsc_radix_free(fsc_ctx *ctx, fsc_radix *a);
fsc_status fsc_radix_len(const fsc_radix *a, size_t *n_blocks);
/* out = a <op> b for FSC_OP_ADD..FSC_OP_EQ (ciphertext, ciphertext).                              */
fsc_status fsc_radix_binary(fsc_ctx *ctx, uint32_t op, const fsc_radix *a, const fsc_radix *b, fsc_radix **out);
/* out = a <op> scalar for ADD, MUL, AND, DIV, REM, SHR, SHL; the scalar is little-endian bytes of any
 * length (for SHR/SHL it is the bit count, taken modulo the width of a).                          */
fsc_status fsc_radix_scalar(fsc_ctx *ctx, uint32_t op, const fsc_radix *a, const uint8_t *scalar_le, size_t n_bytes, fsc_radix **out);
/* a * b without wrapping at |a| blocks: out_blocks result blocks (e.g. 256 for a 256x256-bit product). */
fsc_status fsc_radix_mul_wide(fsc_ctx *ctx, const fsc_radix *a, const fsc_radix *b, size_t out_blocks, fsc_radix **out);
/* a * b + addend, out_blocks result blocks: the addend rides in the product's column sum, one carry propagation for
 * both - the fused form of `k_fhe + (e_fhe * privkey_fhe)` (src/schnorr.rs:274).                  */
fsc_status fsc_radix_mul_add_wide(fsc_ctx *ctx, const fsc_radix *a, const fsc_radix *b, const fsc_radix *addend,
                                  size_t out_blocks, fsc_radix **out);
/* a * scalar + addend, out_blocks result blocks, scalar = little-endian bytes: the product's digit multiples cost two lookups per
 * block and the addend rides in the column sum.  Use: k + e*d with the challenge e taken in PLAINTEXT - e is public by
 * construction (every verifier recomputes it from R, P and the message: src/schnorr.rs:267, :307-345), only d and k need to stay
 * encrypted.  This is not the reference's dataflow (it encrypts e too, src/schnorr.rs:272); 3.5x fewer bootstraps.          */
fsc_status fsc_radix_scalar_mul_add_wide(fsc_ctx *ctx, const fsc_radix *a, const uint8_t *scalar_le, size_t n_bytes,
                                         const fsc_radix *addend, size_t out_blocks, fsc_radix **out);
/* FheUintM::cast_from (src/biguint.rs:110,116,135-137): truncate or zero-extend; no device work.  */
fsc_status fsc_radix_cast(fsc_ctx *ctx, const fsc_radix *a, size_t n_blocks, fsc_radix **out);
fsc_status fsc_radix_slice(fsc_ctx *ctx, const fsc_radix *a, size_t first, size_t n_blocks, fsc_radix **out);
fsc_status fsc_radix_concat(fsc_ctx *ctx, const fsc_radix *const *parts, size_t n_parts, fsc_radix **out);
/* sum of n_operands values, wrapping at n_blocks (one carry-save reduction + one carry propagation). */
fsc_status fsc_radix_sum(fsc_ctx *ctx, const fsc_radix *const *operands, size_t n_operands, size_t n_blocks, fsc_radix **out);
/* cond ? if_true : if_false; cond is a one-block value holding 0 / 1 (FSC_OP_LT / FSC_OP_EQ output). */
fsc_status fsc_radix_select(fsc_ctx *ctx, const fsc_radix *cond, const fsc_radix *if_true, const fsc_radix *if_false, fsc_radix **out);
/* ---- multi-GPU: sharding of wide PBS levels over the ranks of one node ------------------------ */
/* One process per GPU, keys replicated, every rank runs the same operator sequence.  A PBS level of at least
 * `min_width` requests is cut into `world` contiguous slices; this rank bootstraps slice `rank` into
 * buffer[rank * bytes_per_rank ...] and then calls `all_gather`, which must enqueue on the context's stream an
 * all-gather over that buffer (rank r's slice at offset r * bytes_per_rank on every rank), e.g. NCCL over
 * NVLink via torch.distributed.  `buffer` is caller-owned device memory of `capacity_bytes`.  Narrower levels
 * run replicated on every rank with no exchange.  (The reference has no counterpart: rayon threads, one host.) */
typedef int32_t (*fsc_exchange_fn)(void *user, void *buffer, size_t bytes_per_rank);
fsc_status fsc_set_level_exchange(fsc_ctx *ctx, int32_t rank, int32_t world, size_t min_width, void *buffer,
                                  size_t capacity_bytes, fsc_exchange_fn all_gather, void *user);
/* The same sharding with the exchange OWNED BY THE LIBRARY and fused into the blind-rotation kernel (no callback, no
 * NCCL, nothing but these three calls for a Rust host): every rank's block pool is a fixed-capacity allocation mapped
 * into its peers (CUDA IPC between processes, cudaDeviceEnablePeerAccess inside one process); the sample-extraction
 * epilogue of rank r's slice stores each output word straight into the destination slot of EVERY rank's pool over
 * NVLink, and a flag barrier in peer memory (one tiny kernel; st.release.sys / ld.acquire.sys, bounded spin) orders the
 * levels.  No staging buffer, no all-gather, no scatter launch.
 *   1. every rank: fsc_peer_pool_export(ctx, capacity_blocks, handle)       -> 128 opaque bytes
 *   2. the host exchanges the `world` handles by any means (a file, a socket, torch.distributed ...)
 *   3. every rank: fsc_peer_pool_connect(ctx, rank, world, min_width, all_handles)   (handles in rank order)
 * Contract as above: SPMD (identical operator sequence and identical input ciphertexts on every rank), keys
 * replicated; the pool cannot grow past capacity_blocks once exported (FSC_ERR_OOM).  A peer that never arrives
 * makes the barrier give up after a few seconds; the next download reports FSC_ERR_COMM.  Handles are TRUSTED input (they carry
 * device addresses): exchange them only between the ranks of one job.                                                     */
#define FSC_PEER_HANDLE_BYTES 128
fsc_status fsc_peer_pool_export(fsc_ctx *ctx, size_t capacity_blocks, uint8_t *handle_out);
fsc_status fsc_peer_pool_connect(fsc_ctx *ctx, int32_t rank, int32_t world, size_t min_width, const uint8_t *handles);
/* Unmaps the peers' pools and turns sharding off.  All ranks call it together: a host-side barrier before (nobody still writes
 * into a pool) and after (nobody still maps a pool that is about to be reallocated).                                   */
fsc_status fsc_peer_pool_disconnect(fsc_ctx *ctx);
/* bootstraps, PBS levels and sharded levels issued by the radix layer so far on this context     */
fsc_status fsc_radix_stats2(const fsc_ctx *ctx, uint64_t *pbs_count, uint64_t *level_count, uint64_t *sharded_levels);
/* bootstraps and PBS levels issued by the radix layer so far on this context                      */
fsc_status fsc_radix_stats(const fsc_ctx *ctx, uint64_t *pbs_count, uint64_t *level_count);

/* ---- client side (host CPU): key generation, block encryption, decryption ------------------- */
/* Replaces tfhe::generate_keys / FheUintN::try_encrypt / decrypt (src/biguint.rs:26,70,277).  Plain host code.
 * Randomness: ChaCha20 under a 256-bit master key from the OS CSPRNG (getrandom; the reference's tfhe is built with
 * seeder_unix, Cargo.toml:9) for the keys, and under a second OS-drawn 256-bit key per client instance for
 * encryption masks and noise - never derived from a caller seed, a counter or anything persisted.          */
enum { FSC_NOISE_GAUSSIAN = 0, FSC_NOISE_TUNIFORM = 1 };
typedef struct {
    uint32_t noise_kind;            /* FSC_NOISE_GAUSSIAN: standard deviations below (fractions of q)          */
    uint32_t lwe_tuniform_bound;    /* FSC_NOISE_TUNIFORM: noise in [-2^b, 2^b]                                */
    uint32_t glwe_tuniform_bound;
    uint32_t reserved;
    double lwe_noise_std;
    double glwe_noise_std;
} fsc_noise_params;
typedef struct fsc_client fsc_client;
fsc_status fsc_client_keygen(const fsc_params *params, const fsc_noise_params *noise, fsc_client **out);
/* TEST-ONLY deterministic variant: the master key is expanded from a 64-bit seed (at most 64 bits of entropy, anyone
 * who knows the seed re-derives the secret key).  Encryption randomness is still fresh OS entropy ...           */
fsc_status fsc_client_keygen_seeded(const fsc_params *params, const fsc_noise_params *noise, uint64_t seed, fsc_client **out);
/* ... unless a test pins it too (reproducible ciphertexts, e.g. identical inputs on every rank of a multi-GPU run). */
fsc_status fsc_client_set_encryption_seed(fsc_client *c, uint64_t seed);
fsc_status fsc_client_free(fsc_client *c);
const char *fsc_client_last_error(const fsc_client *c);
/* server key material for fsc_keys_upload; the pointers stay valid until fsc_client_free            */
fsc_status fsc_client_server_keys(const fsc_client *c, const uint64_t **bsk, size_t *bsk_words,
                                  const uint64_t **ksk, size_t *ksk_words);
fsc_status fsc_client_secret_keys(const fsc_client *c, const uint64_t **lwe_sk, const uint64_t **glwe_sk);
/* values: one plaintext (message + carry, < message_modulus*carry_modulus) per block -> big LWE blocks */
fsc_status fsc_client_encrypt_blocks(fsc_client *c, const uint8_t *values, size_t n_blocks, uint64_t *out_blocks);
/* values: decoded plaintext incl. the padding bit (0..31); noise (optional): signed phase error per block */
fsc_status fsc_client_decrypt_blocks(fsc_client *c, const uint64_t *blocks, size_t n_blocks, uint8_t *values, int64_t *noise);

/* ---- on-disk formats (host CPU; container layout documented in csrc/keyfile.cpp and INTEGRATION.md) ----
 * Replaces nothing in the reference, which regenerates keys in every test (src/biguint.rs:277); a deployment keeps
 * the client key (parameters + noise + 256-bit master key + secret bits, a few KB: the 123 MB of server key material are
 * re-derived from the master key on load; no encryption state is stored) on the signer's side and ships the expanded
 * server key, without secrets, to the GPU host.  The checksum detects corruption; it is not an authenticator.   */
fsc_status fsc_client_save(const fsc_client *c, const char *path);
fsc_status fsc_client_load(const char *path, fsc_client **out);
fsc_status fsc_server_keys_save(const fsc_client *c, const char *path);
/* bsk and ksk point into ONE allocation: release it with fsc_buffer_free(*bsk)                               */
fsc_status fsc_server_keys_load(const char *path, fsc_params *params, uint64_t **bsk, size_t *bsk_words,
                                uint64_t **ksk, size_t *ksk_words);
/* big LWE blocks (radix digits), n_blocks x (k N + 1) words                                                  */
fsc_status fsc_blocks_save(const char *path, const fsc_params *params, const uint64_t *blocks, size_t n_blocks);
fsc_status fsc_blocks_load(const char *path, fsc_params *params, uint64_t **blocks, size_t *n_blocks);
fsc_status fsc_buffer_free(uint64_t *buffer);

/* ---- measurement & test hooks ---------------------------------------------------------- */
/* CUDA-event timer on the context's stream.                                                      */
fsc_status fsc_timer_start(fsc_ctx *ctx);
fsc_status fsc_timer_stop(fsc_ctx *ctx, float *elapsed_ms);     /* synchronises */
/* number of kernels this context has launched so far                                             */
fsc_status fsc_launch_count(const fsc_ctx *ctx, uint64_t *out);
/* Name of the blind-rotation kernel this context uses for wide batches ("pbs_stream_kernel",
 * "pbs_ring_kernel" or "pbs_pair_kernel"; fixed at fsc_keys_upload).  Static string. */
const char *fsc_pbs_kernel_name(const fsc_ctx *ctx);
/* Measures the device's sustained FP64 FMA rate (TFLOP/s, 2 flops per FMA) with a register-resident
 * FMA-chain kernel: the denominator of the blind rotation's compute roofline.                   */
fsc_status fsc_measure_fp64_peak(fsc_ctx *ctx, double *tflops);
/* Test hook: the Fourier-domain bootstrapping key as the context holds it (n * 32 * 4 * 32 complex doubles = 2 x that many
 * doubles; stream_order = 0: ring-kernel layout [n][slot][g][lane], 1: stream-kernel layout [n][position][g][lane]).
 * FSC_ERR_BAD_ARG if the context does not hold that layout.                                       */
fsc_status fsc_debug_fourier_key(fsc_ctx *ctx, double *out, int32_t stream_order);
/* c = a * b (negacyclic, mod 2^64) through the blind rotation's own FFT; a: count*N torus words,
 * b: count*N small signed integers (|b| < 2^23).  Host buffers.                                  */
fsc_status fsc_debug_negacyclic_mul(fsc_ctx *ctx, const uint64_t *a, const int64_t *b, uint64_t *c, size_t count);

#ifdef __cplusplus
}
#endif
#endif /* FHE_SIGN_CUDA_H */
