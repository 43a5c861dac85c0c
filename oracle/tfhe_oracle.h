/*
 * tfhe_oracle.h — CPU restatement of the TFHE arithmetic behind fhe-sign's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or the CPU baseline.  The product (fhe_sign_b200/) never links or
 * imports this file and fails loudly when its CUDA library is missing.
 *
 * What it restates.  The reference (/root/reference, crate `key-protocol`) performs all
 * ciphertext arithmetic inside the third-party crate `tfhe` 0.10.0 (Cargo.toml:9,
 * Cargo.lock:482-485), whose sources are NOT vendored under /root/reference and cannot be
 * built here (no cargo/rustc, no network).  Every `FheUint32/FheUint64` operator used by
 * src/biguint.rs:110,116,135-143,221-248 and src/perf_test.rs:28-54 bottoms out in
 *     shortint apply_lookup_table = LWE keyswitch (big -> small key) ; programmable bootstrap
 * with the PARAM_MESSAGE_2_CARRY_2_KS_PBS parameter set (ConfigBuilder::default(),
 * src/biguint.rs:276, src/perf_test.rs:9).  This file restates that published algorithm
 * (CGGI/TFHE: signed gadget decomposition, LWE keyswitch, blind rotation by GGSW x GLWE
 * external products with an f64 negacyclic FFT, sample extraction) from the TFHE papers and
 * the tfhe-rs documentation.
 *
 * PARITY STATUS: the reference pins this boundary only at the plaintext level
 * (decrypt(op(encrypt x)) == f(x)).  Ciphertext bits, noise and the
 * CSPRNG stream of tfhe 0.10.0 are "parity unpinned": the oracle uses its own seeded
 * generator (see orc_rng_*).  The plaintext-level known answers of the reference's tests
 *  are pinned by tests/test_oracle_*.py and tests/golden/.
 */
#ifndef TFHE_ORACLE_H
#define TFHE_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_NOISE_GAUSSIAN = 0, ORC_NOISE_TUNIFORM = 1 };

typedef struct {
    uint32_t lwe_dim;        /* n  : small LWE dimension                         */
    uint32_t glwe_dim;       /* k  : GLWE dimension                              */
    uint32_t poly_size;      /* N  : polynomial size (power of two)              */
    uint32_t pbs_base_log;   /* log2(beta_pbs)                                   */
    uint32_t pbs_level;      /* l_pbs                                            */
    uint32_t ks_base_log;    /* log2(beta_ks)                                    */
    uint32_t ks_level;       /* l_ks                                             */
    uint32_t message_modulus;
    uint32_t carry_modulus;
    uint32_t noise_kind;     /* ORC_NOISE_GAUSSIAN | ORC_NOISE_TUNIFORM          */
    uint32_t lwe_tuniform_bound;   /* b: noise uniform-ish in [-2^b, 2^b]        */
    uint32_t glwe_tuniform_bound;
    double   lwe_noise_std;  /* torus units (fraction of q), Gaussian kind       */
    double   glwe_noise_std;
} orc_params;

/* Named presets.  "2_2_gaussian": n=834 (PARAM_MESSAGE_2_CARRY_2_KS_PBS_GAUSSIAN_2M64);
 * "2_2_tuniform": n=887 (…_TUNIFORM_2M64); "toy": same GLWE side, n=48 (fast tests).     */
int orc_params_preset(const char *name, orc_params *out);

typedef struct orc_keys orc_keys;

/* ---- seeded generator ---------------------------------------------------------------- */
typedef struct { uint64_t state; } orc_rng;
void     orc_rng_init(orc_rng *r, uint64_t seed, uint64_t stream);
uint64_t orc_rng_next(orc_rng *r);
int64_t  orc_rng_gaussian_torus(orc_rng *r, double std);
int64_t  orc_rng_tuniform(orc_rng *r, uint32_t bound_log2);

/* ---- keys --------------------------------------------------------------------------- */
orc_keys *orc_keygen(const orc_params *p, uint64_t seed);
void      orc_keys_free(orc_keys *k);
const orc_params *orc_keys_params(const orc_keys *k);
const uint64_t *orc_lwe_sk(const orc_keys *k);    /* n entries in {0,1}                  */
const uint64_t *orc_glwe_sk(const orc_keys *k);   /* k*N entries in {0,1}                */
const uint64_t *orc_bsk(const orc_keys *k);       /* [n][k+1][l][k+1][N] standard domain */
const uint64_t *orc_ksk(const orc_keys *k);       /* [k*N][l_ks][n+1]                    */
size_t orc_bsk_len(const orc_keys *k);
size_t orc_ksk_len(const orc_keys *k);

/* ---- client side -------------------------------------------------------------------- */
/* Encrypt `count` plaintexts (already encoded torus values) under the big (GLWE) key.   */
void orc_encrypt_big(const orc_keys *k, const uint64_t *plain, size_t count,
                     uint64_t seed, uint64_t stream, uint64_t *out /* count*(kN+1) */);
void orc_encrypt_small(const orc_keys *k, const uint64_t *plain, size_t count,
                       uint64_t seed, uint64_t stream, uint64_t *out /* count*(n+1) */);
void orc_phase_big(const orc_keys *k, const uint64_t *ct, size_t count, uint64_t *phase);
void orc_phase_small(const orc_keys *k, const uint64_t *ct, size_t count, uint64_t *phase);
/* round(phase / delta) mod (message_modulus*carry_modulus*2) — includes the padding bit. */
uint64_t orc_decode(const orc_params *p, uint64_t phase);
uint64_t orc_delta(const orc_params *p);

/* ---- server side -------------------------------------------------------------------- */
void orc_decompose(uint64_t x, uint32_t base_log, uint32_t level, int64_t *digits);
uint32_t orc_modswitch(uint64_t x, uint32_t poly_size);
void orc_keyswitch(const orc_keys *k, const uint64_t *in, size_t count, uint64_t *out);
/* Accumulator polynomial of a function table f[0..msg*carry) (values are un-encoded).   */
void orc_make_lut(const orc_params *p, const uint64_t *table, uint64_t *lut /* N */);
/* PBS of `count` small-key ciphertexts; ciphertext c uses LUT polynomial lut_idx[c].    */
void orc_pbs(const orc_keys *k, const uint64_t *in_small, size_t count,
             const uint64_t *luts, const uint32_t *lut_idx, uint64_t *out_big, int nthreads);
void orc_ks_pbs(const orc_keys *k, const uint64_t *in_big, size_t count,
                const uint64_t *luts, const uint32_t *lut_idx, uint64_t *out_big, int nthreads);
/* Blind rotation only (no sample extract): out is count GLWE ciphertexts [(k+1)][N].    */
void orc_blind_rotate(const orc_keys *k, const uint64_t *in_small, size_t count,
                      const uint64_t *luts, const uint32_t *lut_idx, uint64_t *out_glwe,
                      int nthreads);
void orc_sample_extract(const orc_params *p, const uint64_t *glwe, size_t count, uint64_t *out_big);
/* Negacyclic product through the oracle's FFT (unit-test hook): c = a * b, b small int.  */
void orc_negacyclic_mul_fft(uint32_t N, const uint64_t *a, const int64_t *b, uint64_t *c);
/* Exact schoolbook negacyclic product mod 2^64 (ground truth for the FFT hooks).         */
void orc_negacyclic_mul_exact(uint32_t N, const uint64_t *a, const int64_t *b, uint64_t *c);
int  orc_max_threads(void);
/* Overrides OMP_NUM_THREADS for every later call (bench.py under torchrun, which exports OMP_NUM_THREADS=1). */
void orc_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
