/*
 * tfhe_oracle.c — CPU restatement (plain C, OpenMP over ciphertexts) of the TFHE arithmetic
 * that every FheUint operator of the reference bottoms out in.  See tfhe_oracle.h for scope,
 * provenance and parity status.  TEST INFRASTRUCTURE ONLY — never linked into the product.
 *
 * Reference call sites this stands behind (the arithmetic itself lives in tfhe 0.10.0,
 * Cargo.lock:482-485, not vendored):
 *   FheUint32::try_encrypt   src/biguint.rs:26,207        -> orc_encrypt_big
 *   FheUint32::decrypt       src/biguint.rs:70            -> orc_phase_big + orc_decode
 *   every + * >> & on FheUint32/64  src/biguint.rs:110,116,138-143,223-248,
 *                                   src/perf_test.rs:28-54 -> orc_ks_pbs (apply_lookup_table)
 */
#define _GNU_SOURCE
#include "tfhe_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ===================================================================================== */
/* RNG: splitmix64 stream keyed by (seed, stream)                                        */
/* ===================================================================================== */
static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void orc_rng_init(orc_rng *r, uint64_t seed, uint64_t stream) {
    r->state = mix64(seed + 0x9E3779B97F4A7C15ull) ^ mix64(stream * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull);
}
uint64_t orc_rng_next(orc_rng *r) {
    r->state += 0x9E3779B97F4A7C15ull;
    return mix64(r->state);
}
int64_t orc_rng_gaussian_torus(orc_rng *r, double std) {
    /* Box-Muller; one sample per call (the second is discarded to keep the stream simple) */
    double u1 = (double)((orc_rng_next(r) >> 11) + 1) * (1.0 / 9007199254740992.0);
    double u2 = (double)(orc_rng_next(r) >> 11) * (1.0 / 9007199254740992.0);
    double g = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
    return (int64_t)llround(g * std * 18446744073709551616.0);
}
int64_t orc_rng_tuniform(orc_rng *r, uint32_t b) {
    /* values in [-2^b, 2^b]; the two end points have half the weight of the others */
    uint64_t x = orc_rng_next(r) >> (64 - (b + 2));
    return (int64_t)((x >> 1) + (x & 1)) - ((int64_t)1 << b);
}

/* ===================================================================================== */
/* parameters                                                                            */
/* ===================================================================================== */
int orc_params_preset(const char *name, orc_params *o) {
    memset(o, 0, sizeof *o);
    o->glwe_dim = 1; o->poly_size = 2048; o->pbs_level = 1;
    o->ks_base_log = 3; o->ks_level = 5; o->message_modulus = 4; o->carry_modulus = 4;
    if (!strcmp(name, "2_2_gaussian") || !strcmp(name, "toy")) {
        o->lwe_dim = 834; o->pbs_base_log = 23; o->noise_kind = ORC_NOISE_GAUSSIAN;
        o->lwe_noise_std = 3.5539902359442825e-06; o->glwe_noise_std = 2.845267479601915e-15;
        if (!strcmp(name, "toy")) o->lwe_dim = 48;
        return 0;
    }
    if (!strcmp(name, "2_2_tuniform")) {
        o->lwe_dim = 887; o->pbs_base_log = 22; o->noise_kind = ORC_NOISE_TUNIFORM;
        o->lwe_tuniform_bound = 46; o->glwe_tuniform_bound = 17;
        return 0;
    }
    return -1;
}

static inline int64_t noise_lwe(const orc_params *p, orc_rng *r) {
    return p->noise_kind == ORC_NOISE_TUNIFORM ? orc_rng_tuniform(r, p->lwe_tuniform_bound)
                                               : orc_rng_gaussian_torus(r, p->lwe_noise_std);
}
static inline int64_t noise_glwe(const orc_params *p, orc_rng *r) {
    return p->noise_kind == ORC_NOISE_TUNIFORM ? orc_rng_tuniform(r, p->glwe_tuniform_bound)
                                               : orc_rng_gaussian_torus(r, p->glwe_noise_std);
}

/* ===================================================================================== */
/* negacyclic FFT: N reals -> M=N/2 complex, X_k = sum_j (a_j + i a_{j+M}) zeta^{j(4k+1)},  */
/* zeta = exp(2 pi i / 2N).  Forward = twist + radix-4 DIF (natural in, scrambled out);   */
/* inverse = the exact stage-wise reverse + untwist.  Point-wise products happen in       */
/* the scrambled order, so no permutation pass is needed.                                 */
/* ===================================================================================== */
typedef struct {
    uint32_t N, M;
    double *twr, *twi;    /* twist zeta^j, j<M                                                  */
    double *w1r, *w1i, *w2r, *w2i, *w3r, *w3i;   /* radix-4 stage twiddles, stage with quarter q at offset q-1... see plan_get */
    int has_radix2;       /* one leading radix-2 stage when log2 M is odd                        */
    double *r2r, *r2i;    /* twiddles of that radix-2 stage: exp(2 pi i j / M), j < M/2           */
} fft_plan;

static fft_plan *g_plans[32];

/* Radix-4 decimation in frequency (natural in, digit-reversed out) for the forward transform and its exact
 * stage-by-stage inverse (decimation in time) for the backward one; an extra radix-2 stage in front when
 * log2 M is odd.  Point-wise products happen in the scrambled order, so no permutation pass is needed. */
static fft_plan *plan_get(uint32_t N) {
    int lg = 0; while ((1u << lg) < N) lg++;
    fft_plan *pl = g_plans[lg];
    if (pl) return pl;
#pragma omp critical(orc_plan)
    {
        if (!g_plans[lg]) {
            fft_plan *q = (fft_plan *)calloc(1, sizeof *q);
            const long double PI = 3.14159265358979323846264338327950288L;
            uint32_t M = N / 2;
            q->N = N; q->M = M;
            q->twr = (double *)malloc(sizeof(double) * M); q->twi = (double *)malloc(sizeof(double) * M);
            for (uint32_t j = 0; j < M; j++) {
                long double a = PI * (long double)j / (long double)N;
                q->twr[j] = (double)cosl(a); q->twi[j] = (double)sinl(a);
            }
            int lm = lg - 1;
            q->has_radix2 = lm & 1;
            uint32_t R = q->has_radix2 ? M / 2 : M;      /* size handled by the radix-4 stages */
            if (q->has_radix2) {
                q->r2r = (double *)malloc(sizeof(double) * (M / 2)); q->r2i = (double *)malloc(sizeof(double) * (M / 2));
                for (uint32_t j = 0; j < M / 2; j++) {
                    long double a = 2 * PI * (long double)j / (long double)M;
                    q->r2r[j] = (double)cosl(a); q->r2i[j] = (double)sinl(a);
                }
            }
            /* tables: for every quarter size qq = R/4, R/16, ..., 1 the arrays w^j, w^2j, w^3j (j < qq), w = exp(2 pi i / 4qq);
             * stage with quarter qq is stored at offset (qq - 1) / 3 * ... -> simply at offset qq - 1 in arrays of size R */
            q->w1r = (double *)calloc(R, sizeof(double)); q->w1i = (double *)calloc(R, sizeof(double));
            q->w2r = (double *)calloc(R, sizeof(double)); q->w2i = (double *)calloc(R, sizeof(double));
            q->w3r = (double *)calloc(R, sizeof(double)); q->w3i = (double *)calloc(R, sizeof(double));
            for (uint32_t qq = 1; qq < R; qq <<= 2)
                for (uint32_t j = 0; j < qq; j++) {
                    long double a = 2 * PI * (long double)j / (long double)(4 * qq);
                    q->w1r[qq - 1 + j] = (double)cosl(a);     q->w1i[qq - 1 + j] = (double)sinl(a);
                    q->w2r[qq - 1 + j] = (double)cosl(2 * a); q->w2i[qq - 1 + j] = (double)sinl(2 * a);
                    q->w3r[qq - 1 + j] = (double)cosl(3 * a); q->w3i[qq - 1 + j] = (double)sinl(3 * a);
                }
            g_plans[lg] = q;
        }
    }
    return g_plans[lg];
}

/* radix-4 DIF stages over `len` points starting at re/im */
static void r4_fwd(const fft_plan *pl, double *re, double *im, uint32_t len) {
    for (uint32_t q = len >> 2; q >= 1; q >>= 2) {
        const double *w1r = pl->w1r + q - 1, *w1i = pl->w1i + q - 1, *w2r = pl->w2r + q - 1, *w2i = pl->w2i + q - 1,
                     *w3r = pl->w3r + q - 1, *w3i = pl->w3i + q - 1;
        for (uint32_t s = 0; s < len; s += 4 * q) {
            double *ar = re + s, *ai = im + s, *br = ar + q, *bi = ai + q, *cr = br + q, *ci = bi + q, *dr = cr + q, *di = ci + q;
#pragma omp simd
            for (uint32_t j = 0; j < q; j++) {
                const double t0r = ar[j] + cr[j], t0i = ai[j] + ci[j], t1r = ar[j] - cr[j], t1i = ai[j] - ci[j];
                const double t2r = br[j] + dr[j], t2i = bi[j] + di[j];
                const double t3r = -(bi[j] - di[j]), t3i = br[j] - dr[j];          /* i * (b - d) */
                const double y1r = t1r + t3r, y1i = t1i + t3i, y2r = t0r - t2r, y2i = t0i - t2i, y3r = t1r - t3r, y3i = t1i - t3i;
                ar[j] = t0r + t2r; ai[j] = t0i + t2i;
                br[j] = y1r * w1r[j] - y1i * w1i[j]; bi[j] = y1r * w1i[j] + y1i * w1r[j];
                cr[j] = y2r * w2r[j] - y2i * w2i[j]; ci[j] = y2r * w2i[j] + y2i * w2r[j];
                dr[j] = y3r * w3r[j] - y3i * w3i[j]; di[j] = y3r * w3i[j] + y3i * w3r[j];
            }
            if (q == 1) continue;
        }
    }
}
/* exact inverse of r4_fwd up to a factor len */
static void r4_inv(const fft_plan *pl, double *re, double *im, uint32_t len) {
    for (uint32_t q = 1; q < len; q <<= 2) {
        const double *w1r = pl->w1r + q - 1, *w1i = pl->w1i + q - 1, *w2r = pl->w2r + q - 1, *w2i = pl->w2i + q - 1,
                     *w3r = pl->w3r + q - 1, *w3i = pl->w3i + q - 1;
        for (uint32_t s = 0; s < len; s += 4 * q) {
            double *ar = re + s, *ai = im + s, *br = ar + q, *bi = ai + q, *cr = br + q, *ci = bi + q, *dr = cr + q, *di = ci + q;
#pragma omp simd
            for (uint32_t j = 0; j < q; j++) {
                const double y0r = ar[j], y0i = ai[j];
                const double y1r = br[j] * w1r[j] + bi[j] * w1i[j], y1i = bi[j] * w1r[j] - br[j] * w1i[j];   /* times conj(w^j) */
                const double y2r = cr[j] * w2r[j] + ci[j] * w2i[j], y2i = ci[j] * w2r[j] - cr[j] * w2i[j];
                const double y3r = dr[j] * w3r[j] + di[j] * w3i[j], y3i = di[j] * w3r[j] - dr[j] * w3i[j];
                const double s0r = y0r + y2r, s0i = y0i + y2i, s1r = y0r - y2r, s1i = y0i - y2i;
                const double s2r = y1r + y3r, s2i = y1i + y3i;
                const double s3r = y1i - y3i, s3i = -(y1r - y3r);                   /* -i * (y1 - y3) */
                ar[j] = s0r + s2r; ai[j] = s0i + s2i;
                br[j] = s1r + s3r; bi[j] = s1i + s3i;
                cr[j] = s0r - s2r; ci[j] = s0i - s2i;
                dr[j] = s1r - s3r; di[j] = s1i - s3i;
            }
        }
    }
}

/* in: real coefficients as doubles c[0..N); out: re/im [M] (scrambled order) */
static void fft_fwd(const fft_plan *pl, const double *c, double *re, double *im) {
    const uint32_t M = pl->M;
#pragma omp simd
    for (uint32_t j = 0; j < M; j++) {
        double a = c[j], b = c[j + M];
        re[j] = a * pl->twr[j] - b * pl->twi[j];
        im[j] = a * pl->twi[j] + b * pl->twr[j];
    }
    if (pl->has_radix2) {
        const uint32_t h = M / 2;
#pragma omp simd
        for (uint32_t j = 0; j < h; j++) {
            double ur = re[j], ui = im[j], vr = re[j + h], vi = im[j + h];
            double dr = ur - vr, di = ui - vi;
            re[j] = ur + vr; im[j] = ui + vi;
            re[j + h] = dr * pl->r2r[j] - di * pl->r2i[j];
            im[j + h] = dr * pl->r2i[j] + di * pl->r2r[j];
        }
        r4_fwd(pl, re, im, h); r4_fwd(pl, re + h, im + h, h);
    } else {
        r4_fwd(pl, re, im, M);
    }
}

/* in: re/im [M] scrambled (destroyed); out: N real coefficients (scaled by 1/M) */
static void fft_inv(const fft_plan *pl, double *re, double *im, double *c) {
    const uint32_t M = pl->M;
    if (pl->has_radix2) {
        const uint32_t h = M / 2;
        r4_inv(pl, re, im, h); r4_inv(pl, re + h, im + h, h);
#pragma omp simd
        for (uint32_t j = 0; j < h; j++) {
            double vr = re[j + h] * pl->r2r[j] + im[j + h] * pl->r2i[j];
            double vi = im[j + h] * pl->r2r[j] - re[j + h] * pl->r2i[j];
            double ur = re[j], ui = im[j];
            re[j] = ur + vr; im[j] = ui + vi;
            re[j + h] = ur - vr; im[j + h] = ui - vi;
        }
    } else {
        r4_inv(pl, re, im, M);
    }
    const double sc = 1.0 / (double)M;
#pragma omp simd
    for (uint32_t j = 0; j < M; j++) {
        double a = re[j] * sc, b = im[j] * sc;
        c[j]     = a * pl->twr[j] + b * pl->twi[j];    /* times conj(twist) */
        c[j + M] = b * pl->twr[j] - a * pl->twi[j];
    }
}

/* double (any magnitude < 2^116) -> nearest integer mod 2^64 */
static inline uint64_t f64_to_torus(double x) {
    double q = x * (1.0 / 18446744073709551616.0);
    q = nearbyint(q);
    double r = fma(-q, 18446744073709551616.0, x);     /* exact: |r| <= 2^63 */
    r = nearbyint(r);
    if (r >= 9223372036854775808.0) return 0x8000000000000000ull;
    return (uint64_t)(int64_t)r;
}

/* ===================================================================================== */
/* keys                                                                                  */
/* ===================================================================================== */
struct orc_keys {
    orc_params p;
    uint64_t seed;
    uint64_t *lwe_sk, *glwe_sk, *bsk, *ksk;
    size_t bsk_len, ksk_len;
    double *bsk_re, *bsk_im;          /* Fourier BSK: [n][k+1][l][k+1][M] */
};

enum { ST_LWE_SK = 1, ST_GLWE_SK = 2, ST_BSK = 1000, ST_KSK = 2000000 };

/* body[j] += sum_m (a_m * S_m)[j]  (negacyclic, S binary) */
static void add_mask_times_key(const orc_params *p, const uint64_t *glwe_sk, const uint64_t *mask, uint64_t *body) {
    const uint32_t N = p->poly_size;
    for (uint32_t m = 0; m < p->glwe_dim; m++) {
        const uint64_t *a = mask + (size_t)m * N, *S = glwe_sk + (size_t)m * N;
        for (uint32_t t = 0; t < N; t++) {
            if (!S[t]) continue;
            /* a * X^t : coefficient j gets a[j-t], negated on wrap */
            for (uint32_t j = t; j < N; j++) body[j] += a[j - t];
            for (uint32_t j = 0; j < t; j++) body[j] -= a[j + N - t];
        }
    }
}

orc_keys *orc_keygen(const orc_params *p, uint64_t seed) {
    orc_keys *K = (orc_keys *)calloc(1, sizeof *K);
    K->p = *p; K->seed = seed;
    const uint32_t n = p->lwe_dim, k = p->glwe_dim, N = p->poly_size, L = p->pbs_level, M = N / 2;
    orc_rng r;
    K->lwe_sk = (uint64_t *)malloc(sizeof(uint64_t) * n);
    K->glwe_sk = (uint64_t *)malloc(sizeof(uint64_t) * k * N);
    orc_rng_init(&r, seed, ST_LWE_SK);
    for (uint32_t i = 0; i < n; i++) K->lwe_sk[i] = orc_rng_next(&r) >> 63;
    orc_rng_init(&r, seed, ST_GLWE_SK);
    for (uint32_t i = 0; i < k * N; i++) K->glwe_sk[i] = orc_rng_next(&r) >> 63;

    /* bootstrapping key: GGSW(s_i), rows (p, l), each a GLWE of k+1 polynomials */
    const size_t row = (size_t)(k + 1) * N, ggsw = (size_t)(k + 1) * L * row;
    K->bsk_len = (size_t)n * ggsw;
    K->bsk = (uint64_t *)malloc(sizeof(uint64_t) * K->bsk_len);
#pragma omp parallel for schedule(dynamic, 4)
    for (uint32_t i = 0; i < n; i++) {
        orc_rng rr; orc_rng_init(&rr, seed, ST_BSK + i);
        for (uint32_t pp = 0; pp <= k; pp++)
            for (uint32_t l = 0; l < L; l++) {
                uint64_t *g = K->bsk + (size_t)i * ggsw + ((size_t)pp * L + l) * row;
                for (size_t t = 0; t < (size_t)k * N; t++) g[t] = orc_rng_next(&rr);
                uint64_t *body = g + (size_t)k * N;
                for (uint32_t j = 0; j < N; j++) body[j] = (uint64_t)noise_glwe(p, &rr);
                add_mask_times_key(p, K->glwe_sk, g, body);
                const uint64_t fac = (uint64_t)1 << (64 - p->pbs_base_log * (l + 1));
                if (K->lwe_sk[i]) {
                    if (pp < k) { for (uint32_t j = 0; j < N; j++) body[j] -= K->glwe_sk[(size_t)pp * N + j] * fac; }
                    else body[0] += fac;
                }
            }
    }
    /* keyswitching key: for each big-key bit, l_ks LWE encryptions under the small key */
    const uint32_t KL = p->ks_level;
    K->ksk_len = (size_t)k * N * KL * (n + 1);
    K->ksk = (uint64_t *)malloc(sizeof(uint64_t) * K->ksk_len);
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < k * N; i++) {
        orc_rng rr; orc_rng_init(&rr, seed, ST_KSK + i);
        for (uint32_t l = 0; l < KL; l++) {
            uint64_t *c = K->ksk + ((size_t)i * KL + l) * (n + 1);
            uint64_t b = (uint64_t)noise_lwe(p, &rr);
            for (uint32_t t = 0; t < n; t++) { c[t] = orc_rng_next(&rr); b += c[t] * K->lwe_sk[t]; }
            b += K->glwe_sk[i] << (64 - p->ks_base_log * (l + 1));
            c[n] = b;
        }
    }
    /* Fourier-domain BSK */
    const size_t npoly = (size_t)n * (k + 1) * L * (k + 1);
    K->bsk_re = (double *)malloc(sizeof(double) * npoly * M);
    K->bsk_im = (double *)malloc(sizeof(double) * npoly * M);
    const fft_plan *pl = plan_get(N);
#pragma omp parallel
    {
        double *c = (double *)malloc(sizeof(double) * N);
#pragma omp for schedule(static)
        for (size_t q = 0; q < npoly; q++) {
            const uint64_t *src = K->bsk + q * N;
            for (uint32_t j = 0; j < N; j++) c[j] = (double)(int64_t)src[j];
            fft_fwd(pl, c, K->bsk_re + q * M, K->bsk_im + q * M);
        }
        free(c);
    }
    return K;
}

void orc_keys_free(orc_keys *K) {
    if (!K) return;
    free(K->lwe_sk); free(K->glwe_sk); free(K->bsk); free(K->ksk); free(K->bsk_re); free(K->bsk_im); free(K);
}
const orc_params *orc_keys_params(const orc_keys *k) { return &k->p; }
const uint64_t *orc_lwe_sk(const orc_keys *k) { return k->lwe_sk; }
const uint64_t *orc_glwe_sk(const orc_keys *k) { return k->glwe_sk; }
const uint64_t *orc_bsk(const orc_keys *k) { return k->bsk; }
const uint64_t *orc_ksk(const orc_keys *k) { return k->ksk; }
size_t orc_bsk_len(const orc_keys *k) { return k->bsk_len; }
size_t orc_ksk_len(const orc_keys *k) { return k->ksk_len; }

/* ===================================================================================== */
/* client side                                                                           */
/* ===================================================================================== */
void orc_encrypt_big(const orc_keys *K, const uint64_t *plain, size_t count, uint64_t seed, uint64_t stream, uint64_t *out) {
    const size_t d = (size_t)K->p.glwe_dim * K->p.poly_size;
    for (size_t c = 0; c < count; c++) {
        orc_rng r; orc_rng_init(&r, seed, stream + c);
        uint64_t *ct = out + c * (d + 1);
        uint64_t b = (uint64_t)noise_glwe(&K->p, &r) + plain[c];
        for (size_t i = 0; i < d; i++) { ct[i] = orc_rng_next(&r); b += ct[i] * K->glwe_sk[i]; }
        ct[d] = b;
    }
}
void orc_encrypt_small(const orc_keys *K, const uint64_t *plain, size_t count, uint64_t seed, uint64_t stream, uint64_t *out) {
    const size_t d = K->p.lwe_dim;
    for (size_t c = 0; c < count; c++) {
        orc_rng r; orc_rng_init(&r, seed, stream + c);
        uint64_t *ct = out + c * (d + 1);
        uint64_t b = (uint64_t)noise_lwe(&K->p, &r) + plain[c];
        for (size_t i = 0; i < d; i++) { ct[i] = orc_rng_next(&r); b += ct[i] * K->lwe_sk[i]; }
        ct[d] = b;
    }
}
void orc_phase_big(const orc_keys *K, const uint64_t *ct, size_t count, uint64_t *phase) {
    const size_t d = (size_t)K->p.glwe_dim * K->p.poly_size;
    for (size_t c = 0; c < count; c++) {
        const uint64_t *x = ct + c * (d + 1);
        uint64_t b = x[d];
        for (size_t i = 0; i < d; i++) b -= x[i] * K->glwe_sk[i];
        phase[c] = b;
    }
}
void orc_phase_small(const orc_keys *K, const uint64_t *ct, size_t count, uint64_t *phase) {
    const size_t d = K->p.lwe_dim;
    for (size_t c = 0; c < count; c++) {
        const uint64_t *x = ct + c * (d + 1);
        uint64_t b = x[d];
        for (size_t i = 0; i < d; i++) b -= x[i] * K->lwe_sk[i];
        phase[c] = b;
    }
}
uint64_t orc_delta(const orc_params *p) {
    return ((uint64_t)1 << 63) / ((uint64_t)p->message_modulus * p->carry_modulus);
}
uint64_t orc_decode(const orc_params *p, uint64_t phase) {
    const uint64_t delta = orc_delta(p);
    return ((phase + delta / 2) / delta) % ((uint64_t)2 * p->message_modulus * p->carry_modulus);
}

/* ===================================================================================== */
/* server side                                                                           */
/* ===================================================================================== */
/* Signed gadget decomposition, closest representable value first, then balanced digits
 * produced from the least significant level upwards (digits[l] multiplies q / B^(l+1)). */
void orc_decompose(uint64_t x, uint32_t base_log, uint32_t level, int64_t *digits) {
    const uint32_t rep = base_log * level, non_rep = 64 - rep;
    uint64_t state = ((x >> (non_rep - 1)) + 1) >> 1;
    if (rep < 64) state &= ((uint64_t)1 << rep) - 1;
    const uint64_t B = (uint64_t)1 << base_log;
    for (int l = (int)level - 1; l >= 0; l--) {
        uint64_t d = state & (B - 1);
        state >>= base_log;
        uint64_t carry = (((d - 1) | state) & d) >> (base_log - 1);
        state += carry;
        digits[l] = (int64_t)d - (int64_t)(carry << base_log);
    }
}

uint32_t orc_modswitch(uint64_t x, uint32_t N) {
    int lg = 0; while ((1u << lg) < 2 * N) lg++;
    return (uint32_t)((((x >> (64 - lg - 1)) + 1) >> 1) & (2 * N - 1));
}

void orc_keyswitch(const orc_keys *K, const uint64_t *in, size_t count, uint64_t *out) {
    const orc_params *p = &K->p;
    const size_t d = (size_t)p->glwe_dim * p->poly_size, n = p->lwe_dim;
    const uint32_t L = p->ks_level;
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < count; c++) {
        const uint64_t *x = in + c * (d + 1);
        uint64_t *y = out + c * (n + 1);
        int64_t dg[64];
        memset(y, 0, sizeof(uint64_t) * (n + 1));
        y[n] = x[d];
        for (size_t i = 0; i < d; i++) {
            orc_decompose(x[i], p->ks_base_log, L, dg);
            for (uint32_t l = 0; l < L; l++) {
                if (!dg[l]) continue;
                const uint64_t *row = K->ksk + (i * L + l) * (n + 1);
                const uint64_t m = (uint64_t)dg[l];
                for (size_t t = 0; t <= n; t++) y[t] -= m * row[t];
            }
        }
    }
}

void orc_make_lut(const orc_params *p, const uint64_t *table, uint64_t *lut) {
    const uint32_t N = p->poly_size, space = p->message_modulus * p->carry_modulus, box = N / space, half = box / 2;
    const uint64_t delta = orc_delta(p);
    uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * N);
    for (uint32_t i = 0; i < space; i++)
        for (uint32_t j = 0; j < box; j++) tmp[i * box + j] = table[i] * delta;
    for (uint32_t j = 0; j < N; j++) {
        uint32_t s = j + half;
        lut[j] = s < N ? tmp[s] : (uint64_t)0 - tmp[s - N];
    }
    free(tmp);
}

/* out = X^t * in (negacyclic), t in [0, 2N) */
static void poly_rotate(uint32_t N, const uint64_t *in, uint32_t t, uint64_t *out) {
    for (uint32_t j = 0; j < N; j++) {
        uint32_t s = (j + 2 * N - t) & (2 * N - 1);
        out[j] = s < N ? in[s] : (uint64_t)0 - in[s - N];
    }
}

typedef struct {
    double *c, *fre, *fim, *ore, *oim;
    uint64_t *rot, *acc;
} pbs_scratch;

static void scratch_init(pbs_scratch *s, const orc_params *p) {
    const uint32_t N = p->poly_size, M = N / 2, k1 = p->glwe_dim + 1;
    s->c = (double *)malloc(sizeof(double) * N);
    s->fre = (double *)malloc(sizeof(double) * M); s->fim = (double *)malloc(sizeof(double) * M);
    s->ore = (double *)malloc(sizeof(double) * M * k1); s->oim = (double *)malloc(sizeof(double) * M * k1);
    s->rot = (uint64_t *)malloc(sizeof(uint64_t) * N * k1);
    s->acc = (uint64_t *)malloc(sizeof(uint64_t) * N * k1);
}
static void scratch_free(pbs_scratch *s) {
    free(s->c); free(s->fre); free(s->fim); free(s->ore); free(s->oim); free(s->rot); free(s->acc);
}

/* acc += GGSW_i (x) diff, diff given in s->rot ([k+1][N]) */
static void external_product_add(const orc_keys *K, uint32_t i, pbs_scratch *s) {
    const orc_params *p = &K->p;
    const uint32_t N = p->poly_size, M = N / 2, k1 = p->glwe_dim + 1, L = p->pbs_level;
    const fft_plan *pl = plan_get(N);
    memset(s->ore, 0, sizeof(double) * M * k1); memset(s->oim, 0, sizeof(double) * M * k1);
    int64_t dg[64];
    for (uint32_t pp = 0; pp < k1; pp++)
        for (uint32_t l = 0; l < L; l++) {
            const uint64_t *d = s->rot + (size_t)pp * N;
            if (L == 1) {            /* single level: closest multiple of q/B, balanced digit (same rule as orc_decompose) */
                const uint32_t B = p->pbs_base_log;
                const uint64_t mask = ((uint64_t)1 << B) - 1, half = (uint64_t)1 << (B - 1);
                for (uint32_t j = 0; j < N; j++) {
                    const uint64_t st = (((d[j] >> (64 - B - 1)) + 1) >> 1) & mask;
                    s->c[j] = (double)(st > half ? (int64_t)st - ((int64_t)1 << B) : (int64_t)st);
                }
            } else
                for (uint32_t j = 0; j < N; j++) { orc_decompose(d[j], p->pbs_base_log, L, dg); s->c[j] = (double)dg[l]; }
            fft_fwd(pl, s->c, s->fre, s->fim);
            for (uint32_t q = 0; q < k1; q++) {
                const size_t off = ((((size_t)i * k1 + pp) * L + l) * k1 + q) * M;
                const double *gr = K->bsk_re + off, *gi = K->bsk_im + off;
                double *orr = s->ore + (size_t)q * M, *oii = s->oim + (size_t)q * M;
#pragma omp simd
                for (uint32_t t = 0; t < M; t++) {
                    orr[t] += s->fre[t] * gr[t] - s->fim[t] * gi[t];
                    oii[t] += s->fre[t] * gi[t] + s->fim[t] * gr[t];
                }
            }
        }
    for (uint32_t q = 0; q < k1; q++) {
        fft_inv(pl, s->ore + (size_t)q * M, s->oim + (size_t)q * M, s->c);
        uint64_t *a = s->acc + (size_t)q * N;
        for (uint32_t j = 0; j < N; j++) a[j] += f64_to_torus(s->c[j]);
    }
}

static void blind_rotate_one(const orc_keys *K, const uint64_t *ct, const uint64_t *lut, pbs_scratch *s) {
    const orc_params *p = &K->p;
    const uint32_t N = p->poly_size, k = p->glwe_dim, n = p->lwe_dim;
    memset(s->acc, 0, sizeof(uint64_t) * N * k);
    uint32_t bt = orc_modswitch(ct[n], N);
    poly_rotate(N, lut, (2 * N - bt) & (2 * N - 1), s->acc + (size_t)k * N);
    for (uint32_t i = 0; i < n; i++) {
        uint32_t at = orc_modswitch(ct[i], N);
        if (!at) continue;
        for (uint32_t q = 0; q <= k; q++) {
            uint64_t *r = s->rot + (size_t)q * N, *a = s->acc + (size_t)q * N;
            poly_rotate(N, a, at, r);
            for (uint32_t j = 0; j < N; j++) r[j] -= a[j];
        }
        external_product_add(K, i, s);
    }
}

void orc_sample_extract(const orc_params *p, const uint64_t *glwe, size_t count, uint64_t *out) {
    const uint32_t N = p->poly_size, k = p->glwe_dim;
    for (size_t c = 0; c < count; c++) {
        const uint64_t *g = glwe + c * (size_t)(k + 1) * N;
        uint64_t *o = out + c * ((size_t)k * N + 1);
        for (uint32_t m = 0; m < k; m++) {
            o[(size_t)m * N] = g[(size_t)m * N];
            for (uint32_t j = 1; j < N; j++) o[(size_t)m * N + j] = (uint64_t)0 - g[(size_t)m * N + N - j];
        }
        o[(size_t)k * N] = g[(size_t)k * N];
    }
}

void orc_blind_rotate(const orc_keys *K, const uint64_t *in_small, size_t count, const uint64_t *luts,
                      const uint32_t *lut_idx, uint64_t *out_glwe, int nthreads) {
    const orc_params *p = &K->p;
    const uint32_t N = p->poly_size, k1 = p->glwe_dim + 1, n = p->lwe_dim;
    if (nthreads <= 0) nthreads = orc_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
        pbs_scratch s; scratch_init(&s, p);
#pragma omp for schedule(dynamic, 1)
        for (size_t c = 0; c < count; c++) {
            blind_rotate_one(K, in_small + c * (n + 1), luts + (size_t)(lut_idx ? lut_idx[c] : 0) * N, &s);
            memcpy(out_glwe + c * (size_t)k1 * N, s.acc, sizeof(uint64_t) * k1 * N);
        }
        scratch_free(&s);
    }
}

void orc_pbs(const orc_keys *K, const uint64_t *in_small, size_t count, const uint64_t *luts,
             const uint32_t *lut_idx, uint64_t *out_big, int nthreads) {
    const orc_params *p = &K->p;
    const size_t g = (size_t)(p->glwe_dim + 1) * p->poly_size;
    uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * g * count);
    orc_blind_rotate(K, in_small, count, luts, lut_idx, tmp, nthreads);
    orc_sample_extract(p, tmp, count, out_big);
    free(tmp);
}

void orc_ks_pbs(const orc_keys *K, const uint64_t *in_big, size_t count, const uint64_t *luts,
                const uint32_t *lut_idx, uint64_t *out_big, int nthreads) {
    const orc_params *p = &K->p;
    uint64_t *small = (uint64_t *)malloc(sizeof(uint64_t) * (p->lwe_dim + 1) * count);
    orc_keyswitch(K, in_big, count, small);
    orc_pbs(K, small, count, luts, lut_idx, out_big, nthreads);
    free(small);
}

void orc_negacyclic_mul_fft(uint32_t N, const uint64_t *a, const int64_t *b, uint64_t *c) {
    const fft_plan *pl = plan_get(N);
    const uint32_t M = N / 2;
    double *t = (double *)malloc(sizeof(double) * N);
    double *ar = (double *)malloc(sizeof(double) * M), *ai = (double *)malloc(sizeof(double) * M);
    double *br = (double *)malloc(sizeof(double) * M), *bi = (double *)malloc(sizeof(double) * M);
    for (uint32_t j = 0; j < N; j++) t[j] = (double)(int64_t)a[j];
    fft_fwd(pl, t, ar, ai);
    for (uint32_t j = 0; j < N; j++) t[j] = (double)b[j];
    fft_fwd(pl, t, br, bi);
    for (uint32_t j = 0; j < M; j++) {
        double x = ar[j] * br[j] - ai[j] * bi[j], y = ar[j] * bi[j] + ai[j] * br[j];
        ar[j] = x; ai[j] = y;
    }
    fft_inv(pl, ar, ai, t);
    for (uint32_t j = 0; j < N; j++) c[j] = f64_to_torus(t[j]);
    free(t); free(ar); free(ai); free(br); free(bi);
}

void orc_negacyclic_mul_exact(uint32_t N, const uint64_t *a, const int64_t *b, uint64_t *c) {
    memset(c, 0, sizeof(uint64_t) * N);
    for (uint32_t i = 0; i < N; i++) {
        const uint64_t bi = (uint64_t)b[i];
        if (!bi) continue;
        for (uint32_t j = 0; j < N - i; j++) c[i + j] += a[j] * bi;
        for (uint32_t j = N - i; j < N; j++) c[i + j - N] -= a[j] * bi;
    }
}

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);      /* launchers such as torchrun export OMP_NUM_THREADS=1 */
#else
    (void)n;
#endif
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
