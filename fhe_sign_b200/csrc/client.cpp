// client.cpp — the client side of the path: key generation, block encryption and decryption on the host CPU.
//
// The reference does this with tfhe's ClientKey (generate_keys `src/biguint.rs:277`, FheUint32::try_encrypt
// `src/biguint.rs:26,207`, decrypt `src/biguint.rs:70`); SURVEY.md 8b keeps it on the host.  This is plain
// integer code (LWE / GLWE encryption under binary secret keys, q = 2^64) plus noise sampling; the server-side
// key material it produces is exactly what fsc_keys_upload consumes.  Randomness: ChaCha20 streams under a 256-bit
// master key drawn from the OS (getrandom; what tfhe's seeder_unix feature does, Cargo.toml:9) for the secret keys and
// the server key material, and under a SECOND 256-bit key drawn from the OS per client instance for encryption masks
// and noise, so that no persisted or caller-chosen state ever determines encryption randomness.
// fsc_client_keygen_seeded (64-bit seed -> master key) exists for reproducible tests only.
// Nothing here runs on the GPU and nothing on the GPU path depends on it.
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/random.h>

#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fhe_sign_cuda.h"
#include "client_internal.h"

namespace {

// ---- ChaCha20 keystream -------------------------------------------------------------------
struct ChaCha {
    uint32_t st[16];
    uint32_t buf[16];
    int pos = 16;
    static uint32_t rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
    ChaCha(const uint8_t key[32], uint64_t stream) {      // 256-bit key, 64-bit block counter, 64-bit stream id as nonce
        static const uint32_t sigma[4] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
        memcpy(st, sigma, 16);
        memcpy(st + 4, key, 32);
        st[12] = 0; st[13] = 0;
        st[14] = (uint32_t)stream; st[15] = (uint32_t)(stream >> 32);
    }
    void block() {
        uint32_t x[16];
        memcpy(x, st, 64);
#define QR(a, b, c, d) x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12); \
                       x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7)
        for (int r = 0; r < 10; ++r) {
            QR(0, 4, 8, 12); QR(1, 5, 9, 13); QR(2, 6, 10, 14); QR(3, 7, 11, 15);
            QR(0, 5, 10, 15); QR(1, 6, 11, 12); QR(2, 7, 8, 13); QR(3, 4, 9, 14);
        }
#undef QR
        for (int i = 0; i < 16; ++i) buf[i] = x[i] + st[i];
        if (++st[12] == 0) ++st[13];
        pos = 0;
    }
    uint64_t next() {
        if (pos >= 16) block();
        const uint64_t v = (uint64_t)buf[pos] | ((uint64_t)buf[pos + 1] << 32);
        pos += 2;
        return v;
    }
};

struct Noise {
    uint32_t kind; double stddev; uint32_t bound_log2;
    int64_t sample(ChaCha& r) const {
        if (kind == FSC_NOISE_TUNIFORM) {
            const uint64_t x = r.next() >> (64 - (bound_log2 + 2));
            return (int64_t)((x >> 1) + (x & 1)) - ((int64_t)1 << bound_log2);
        }
        const double u1 = (double)((r.next() >> 11) + 1) * (1.0 / 9007199254740992.0);
        const double u2 = (double)(r.next() >> 11) * (1.0 / 9007199254740992.0);
        const double g = sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
        return (int64_t)llround(g * stddev * 18446744073709551616.0);
    }
};

// 256 bits from the OS CSPRNG (getrandom(2), /dev/urandom as a fallback); false if neither works
bool os_entropy(uint8_t out[32]) {
    size_t got = 0;
    while (got < 32) {
        const ssize_t r = getrandom(out + got, 32 - got, 0);
        if (r > 0) { got += (size_t)r; continue; }
        if (r < 0 && errno == EINTR) continue;
        break;
    }
    if (got == 32) return true;
    FILE* f = fopen("/dev/urandom", "rb");
    if (!f) return false;
    const bool ok = fread(out, 1, 32, f) == 32;
    fclose(f);
    return ok;
}

// test-only: a 64-bit seed expanded into a master key (splitmix64) - at most 64 bits of entropy by construction
void key_from_seed(uint64_t seed, uint8_t key[32]) {
    uint64_t z = seed;
    for (int i = 0; i < 4; ++i) {
        z += 0x9E3779B97F4A7C15ull;
        uint64_t x = z;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        x ^= x >> 31;
        memcpy(key + 8 * i, &x, 8);
    }
}

}  // namespace

struct fsc_client {
    fsc_params p;
    fsc_noise_params np;
    uint8_t master[32];            // keys the secret-key and server-key streams (persisted in the client key file)
    uint8_t enc_key[32];           // keys the encryption streams: fresh from the OS per instance, never persisted
    std::vector<uint64_t> lwe_sk, glwe_sk, bsk, ksk;
    uint64_t enc_counter = 0;      // stream id under enc_key (unique per block within this instance)
    std::string err;
    Noise lwe_noise() const { return {np.noise_kind, np.lwe_noise_std, np.lwe_tuniform_bound}; }
    Noise glwe_noise() const { return {np.noise_kind, np.glwe_noise_std, np.glwe_tuniform_bound}; }
};

static thread_local std::string g_client_error;

static void parallel_for(size_t n, const std::function<void(size_t)>& fn) {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 4;
    if (nt > 32) nt = 32;
    if (n < nt) nt = (unsigned)n;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([=, &fn]() { for (size_t i = t; i < n; i += nt) fn(i); });
    for (auto& x : th) x.join();
}

// body[j] += (a * S)[j], negacyclic, S binary
static void add_mask_times_key(size_t N, const uint64_t* a, const uint64_t* S, uint64_t* body) {
    for (size_t t = 0; t < N; ++t) {
        if (!S[t]) continue;
        for (size_t j = t; j < N; ++j) body[j] += a[j - t];
        for (size_t j = 0; j < t; ++j) body[j] -= a[j + N - t];
    }
}

fsc_params fsc_client_params(const fsc_client* c) { return c->p; }
fsc_noise_params fsc_client_noise(const fsc_client* c) { return c->np; }
const uint8_t* fsc_client_master_key(const fsc_client* c) { return c->master; }
void fsc_client_set_error(const std::string& msg) { g_client_error = msg; }

bool fsc_params_plausible(const fsc_params& p, const fsc_noise_params& np, std::string* why) {
    auto bad = [&](const char* m) { if (why) *why = m; return false; };
    if (p.glwe_dim == 0 || p.glwe_dim > 8 || p.poly_size < 2 || p.poly_size > (1u << 16) || (p.poly_size & (p.poly_size - 1)))
        return bad("glwe_dim / poly_size out of range");
    if (p.lwe_dim == 0 || p.lwe_dim > (1u << 14)) return bad("lwe_dim out of range");
    if (p.pbs_level == 0 || p.pbs_base_log == 0 || (uint64_t)p.pbs_base_log * p.pbs_level >= 64) return bad("pbs decomposition out of range");
    if (p.ks_level == 0 || p.ks_base_log == 0 || (uint64_t)p.ks_base_log * p.ks_level >= 64) return bad("keyswitch decomposition out of range");
    const uint64_t space = (uint64_t)p.message_modulus * p.carry_modulus;
    if (space == 0 || space > 256 || (space & (space - 1))) return bad("message * carry modulus must be a power of two up to 256");
    if (np.noise_kind != FSC_NOISE_GAUSSIAN && np.noise_kind != FSC_NOISE_TUNIFORM) return bad("unknown noise kind");
    if (np.lwe_tuniform_bound > 61 || np.glwe_tuniform_bound > 61) return bad("tuniform bound above 61 bits");
    if (np.noise_kind == FSC_NOISE_GAUSSIAN && (!(np.lwe_noise_std >= 0.0 && np.lwe_noise_std < 0.25) || !(np.glwe_noise_std >= 0.0 && np.glwe_noise_std < 0.25)))
        return bad("gaussian standard deviation out of range");
    return true;
}

static fsc_status keygen_from_master(const fsc_params* params, const fsc_noise_params* noise, const uint8_t master[32], fsc_client** out) {
    if (!params || !noise || !out || !master) { g_client_error = "null argument"; return FSC_ERR_BAD_ARG; }
    *out = nullptr;
    const fsc_params& p = *params;
    std::string why;
    if (!fsc_params_plausible(p, *noise, &why)) { g_client_error = "invalid parameter set: " + why; return FSC_ERR_PARAMS; }
    try {
        fsc_client* c = new fsc_client();
        c->p = p; c->np = *noise;
        memcpy(c->master, master, 32);
        if (!os_entropy(c->enc_key)) { delete c; g_client_error = "no OS entropy source (getrandom, /dev/urandom)"; return FSC_ERR_INTERNAL; }
        const uint8_t* mk = c->master;
        const size_t n = p.lwe_dim, k = p.glwe_dim, N = p.poly_size, L = p.pbs_level, KL = p.ks_level;
        c->lwe_sk.resize(n); c->glwe_sk.resize(k * N);
        { ChaCha r(mk, 1); for (auto& v : c->lwe_sk) v = r.next() >> 63; }
        { ChaCha r(mk, 2); for (auto& v : c->glwe_sk) v = r.next() >> 63; }
        // bootstrapping key: GGSW(s_i), rows (row polynomial p, level l), each a GLWE of k+1 polynomials
        const size_t row = (k + 1) * N, ggsw = (k + 1) * L * row;
        c->bsk.assign(n * ggsw, 0);
        const Noise gn = c->glwe_noise(), ln = c->lwe_noise();
        parallel_for(n, [&](size_t i) {
            ChaCha r(mk, 1000 + i);
            for (size_t pp = 0; pp <= k; ++pp)
                for (size_t l = 0; l < L; ++l) {
                    uint64_t* g = c->bsk.data() + i * ggsw + (pp * L + l) * row;
                    for (size_t t = 0; t < k * N; ++t) g[t] = r.next();
                    uint64_t* body = g + k * N;
                    for (size_t j = 0; j < N; ++j) body[j] = (uint64_t)gn.sample(r);
                    for (size_t m = 0; m < k; ++m) add_mask_times_key(N, g + m * N, c->glwe_sk.data() + m * N, body);
                    const uint64_t fac = (uint64_t)1 << (64 - p.pbs_base_log * (l + 1));
                    if (c->lwe_sk[i]) {
                        if (pp < k) { for (size_t j = 0; j < N; ++j) body[j] -= c->glwe_sk[pp * N + j] * fac; }
                        else body[0] += fac;
                    }
                }
        });
        // keyswitching key: l_ks LWE encryptions under the small key per big-key bit
        c->ksk.assign(k * N * KL * (n + 1), 0);
        parallel_for(k * N, [&](size_t i) {
            ChaCha r(mk, 2000000 + i);
            for (size_t l = 0; l < KL; ++l) {
                uint64_t* ct = c->ksk.data() + (i * KL + l) * (n + 1);
                uint64_t b = (uint64_t)ln.sample(r);
                for (size_t t = 0; t < n; ++t) { ct[t] = r.next(); b += ct[t] * c->lwe_sk[t]; }
                ct[n] = b + (c->glwe_sk[i] << (64 - p.ks_base_log * (l + 1)));
            }
        });
        *out = c;
        return FSC_OK;
    } catch (const std::bad_alloc&) {
        g_client_error = "host allocation failed"; return FSC_ERR_OOM;
    } catch (...) {
        g_client_error = "key generation failed"; return FSC_ERR_INTERNAL;
    }
}

fsc_status fsc_client_keygen_from_master(const fsc_params* params, const fsc_noise_params* noise, const uint8_t* master32, fsc_client** out) {
    return keygen_from_master(params, noise, master32, out);      // keyfile.cpp: rebuilds a client from its persisted master key
}

extern "C" {

fsc_status fsc_client_keygen(const fsc_params* params, const fsc_noise_params* noise, fsc_client** out) {
    uint8_t master[32];
    if (!os_entropy(master)) { g_client_error = "no OS entropy source (getrandom, /dev/urandom)"; if (out) *out = nullptr; return FSC_ERR_INTERNAL; }
    const fsc_status st = keygen_from_master(params, noise, master, out);
    memset(master, 0, sizeof(master));
    return st;
}

fsc_status fsc_client_keygen_seeded(const fsc_params* params, const fsc_noise_params* noise, uint64_t seed, fsc_client** out) {
    uint8_t master[32];
    key_from_seed(seed, master);
    return keygen_from_master(params, noise, master, out);
}

fsc_status fsc_client_set_encryption_seed(fsc_client* c, uint64_t seed) {
    if (!c) return FSC_ERR_BAD_ARG;
    key_from_seed(seed ^ 0xE7C0DE5EEDull, c->enc_key);
    c->enc_counter = 0;
    return FSC_OK;
}

fsc_status fsc_client_free(fsc_client* c) { delete c; return FSC_OK; }
const char* fsc_client_last_error(const fsc_client* c) { return c ? c->err.c_str() : g_client_error.c_str(); }

fsc_status fsc_client_server_keys(const fsc_client* c, const uint64_t** bsk, size_t* bsk_words, const uint64_t** ksk, size_t* ksk_words) {
    if (!c || !bsk || !bsk_words || !ksk || !ksk_words) return FSC_ERR_BAD_ARG;
    *bsk = c->bsk.data(); *bsk_words = c->bsk.size(); *ksk = c->ksk.data(); *ksk_words = c->ksk.size();
    return FSC_OK;
}

fsc_status fsc_client_secret_keys(const fsc_client* c, const uint64_t** lwe_sk, const uint64_t** glwe_sk) {
    if (!c) return FSC_ERR_BAD_ARG;
    if (lwe_sk) *lwe_sk = c->lwe_sk.data();
    if (glwe_sk) *glwe_sk = c->glwe_sk.data();
    return FSC_OK;
}

fsc_status fsc_client_encrypt_blocks(fsc_client* c, const uint8_t* values, size_t n_blocks, uint64_t* out_blocks) {
    if (!c) return FSC_ERR_BAD_ARG;
    if ((!values || !out_blocks) && n_blocks) { c->err = "null argument"; return FSC_ERR_BAD_ARG; }
    const size_t d = (size_t)c->p.glwe_dim * c->p.poly_size;
    const uint64_t space = (uint64_t)c->p.message_modulus * c->p.carry_modulus, delta = ((uint64_t)1 << 63) / space;
    for (size_t i = 0; i < n_blocks; ++i)
        if (values[i] >= space) { c->err = "block value exceeds the plaintext space"; return FSC_ERR_BAD_ARG; }
    const Noise gn = c->glwe_noise();
    const uint64_t base = c->enc_counter;
    c->enc_counter += n_blocks;
    parallel_for(n_blocks, [&](size_t i) {
        ChaCha r(c->enc_key, base + i);
        uint64_t* ct = out_blocks + i * (d + 1);
        uint64_t b = (uint64_t)gn.sample(r) + (uint64_t)values[i] * delta;
        for (size_t t = 0; t < d; ++t) { ct[t] = r.next(); b += ct[t] * c->glwe_sk[t]; }
        ct[d] = b;
    });
    return FSC_OK;
}

fsc_status fsc_client_decrypt_blocks(fsc_client* c, const uint64_t* blocks, size_t n_blocks, uint8_t* values, int64_t* noise) {
    if (!c) return FSC_ERR_BAD_ARG;
    if ((!blocks || !values) && n_blocks) { c->err = "null argument"; return FSC_ERR_BAD_ARG; }
    const size_t d = (size_t)c->p.glwe_dim * c->p.poly_size;
    const uint64_t space = (uint64_t)c->p.message_modulus * c->p.carry_modulus, delta = ((uint64_t)1 << 63) / space;
    parallel_for(n_blocks, [&](size_t i) {
        const uint64_t* ct = blocks + i * (d + 1);
        uint64_t ph = ct[d];
        for (size_t t = 0; t < d; ++t) ph -= ct[t] * c->glwe_sk[t];
        const uint64_t m = ((ph + delta / 2) / delta) % (2 * space);          // includes the padding bit
        values[i] = (uint8_t)m;
        if (noise) noise[i] = (int64_t)(ph - m * delta);
    });
    return FSC_OK;
}

}  // extern "C"
