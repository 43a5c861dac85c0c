// keyfile.cpp — on-disk formats of the client side (host CPU): seeded client key, expanded server key, ciphertext
// blocks.  SURVEY.md section 8(f) row 4: lets a Rust host persist what the reference regenerates in every test
// (tfhe::generate_keys, src/biguint.rs:277) and hand a server its key material without the secrets.
//
// One container for all three ("FSCFILE1"), little-endian, 8-byte aligned:
//
//   0   char[8]  magic "FSCFILE1"
//   8   u32      version (1)
//   12  u32      kind: 1 client key (seeded), 2 server key (expanded), 3 ciphertext blocks
//   16  u32[10]  fsc_params   (lwe_dim, glwe_dim, poly_size, pbs_base_log, pbs_level, ks_base_log, ks_level,
//                              message_modulus, carry_modulus, acc_bits)
//   56  fsc_noise_params (32 bytes; zero for kinds 2 and 3)
//   88  u64      reserved (0)
//   96  u64      aux  (kind 2: bsk word count; kind 3: number of blocks; kind 1: 0)
//   104 u64      payload words (u64 count)
//   112 payload  kind 1: 256-bit master key (4 words) | lwe_sk | glwe_sk as bit-per-word u64 (the keys are re-derived
//                        from the master key on load and must equal these; bsk / ksk are NOT stored: 123 MB regenerate
//                        in well under a second).  NO encryption state is stored: encryption masks and noise come from
//                        a fresh OS-entropy stream per client instance, so save -> encrypt -> load, or one file loaded
//                        by two processes, can never replay a mask.  The file holds the secret key: protect it as such.
//                kind 2: bsk [n][k+1][l][k+1][N] | ksk [kN][l_ks][n+1]     (the layout fsc_keys_upload takes)
//                kind 3: n_blocks x (k N + 1) words, mask first, body last
//   ...  u64     FNV-1a 64 checksum of every byte before it (detects truncation / corruption; it is NOT a MAC -
//                authenticate key files by other means if they cross a trust boundary).  Header fields that feed shift
//                amounts or allocation sizes are range-checked before use.
//
// No C++ exception crosses the boundary; errors come back as fsc_status with fsc_client_last_error(NULL).
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <new>
#include <string>
#include <vector>

#include "../../include/fhe_sign_cuda.h"
#include "client_internal.h"

namespace {

constexpr char kMagic[8] = {'F', 'S', 'C', 'F', 'I', 'L', 'E', '1'};
constexpr uint32_t kVersion = 2;      // 2: master key in the payload, no encryption counter

struct Header {
    char magic[8];
    uint32_t version, kind;
    uint32_t params[10];
    fsc_noise_params noise;
    uint64_t seed, aux, payload_words;
};
static_assert(sizeof(Header) == 112, "file header layout");

uint64_t fnv1a(const void* data, size_t bytes, uint64_t h) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
constexpr uint64_t kFnvBasis = 14695981039346656037ull;

fsc_status fail(fsc_status code, const std::string& msg) { fsc_client_set_error(msg); return code; }

fsc_status write_file(const char* path, Header h, const std::vector<std::pair<const uint64_t*, size_t>>& parts) {
    h.payload_words = 0;
    for (auto& p : parts) h.payload_words += p.second;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(FSC_ERR_BAD_ARG, std::string("cannot open for writing: ") + path);
    uint64_t sum = fnv1a(&h, sizeof(h), kFnvBasis);
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    for (auto& p : parts) {
        if (!ok) break;
        if (p.second) ok = fwrite(p.first, 8, p.second, f) == p.second;
        sum = fnv1a(p.first, p.second * 8, sum);
    }
    ok = ok && fwrite(&sum, 8, 1, f) == 1;
    ok = (fclose(f) == 0) && ok;
    return ok ? FSC_OK : fail(FSC_ERR_INTERNAL, std::string("short write: ") + path);
}

// reads and verifies a whole file of the expected kind
fsc_status read_file(const char* path, uint32_t kind, Header& h, std::vector<uint64_t>& payload) {
    FILE* f = fopen(path, "rb");
    if (!f) return fail(FSC_ERR_BAD_ARG, std::string("cannot open: ") + path);
    fsc_status st = FSC_OK;
    uint64_t sum = 0;
    if (fread(&h, sizeof(h), 1, f) != 1) st = fail(FSC_ERR_BAD_ARG, "file shorter than its header");
    else if (memcmp(h.magic, kMagic, 8) != 0) st = fail(FSC_ERR_BAD_ARG, "not an FSCFILE1 container");
    else if (h.version != kVersion) st = fail(FSC_ERR_PARAMS, "unsupported file version " + std::to_string(h.version));
    else if (h.kind != kind) st = fail(FSC_ERR_BAD_ARG, "file holds kind " + std::to_string(h.kind) + ", expected " + std::to_string(kind));
    else if (h.payload_words > ((uint64_t)1 << 34)) st = fail(FSC_ERR_BAD_ARG, "implausible payload size");
    if (st == FSC_OK) {
        try { payload.resize(h.payload_words); } catch (const std::bad_alloc&) { st = fail(FSC_ERR_OOM, "host allocation failed"); }
    }
    if (st == FSC_OK && h.payload_words && fread(payload.data(), 8, h.payload_words, f) != h.payload_words)
        st = fail(FSC_ERR_BAD_ARG, "truncated payload");
    if (st == FSC_OK && fread(&sum, 8, 1, f) != 1) st = fail(FSC_ERR_BAD_ARG, "missing checksum");
    fclose(f);
    if (st == FSC_OK && sum != fnv1a(payload.data(), payload.size() * 8, fnv1a(&h, sizeof(h), kFnvBasis)))
        st = fail(FSC_ERR_BAD_ARG, "checksum mismatch (corrupt file)");
    return st;
}

Header make_header(uint32_t kind, const fsc_params& p) {
    Header h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, kMagic, 8);
    h.version = kVersion; h.kind = kind;
    memcpy(h.params, &p, sizeof(h.params));
    return h;
}
static_assert(sizeof(fsc_params) == 40, "fsc_params is ten u32");

bool words_match(const fsc_params& p, uint64_t bsk_words, uint64_t ksk_words) {
    const uint64_t n = p.lwe_dim, k = p.glwe_dim, N = p.poly_size;
    return bsk_words == n * (k + 1) * p.pbs_level * (k + 1) * N && ksk_words == k * N * p.ks_level * (n + 1);
}

}  // namespace

extern "C" {

fsc_status fsc_client_save(const fsc_client* c, const char* path) {
    if (!c || !path) return fail(FSC_ERR_BAD_ARG, "null argument");
    Header h = make_header(1, fsc_client_params(c));
    h.noise = fsc_client_noise(c);
    const uint64_t *lwe = nullptr, *glwe = nullptr;
    fsc_client_secret_keys(c, &lwe, &glwe);
    const fsc_params p = fsc_client_params(c);
    uint64_t master[4];
    memcpy(master, fsc_client_master_key(c), 32);
    const fsc_status st = write_file(path, h, {{master, 4}, {lwe, p.lwe_dim}, {glwe, (size_t)p.glwe_dim * p.poly_size}});
    memset(master, 0, sizeof(master));
    return st;
}

fsc_status fsc_client_load(const char* path, fsc_client** out) {
    if (!path || !out) return fail(FSC_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    Header h;
    std::vector<uint64_t> payload;
    fsc_status st = read_file(path, 1, h, payload);
    if (st != FSC_OK) return st;
    fsc_params p;
    memcpy(&p, h.params, sizeof(p));
    std::string why;
    if (!fsc_params_plausible(p, h.noise, &why)) return fail(FSC_ERR_PARAMS, "key file header: " + why);      // before any shift or allocation derived from it
    if (payload.size() != 4 + (uint64_t)p.lwe_dim + (uint64_t)p.glwe_dim * p.poly_size) return fail(FSC_ERR_BAD_ARG, "secret key size does not match the parameters");
    fsc_client* c = nullptr;
    st = fsc_client_keygen_from_master(&p, &h.noise, reinterpret_cast<const uint8_t*>(payload.data()), &c);
    if (st != FSC_OK) return st;
    const uint64_t *lwe = nullptr, *glwe = nullptr;
    fsc_client_secret_keys(c, &lwe, &glwe);
    if (memcmp(lwe, payload.data() + 4, (size_t)p.lwe_dim * 8) != 0 ||
        memcmp(glwe, payload.data() + 4 + p.lwe_dim, (size_t)p.glwe_dim * p.poly_size * 8) != 0) {
        fsc_client_free(c);
        return fail(FSC_ERR_INTERNAL, "the master key does not regenerate the stored secret key (different library version?)");
    }
    *out = c;      // encryption randomness of the new instance is fresh OS entropy (nothing of it is persisted)
    return FSC_OK;
}

fsc_status fsc_server_keys_save(const fsc_client* c, const char* path) {
    if (!c || !path) return fail(FSC_ERR_BAD_ARG, "null argument");
    const uint64_t *bsk = nullptr, *ksk = nullptr;
    size_t bw = 0, kw = 0;
    fsc_client_server_keys(c, &bsk, &bw, &ksk, &kw);
    Header h = make_header(2, fsc_client_params(c));
    h.aux = bw;      // split point of the payload
    return write_file(path, h, {{bsk, bw}, {ksk, kw}});
}

fsc_status fsc_server_keys_load(const char* path, fsc_params* params, uint64_t** bsk, size_t* bsk_words, uint64_t** ksk, size_t* ksk_words) {
    if (!path || !params || !bsk || !bsk_words || !ksk || !ksk_words) return fail(FSC_ERR_BAD_ARG, "null argument");
    *bsk = *ksk = nullptr; *bsk_words = *ksk_words = 0;
    Header h;
    std::vector<uint64_t> payload;
    fsc_status st = read_file(path, 2, h, payload);
    if (st != FSC_OK) return st;
    fsc_params p;
    memcpy(&p, h.params, sizeof(p));
    fsc_noise_params none;
    memset(&none, 0, sizeof(none));
    std::string why;
    if (!fsc_params_plausible(p, none, &why)) return fail(FSC_ERR_PARAMS, "key file header: " + why);
    if (h.aux > payload.size() || !words_match(p, h.aux, payload.size() - h.aux)) return fail(FSC_ERR_BAD_ARG, "key sizes do not match the parameters");
    uint64_t* buf = new (std::nothrow) uint64_t[payload.size() ? payload.size() : 1];
    if (!buf) return fail(FSC_ERR_OOM, "host allocation failed");
    memcpy(buf, payload.data(), payload.size() * 8);
    *params = p;
    *bsk = buf; *bsk_words = (size_t)h.aux;
    *ksk = buf + h.aux; *ksk_words = payload.size() - (size_t)h.aux;
    return FSC_OK;
}

fsc_status fsc_blocks_save(const char* path, const fsc_params* params, const uint64_t* blocks, size_t n_blocks) {
    if (!path || !params || (!blocks && n_blocks)) return fail(FSC_ERR_BAD_ARG, "null argument");
    Header h = make_header(3, *params);
    h.aux = n_blocks;
    return write_file(path, h, {{blocks, n_blocks * ((size_t)params->glwe_dim * params->poly_size + 1)}});
}

fsc_status fsc_blocks_load(const char* path, fsc_params* params, uint64_t** blocks, size_t* n_blocks) {
    if (!path || !params || !blocks || !n_blocks) return fail(FSC_ERR_BAD_ARG, "null argument");
    *blocks = nullptr; *n_blocks = 0;
    Header h;
    std::vector<uint64_t> payload;
    fsc_status st = read_file(path, 3, h, payload);
    if (st != FSC_OK) return st;
    fsc_params p;
    memcpy(&p, h.params, sizeof(p));
    fsc_noise_params none;
    memset(&none, 0, sizeof(none));
    std::string why;
    if (!fsc_params_plausible(p, none, &why)) return fail(FSC_ERR_PARAMS, "block file header: " + why);
    const uint64_t words = (uint64_t)p.glwe_dim * p.poly_size + 1;
    if (h.aux > ((uint64_t)1 << 34) / words || payload.size() != h.aux * words) return fail(FSC_ERR_BAD_ARG, "block count does not match the payload");
    uint64_t* buf = new (std::nothrow) uint64_t[payload.size() ? payload.size() : 1];
    if (!buf) return fail(FSC_ERR_OOM, "host allocation failed");
    memcpy(buf, payload.data(), payload.size() * 8);
    *params = p; *blocks = buf; *n_blocks = (size_t)h.aux;
    return FSC_OK;
}

/* frees a buffer returned by fsc_server_keys_load (pass the bsk pointer) or fsc_blocks_load */
fsc_status fsc_buffer_free(uint64_t* buffer) {
    delete[] buffer;
    return FSC_OK;
}

}  // extern "C"
