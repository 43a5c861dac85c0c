// oracle_backend.cpp — TEST INFRASTRUCTURE.  The radix layer's device side (fsc::RadixBackend) executed by the CPU
// oracle (oracle/tfhe_oracle.c): slots hold real big-LWE ciphertexts in host memory, a level is the linear
// combinations on the host followed by orc_ks_pbs (keyswitch + programmable bootstrap, OpenMP over the requests).
// Linked with the product's radix.cpp / radix_capi.cpp into tests/host/libfsc_orc.so it gives
//   * a second, ciphertext-carrying check of every radix circuit that does not need a GPU (tests/test_oracle_radix.py),
//   * the CPU timing of the reference's operator list (src/perf_test.rs:27-80; BASELINE.json configs[0]) and of
//     `k + e * d` (src/schnorr.rs:274) on the host cores: bench.py's cpu_baseline leg and --impl reference.
// Only tests/ and bench.py's CPU legs load it; the product never does.
#include <stdint.h>
#include <string.h>

#include <map>
#include <vector>

#include "../../fhe_sign_b200/csrc/ctx.h"
#include "../../oracle/tfhe_oracle.h"

namespace {

class OracleBackend : public fsc::RadixBackend {
public:
    const orc_keys* K;
    const orc_params* p;
    size_t words, N;
    uint64_t delta;
    int nthreads;
    std::vector<std::vector<uint64_t>> slot;      // one big-LWE ciphertext per slot
    std::vector<int32_t> free_list;
    std::vector<uint64_t> lut_polys;              // [n_luts][N]
    std::map<fsc::LutTable, int32_t> lut_ids;

    OracleBackend(const orc_keys* keys, int threads) : K(keys), p(orc_keys_params(keys)), nthreads(threads) {
        N = p->poly_size;
        words = (size_t)p->glwe_dim * p->poly_size + 1;
        delta = orc_delta(p);
    }
    int32_t alloc_slot() override {
        if (!free_list.empty()) { int32_t s = free_list.back(); free_list.pop_back(); return s; }
        slot.emplace_back(words, 0);
        return (int32_t)slot.size() - 1;
    }
    void free_slot(int32_t s) override { free_list.push_back(s); }
    int32_t lut_id(const fsc::LutTable& t) override {
        auto it = lut_ids.find(t);
        if (it != lut_ids.end()) return it->second;
        uint64_t table[fsc::kSpace];
        for (int i = 0; i < fsc::kSpace; ++i) table[i] = t[i];
        const size_t id = lut_polys.size() / N;
        lut_polys.resize((id + 1) * N);
        orc_make_lut(p, table, lut_polys.data() + id * N);
        lut_ids[t] = (int32_t)id;
        return (int32_t)id;
    }
    void lincomb(const std::vector<std::pair<int32_t, int32_t>>& terms, int32_t cst, uint64_t* out) const {
        memset(out, 0, words * 8);
        for (const auto& t : terms) {
            const uint64_t c = (uint64_t)(int64_t)t.second;
            const uint64_t* s = slot[t.first].data();
            for (size_t w = 0; w < words; ++w) out[w] += c * s[w];
        }
        out[words - 1] += (uint64_t)(int64_t)cst * delta;
    }
    void run_level(const std::vector<fsc::LevelReq>& reqs) override {
        const size_t count = reqs.size();
        if (!count) return;
        std::vector<uint64_t> in(count * words), out(count * words);
        std::vector<uint32_t> idx(count);
        for (size_t i = 0; i < count; ++i) {
            lincomb(reqs[i].terms, reqs[i].cst, in.data() + i * words);
            idx[i] = (uint32_t)reqs[i].lut;
        }
        orc_ks_pbs(K, in.data(), count, lut_polys.data(), idx.data(), out.data(), nthreads);
        for (size_t i = 0; i < count; ++i) memcpy(slot[reqs[i].dst].data(), out.data() + i * words, words * 8);
    }
    void run_linear(const std::vector<fsc::LinReq>& reqs) override {
        std::vector<uint64_t> out(reqs.size() * words);
        for (size_t i = 0; i < reqs.size(); ++i) lincomb(reqs[i].terms, reqs[i].cst, out.data() + i * words);
        for (size_t i = 0; i < reqs.size(); ++i) memcpy(slot[reqs[i].dst].data(), out.data() + i * words, words * 8);
    }
    size_t words_per_block() const override { return words; }
    void import_blocks(const uint64_t* host, size_t n, const int32_t* slots) override {
        for (size_t i = 0; i < n; ++i) memcpy(slot[slots[i]].data(), host + i * words, words * 8);
    }
    void export_blocks(const int32_t* slots, size_t n, uint64_t* host) override {
        for (size_t i = 0; i < n; ++i) memcpy(host + i * words, slot[slots[i]].data(), words * 8);
    }
};

}  // namespace

extern "C" {

fsc_status fsc_map_exception(const std::exception&) { return FSC_ERR_INTERNAL; }

// keys: an orc_keys* owned by the caller (must outlive the context); threads: OpenMP threads per level (0 = all)
fsc_status fscorc_ctx_create(const void* keys, int32_t threads, fsc_ctx** out) {
    if (!keys || !out) return FSC_ERR_BAD_ARG;
    fsc_ctx* c = new fsc_ctx();
    OracleBackend* ob = new OracleBackend(static_cast<const orc_keys*>(keys), threads);
    c->rb = ob;
    c->ev = new fsc::Evaluator(ob);
    *out = c;
    return FSC_OK;
}
fsc_status fscorc_ctx_destroy(fsc_ctx* c) {
    if (!c) return FSC_ERR_BAD_ARG;
    delete c->ev;
    delete c->rb;
    delete c;
    return FSC_OK;
}
const char* fscorc_last_error(const fsc_ctx* c) { return c ? c->err.c_str() : ""; }
}
