#include <stdio.h>
#include <stdlib.h>
// mock_backend.cpp — TEST INFRASTRUCTURE.  A plaintext stand-in for the device side of the radix
// layer (fsc::RadixBackend): every "ciphertext" is the integer it would decrypt to, modulo 32 (4
// plaintext bits + the padding bit), and a bootstrap is the negacyclic table lookup a real PBS
// performs.  Linked with the product's radix.cpp / radix_capi.cpp into tests/host/libfsc_mock.so, it
// lets the CPU test-suite run every radix circuit on thousands of inputs, and flags any request
// whose input would have touched the padding bit or exceeded the noise budget on real hardware.
#include <stdint.h>

#include <map>
#include <vector>

#include "../../fhe_sign_b200/csrc/ctx.h"

namespace {

class MockBackend : public fsc::RadixBackend {
public:
    std::vector<int> val;
    std::vector<int32_t> free_list;
    std::vector<fsc::LutTable> luts;
    std::map<fsc::LutTable, int32_t> lut_ids;
    uint64_t violations = 0, max_batch = 0;
    uint64_t delta = (uint64_t)1 << 59;

    int32_t alloc_slot() override {
        if (!free_list.empty()) { int32_t s = free_list.back(); free_list.pop_back(); return s; }
        val.push_back(0);
        return (int32_t)val.size() - 1;
    }
    void free_slot(int32_t s) override { free_list.push_back(s); }
    int32_t lut_id(const fsc::LutTable& t) override {
        auto it = lut_ids.find(t);
        if (it != lut_ids.end()) return it->second;
        luts.push_back(t);
        lut_ids[t] = (int32_t)luts.size() - 1;
        return (int32_t)luts.size() - 1;
    }
    int lin(const std::vector<std::pair<int32_t, int32_t>>& terms, int cst) const {
        long v = cst;
        for (auto& t : terms) v += (long)t.second * val[t.first];
        return (int)(((v % 32) + 32) % 32);
    }
    // same slicing as the CUDA backend, on a host buffer: one "ciphertext" = 2049 words, body last
    void run_level_sharded(const std::vector<fsc::LevelReq>& reqs) {
        size_t per, lo, hi;
        fsc::shard_range(reqs.size(), exchange.rank, exchange.world, &per, &lo, &hi);
        const size_t slice_bytes = per * 2049 * 8;
        if (slice_bytes * exchange.world > exchange.capacity) throw fsc::RadixError("level exchange buffer too small");
        uint64_t* xb = static_cast<uint64_t*>(exchange.buffer);
        for (size_t i = lo; i < hi; ++i) {
            const int v = lin(reqs[i].terms, reqs[i].cst);
            if (v >= 16) ++violations;
            const fsc::LutTable& t = luts[reqs[i].lut];
            const int o = v < 16 ? t[v] : (32 - t[v - 16]) % 32;
            for (size_t w = 0; w < 2048; ++w) xb[i * 2049 + w] = 0;
            xb[i * 2049 + 2048] = (uint64_t)o * delta;
        }
        if (exchange.all_gather(exchange.user, exchange.buffer, slice_bytes) != 0) throw fsc::RadixError("all-gather failed");
        for (size_t i = 0; i < reqs.size(); ++i) val[reqs[i].dst] = (int)((xb[i * 2049 + 2048] + delta / 2) / delta) % 32;
        ++sharded_levels;
    }
    void run_level(const std::vector<fsc::LevelReq>& reqs) override {
        if (getenv("FSC_MOCK_TRACE")) fprintf(stderr, "level %zu\n", reqs.size());      // level widths, for schedule studies
        if (exchange.active(reqs.size())) { run_level_sharded(reqs); return; }
        std::vector<int> out(reqs.size());
        for (size_t i = 0; i < reqs.size(); ++i) {
            const int v = lin(reqs[i].terms, reqs[i].cst);
            if (v >= 16) ++violations;                                  // padding bit set: a real PBS would negate
            const fsc::LutTable& t = luts[reqs[i].lut];
            out[i] = v < 16 ? t[v] : (32 - t[v - 16]) % 32;
        }
        for (size_t i = 0; i < reqs.size(); ++i) val[reqs[i].dst] = out[i];
        if (reqs.size() > max_batch) max_batch = reqs.size();
    }
    void run_linear(const std::vector<fsc::LinReq>& reqs) override {
        std::vector<int> out(reqs.size());
        for (size_t i = 0; i < reqs.size(); ++i) out[i] = lin(reqs[i].terms, reqs[i].cst);
        for (size_t i = 0; i < reqs.size(); ++i) val[reqs[i].dst] = out[i];
    }
    size_t words_per_block() const override { return 2049; }
    void import_blocks(const uint64_t* host, size_t n, const int32_t* slots) override {
        // the tests hand in trivial (noiseless, zero-mask) LWE ciphertexts: the body is value * delta
        for (size_t i = 0; i < n; ++i) val[slots[i]] = (int)((host[i * 2049 + 2048] + delta / 2) / delta) % 32;
    }
    void export_blocks(const int32_t* slots, size_t n, uint64_t* host) override {
        for (size_t i = 0; i < n; ++i) {
            for (size_t w = 0; w < 2048; ++w) host[i * 2049 + w] = 0;
            host[i * 2049 + 2048] = (uint64_t)val[slots[i]] * delta;
        }
    }
};

}  // namespace

extern "C" {

fsc_status fsc_map_exception(const std::exception&) { return FSC_ERR_INTERNAL; }

fsc_status fscmock_ctx_create(fsc_ctx** out) {
    fsc_ctx* c = new fsc_ctx();
    MockBackend* mb = new MockBackend();
    c->rb = mb;
    c->ev = new fsc::Evaluator(mb);
    *out = c;
    return FSC_OK;
}
fsc_status fscmock_ctx_destroy(fsc_ctx* c) {
    if (!c) return FSC_ERR_BAD_ARG;
    delete c->ev;
    delete c->rb;
    delete c;
    return FSC_OK;
}
const char* fscmock_last_error(const fsc_ctx* c) { return c ? c->err.c_str() : ""; }
// violations: bootstraps whose input had the padding bit set; live: slots currently allocated
fsc_status fscmock_counters(const fsc_ctx* c, uint64_t* violations, uint64_t* live_slots, uint64_t* max_batch) {
    const MockBackend* mb = static_cast<const MockBackend*>(c->rb);
    if (violations) *violations = mb->violations;
    if (live_slots) *live_slots = mb->val.size() - mb->free_list.size();
    if (max_batch) *max_batch = mb->max_batch;
    return FSC_OK;
}
}
