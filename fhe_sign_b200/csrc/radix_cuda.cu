// radix_cuda.cu — device block pool and level execution for the radix layer.
// A level (radix.h) = upload of a few index arrays, one lincomb launch producing the packed PBS
// inputs, one keyswitch launch, one PBS launch scattering its outputs into the pool.
#include <cuda_runtime.h>
#include <string.h>

#include <map>
#include <vector>

#include "engine.h"
#include "radix_cuda.h"

namespace fsc {

namespace {

constexpr int kStages = 4;          // pinned staging buffers in flight

class CudaBackend : public RadixBackend {
public:
    explicit CudaBackend(Engine* e) : eng(e) {
        words = (size_t)e->p.glwe_dim * e->p.poly_size + 1;
        delta = ((uint64_t)1 << 63) / (e->p.message_modulus * e->p.carry_modulus);
        e->use();
        for (int i = 0; i < kStages; ++i) FSC_CUDA_CHECK(cudaEventCreateWithFlags(&stage_ev[i], cudaEventDisableTiming));
    }
    ~CudaBackend() override {
        cudaSetDevice(eng->dev);
        cudaStreamSynchronize(eng->stream);
        if (pool) cudaFree(pool);
        if (luts.d) cudaFree(luts.d);
        if (stage_big) cudaFree(stage_big);
        if (dev_idx) cudaFree(dev_idx);
        for (int i = 0; i < kStages; ++i) {
            if (pinned[i]) cudaFreeHost(pinned[i]);
            if (stage_ev[i]) cudaEventDestroy(stage_ev[i]);
        }
    }

    // ---- slots ----------------------------------------------------------------------------
    int32_t alloc_slot() override {
        if (!free_list.empty()) { int32_t s = free_list.back(); free_list.pop_back(); return s; }
        if (next == cap) grow_pool(cap ? cap * 2 : 4096);
        return (int32_t)next++;
    }
    void free_slot(int32_t s) override { free_list.push_back(s); }

    void grow_pool(size_t new_cap) {
        eng->use();
        uint64_t* np = nullptr;
        FSC_CUDA_CHECK(cudaMalloc(&np, new_cap * words * 8));
        if (pool) {
            cudaError_t e = cudaMemcpyAsync(np, pool, next * words * 8, cudaMemcpyDeviceToDevice, eng->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(eng->stream);
            if (e != cudaSuccess) { cudaFree(np); FSC_CUDA_CHECK(e); }
            cudaFree(pool);
        }
        pool = np; cap = new_cap;
    }

    // ---- lookup tables ----------------------------------------------------------------------
    int32_t lut_id(const LutTable& t) override {
        auto it = lut_ids.find(t);
        if (it != lut_ids.end()) return it->second;
        eng->use();
        const size_t N = eng->p.poly_size;
        if (luts.n == lut_cap) {
            const size_t nc = lut_cap ? lut_cap * 2 : 256;
            uint64_t* nd = nullptr;
            FSC_CUDA_CHECK(cudaMalloc(&nd, nc * N * 8));
            if (luts.d) {
                cudaError_t e = cudaMemcpyAsync(nd, luts.d, luts.n * N * 8, cudaMemcpyDeviceToDevice, eng->stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(eng->stream);
                if (e != cudaSuccess) { cudaFree(nd); FSC_CUDA_CHECK(e); }
                cudaFree(luts.d);
            }
            luts.d = nd; lut_cap = nc;
        }
        std::vector<uint64_t> table(kSpace), poly(N);
        for (int i = 0; i < kSpace; ++i) table[i] = t[i];
        build_lut_poly(eng->p, table.data(), poly.data());
        // pageable source: the copy is staged before the call returns, `poly` may die afterwards
        FSC_CUDA_CHECK(cudaMemcpyAsync(luts.d + luts.n * N, poly.data(), N * 8, cudaMemcpyHostToDevice, eng->stream));
        FSC_CUDA_CHECK(cudaStreamSynchronize(eng->stream));
        const int32_t id = (int32_t)luts.n++;
        lut_ids[t] = id;
        return id;
    }

    // ---- staging ------------------------------------------------------------------------------
    // returns a pinned buffer of at least `bytes`, safe to overwrite
    char* acquire_pinned(size_t bytes) {
        cur = (cur + 1) % kStages;
        FSC_CUDA_CHECK(cudaEventSynchronize(stage_ev[cur]));
        if (bytes > pinned_cap[cur]) {
            if (pinned[cur]) cudaFreeHost(pinned[cur]);
            pinned[cur] = nullptr; pinned_cap[cur] = 0;
            size_t c = 1 << 16;
            while (c < bytes) c *= 2;
            FSC_CUDA_CHECK(cudaMallocHost(&pinned[cur], c));
            pinned_cap[cur] = c;
        }
        return static_cast<char*>(pinned[cur]);
    }
    void ensure_dev_idx(size_t bytes) {
        if (bytes <= dev_idx_cap) return;
        FSC_CUDA_CHECK(cudaStreamSynchronize(eng->stream));
        if (dev_idx) cudaFree(dev_idx);
        dev_idx = nullptr; dev_idx_cap = 0;
        size_t c = 1 << 16;
        while (c < bytes) c *= 2;
        FSC_CUDA_CHECK(cudaMalloc(&dev_idx, c));
        dev_idx_cap = c;
    }
    void ensure_stage_big(size_t count) {
        if (count <= stage_big_cap) return;
        FSC_CUDA_CHECK(cudaStreamSynchronize(eng->stream));
        if (stage_big) cudaFree(stage_big);
        stage_big = nullptr; stage_big_cap = 0;
        size_t c = 1024;
        while (c < count) c *= 2;
        FSC_CUDA_CHECK(cudaMalloc(&stage_big, c * words * 8));
        stage_big_cap = c;
    }

    struct Csr {
        size_t count = 0, terms = 0;
        const int32_t *row_ptr = nullptr, *slot = nullptr, *coef = nullptr, *cst = nullptr, *dst = nullptr;
        const uint32_t* lut = nullptr;
    };

    template <class Req>
    Csr upload(const std::vector<Req>& reqs, bool with_lut) {
        size_t terms = 0;
        for (const auto& r : reqs) terms += r.terms.size();
        const size_t count = reqs.size();
        const size_t n_i32 = (count + 1) + 2 * terms + 3 * count;
        const size_t bytes = n_i32 * 4;
        char* h = acquire_pinned(bytes);
        ensure_dev_idx(bytes);
        int32_t* row_ptr = reinterpret_cast<int32_t*>(h);
        int32_t* slot = row_ptr + count + 1;
        int32_t* coef = slot + terms;
        int32_t* cst = coef + terms;
        int32_t* dst = cst + count;
        uint32_t* lut = reinterpret_cast<uint32_t*>(dst + count);
        size_t t = 0;
        for (size_t i = 0; i < count; ++i) {
            row_ptr[i] = (int32_t)t;
            for (const auto& e : reqs[i].terms) { slot[t] = e.first; coef[t] = e.second; ++t; }
            cst[i] = reqs[i].cst;
            dst[i] = reqs[i].dst;
            lut[i] = with_lut ? (uint32_t)lut_of(reqs[i]) : 0;
        }
        row_ptr[count] = (int32_t)t;
        FSC_CUDA_CHECK(cudaMemcpyAsync(dev_idx, h, bytes, cudaMemcpyHostToDevice, eng->stream));
        FSC_CUDA_CHECK(cudaEventRecord(stage_ev[cur], eng->stream));
        Csr c;
        c.count = count; c.terms = terms;
        const int32_t* d = reinterpret_cast<const int32_t*>(dev_idx);
        c.row_ptr = d; c.slot = d + count + 1; c.coef = c.slot + terms; c.cst = c.coef + terms; c.dst = c.cst + count;
        c.lut = reinterpret_cast<const uint32_t*>(c.dst + count);
        return c;
    }
    static int32_t lut_of(const LevelReq& r) { return r.lut; }
    static int32_t lut_of(const LinReq&) { return 0; }

    // ---- levels ---------------------------------------------------------------------------------
    void run_level(const std::vector<LevelReq>& reqs) override {
        if (reqs.empty()) return;
        eng->use();
        if (exchange.active(reqs.size())) { run_level_sharded(reqs); return; }
        ensure_stage_big(reqs.size());
        eng->ensure_scratch(reqs.size());
        const Csr c = upload(reqs, true);
        launch_lincomb(pool, c.row_ptr, c.slot, c.coef, c.cst, delta, stage_big, nullptr, (int)c.count, (int)words, eng->stream);
        ++eng->launches;
        FSC_CUDA_CHECK(cudaGetLastError());
        eng->keyswitch(stage_big, eng->scratch_small, c.count);
        eng->pbs(eng->scratch_small, &luts, c.lut, pool, c.count, c.dst);
    }
    // this rank bootstraps its contiguous slice into the exchange buffer, all ranks all-gather, every rank scatters
    void run_level_sharded(const std::vector<LevelReq>& reqs) {
        size_t per, lo, hi;
        shard_range(reqs.size(), exchange.rank, exchange.world, &per, &lo, &hi);
        const size_t slice_bytes = per * words * 8;
        if (slice_bytes * exchange.world > exchange.capacity) throw Error(FSC_ERR_COMM, "level exchange buffer too small for this level");
        const size_t mine = hi - lo;
        ensure_stage_big(std::max<size_t>(mine, 1));
        eng->ensure_scratch(std::max<size_t>(mine, 1));
        const Csr c = upload(reqs, true);
        uint64_t* xb = static_cast<uint64_t*>(exchange.buffer);
        if (mine) {
            launch_lincomb(pool, c.row_ptr + lo, c.slot, c.coef, c.cst + lo, delta, stage_big, nullptr, (int)mine, (int)words, eng->stream);
            ++eng->launches;
            FSC_CUDA_CHECK(cudaGetLastError());
            eng->keyswitch(stage_big, eng->scratch_small, mine);
            eng->pbs(eng->scratch_small, &luts, c.lut + lo, xb + lo * words, mine, nullptr);
        }
        if (exchange.all_gather(exchange.user, exchange.buffer, slice_bytes) != 0) throw Error(FSC_ERR_COMM, "level all-gather failed");
        launch_scatter(xb, c.dst, pool, (int)reqs.size(), (int)words, eng->stream);
        ++eng->launches;
        FSC_CUDA_CHECK(cudaGetLastError());
        ++sharded_levels;
    }
    void run_linear(const std::vector<LinReq>& reqs) override {
        if (reqs.empty()) return;
        eng->use();
        const Csr c = upload(reqs, false);
        launch_lincomb(pool, c.row_ptr, c.slot, c.coef, c.cst, delta, pool, c.dst, (int)c.count, (int)words, eng->stream);
        ++eng->launches;
        FSC_CUDA_CHECK(cudaGetLastError());
    }

    // ---- host transfer -----------------------------------------------------------------------------
    size_t words_per_block() const override { return words; }
    void import_blocks(const uint64_t* host, size_t n, const int32_t* slots) override {
        eng->use();
        for (size_t i = 0; i < n; ++i)
            FSC_CUDA_CHECK(cudaMemcpyAsync(pool + (size_t)slots[i] * words, host + i * words, words * 8, cudaMemcpyHostToDevice, eng->stream));
        FSC_CUDA_CHECK(cudaStreamSynchronize(eng->stream));
    }
    void export_blocks(const int32_t* slots, size_t n, uint64_t* host) override {
        eng->use();
        for (size_t i = 0; i < n; ++i)
            FSC_CUDA_CHECK(cudaMemcpyAsync(host + i * words, pool + (size_t)slots[i] * words, words * 8, cudaMemcpyDeviceToHost, eng->stream));
        FSC_CUDA_CHECK(cudaStreamSynchronize(eng->stream));
    }

private:
    Engine* eng;
    size_t words = 0;
    uint64_t delta = 0;
    uint64_t* pool = nullptr;
    size_t cap = 0, next = 0;
    std::vector<int32_t> free_list;
    Luts luts;
    size_t lut_cap = 0;
    std::map<LutTable, int32_t> lut_ids;
    uint64_t* stage_big = nullptr;
    size_t stage_big_cap = 0;
    void* dev_idx = nullptr;
    size_t dev_idx_cap = 0;
    void* pinned[kStages] = {nullptr, nullptr, nullptr, nullptr};
    size_t pinned_cap[kStages] = {0, 0, 0, 0};
    cudaEvent_t stage_ev[kStages] = {nullptr, nullptr, nullptr, nullptr};
    int cur = 0;
};

}  // namespace

RadixBackend* make_cuda_backend(Engine* eng) { return new CudaBackend(eng); }

}  // namespace fsc
