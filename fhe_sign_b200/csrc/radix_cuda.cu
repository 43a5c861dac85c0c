// radix_cuda.cu — device block pool and level execution for the radix layer.
// A level (radix.h) = upload of a few index arrays, one lincomb launch producing the packed PBS
// inputs, one keyswitch launch, one PBS launch scattering its outputs into the pool.
#include <chrono>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <map>
#include <string>
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
        try { peer_disconnect(); } catch (...) {}
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

    // The allocation carries a 4 KB tail behind the slots: the flag words of the peer barrier (one IPC handle covers both).
    static constexpr size_t kFlagTail = 4096;
    uint64_t* flag_area() const { return pool + cap * words; }      // [0, 8): arrivals per rank; [8]: barrier timeout record
    void grow_pool(size_t new_cap) {
        if (peer_fixed) throw Error(FSC_ERR_OOM, "block pool exhausted: a peer-mapped pool cannot grow past the capacity given to fsc_peer_pool_export");
        eng->use();
        uint64_t* np = nullptr;
        FSC_CUDA_CHECK(cudaMalloc(&np, new_cap * words * 8 + kFlagTail));
        {
            const cudaError_t e = cudaMemsetAsync(np + new_cap * words, 0, kFlagTail, eng->stream);
            if (e != cudaSuccess) { cudaFree(np); FSC_CUDA_CHECK(e); }
        }
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
    struct LevelRange {      // NVTX: one range per PBS level, named by its width (and whether it is sharded)
        LevelRange(size_t width, bool sharded) {
            char name[64];
            snprintf(name, sizeof(name), sharded ? "level %zu (sharded)" : "level %zu", width);
            nvtxRangePushA(name);
        }
        ~LevelRange() { nvtxRangePop(); }
    };
    void run_level(const std::vector<LevelReq>& reqs) override {
        if (reqs.empty()) return;
        eng->use();
        LevelRange range(reqs.size(), exchange.active(reqs.size()));
        if (exchange.active(reqs.size())) {
            if (exchange.peer) run_level_peer(reqs); else run_level_sharded(reqs);
            return;
        }
        peer_dirty = true;      // local work since the last barrier: a peer must not overwrite slots this level may still read
        static const bool trace = getenv("FSC_LEVEL_TRACE") != nullptr;      // diagnostic: width and device time of every level (synchronises)
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (trace) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, eng->stream); }
        struct TraceEnd {
            cudaEvent_t a, b; cudaStream_t st; size_t w;
            ~TraceEnd() {
                if (!a) return;
                cudaEventRecord(b, st); cudaEventSynchronize(b);
                float ms = 0; cudaEventElapsedTime(&ms, a, b);
                fprintf(stderr, "level trace: width %zu %.3f ms (%.0f PBS/s)\n", w, ms, w / (ms * 1e-3));
                cudaEventDestroy(a); cudaEventDestroy(b);
            }
        } trace_end{e0, e1, eng->stream, reqs.size()};
        const auto h0 = std::chrono::steady_clock::now();
        ensure_stage_big(reqs.size());
        eng->ensure_scratch(reqs.size());
        const auto h1 = std::chrono::steady_clock::now();
        const Csr c = upload(reqs, true);
        const auto h2 = std::chrono::steady_clock::now();
        if (trace) fprintf(stderr, "   host: ensure buffers %.3f ms, index build + upload call %.3f ms\n",
                           std::chrono::duration<double, std::milli>(h1 - h0).count(), std::chrono::duration<double, std::milli>(h2 - h1).count());
        cudaEvent_t t1 = nullptr, t2 = nullptr, t3 = nullptr, t4 = nullptr;
        if (trace) { cudaEventCreate(&t1); cudaEventCreate(&t2); cudaEventCreate(&t3); cudaEventCreate(&t4); cudaEventRecord(t1, eng->stream); }
        launch_lincomb(pool, c.row_ptr, c.slot, c.coef, c.cst, delta, stage_big, nullptr, (int)c.count, (int)words, eng->stream);
        ++eng->launches;
        FSC_CUDA_CHECK(cudaGetLastError());
        if (trace) cudaEventRecord(t2, eng->stream);
        eng->keyswitch(stage_big, eng->scratch_small, c.count);
        if (trace) cudaEventRecord(t3, eng->stream);
        eng->pbs(eng->scratch_small, &luts, c.lut, pool, c.count, c.dst);
        if (trace) {
            cudaEventRecord(t4, eng->stream); cudaEventSynchronize(t4);
            float a = 0, b = 0, d = 0, h = 0;
            cudaEventElapsedTime(&h, e0, t1); cudaEventElapsedTime(&a, t1, t2); cudaEventElapsedTime(&b, t2, t3); cudaEventElapsedTime(&d, t3, t4);
            fprintf(stderr, "   host prep + index upload %.3f ms | linear combinations %.3f | keyswitch %.3f | blind rotation %.3f\n", h, a, b, d);
            cudaEventDestroy(t1); cudaEventDestroy(t2); cudaEventDestroy(t3); cudaEventDestroy(t4);
        }
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
    // ---- fused exchange over peer-mapped pools ---------------------------------------------------------------
    // Rank r bootstraps its contiguous slice; the kernel's sample-extraction epilogue stores every output word into the
    // destination slot of EVERY rank's pool (own + peers over NVLink), so the next level finds its inputs in place.
    //   barrier A (only if this rank did local-only work since the last barrier): every rank has finished all earlier
    //              kernels, so no rank still reads a slot that a faster peer is about to overwrite;
    //   lincomb + keyswitch + blind rotation of the slice (destinations: all pools);
    //   barrier B: every rank's slice has landed in this rank's pool.
    void peer_barrier() {
        launch_peer_barrier(peer_flags, flag_area() + 8, exchange.rank, exchange.world, ++peer_seq, eng->stream);
        ++eng->launches;
        FSC_CUDA_CHECK(cudaGetLastError());
    }
    void run_level_peer(const std::vector<LevelReq>& reqs) {
        size_t per, lo, hi;
        shard_range(reqs.size(), exchange.rank, exchange.world, &per, &lo, &hi);
        const size_t mine = hi - lo;
        ensure_stage_big(std::max<size_t>(mine, 1));
        eng->ensure_scratch(std::max<size_t>(mine, 1));
        const Csr c = upload(reqs, true);
        if (peer_dirty) peer_barrier();
        if (mine) {
            launch_lincomb(pool, c.row_ptr + lo, c.slot, c.coef, c.cst + lo, delta, stage_big, nullptr, (int)mine, (int)words, eng->stream);
            ++eng->launches;
            FSC_CUDA_CHECK(cudaGetLastError());
            eng->keyswitch(stage_big, eng->scratch_small, mine);
            eng->pbs(eng->scratch_small, &luts, c.lut + lo, pool, mine, c.dst + lo, &peer_dests);
        }
        peer_barrier();
        peer_dirty = false;
        ++sharded_levels;
    }
    void check_peer_error() {
        if (!exchange.peer) return;
        uint64_t bad = 0;
        FSC_CUDA_CHECK(cudaMemcpyAsync(&bad, flag_area() + 8, 8, cudaMemcpyDeviceToHost, eng->stream));
        FSC_CUDA_CHECK(cudaStreamSynchronize(eng->stream));
        if (bad) throw Error(FSC_ERR_COMM, "level exchange: a peer did not reach barrier " + std::to_string(bad) + " in time (ranks out of step, or a rank died)");
    }

    struct PeerHandle {                  // the 128 opaque bytes of fsc_peer_pool_export
        cudaIpcMemHandle_t ipc;          // 64 bytes
        uint64_t pid, ptr, capacity, words;
        int32_t device, pad;
    };
    static_assert(sizeof(PeerHandle) <= kPeerHandleBytes, "peer handle blob");

    void peer_export(size_t capacity_blocks, uint8_t* out) override {
        eng->use();
        if (exchange.peer) throw Error(FSC_ERR_BAD_ARG, "peer pool already connected");
        FSC_CUDA_CHECK(cudaStreamSynchronize(eng->stream));
        peer_fixed = false;
        if (cap < capacity_blocks) grow_pool(capacity_blocks);      // keeps the blocks already imported
        FSC_CUDA_CHECK(cudaStreamSynchronize(eng->stream));
        peer_fixed = true;
        PeerHandle h;
        memset(&h, 0, sizeof(h));
        FSC_CUDA_CHECK(cudaIpcGetMemHandle(&h.ipc, pool));
        h.pid = (uint64_t)getpid(); h.ptr = (uint64_t)(uintptr_t)pool; h.capacity = cap; h.words = words; h.device = eng->dev;
        memset(out, 0, kPeerHandleBytes);
        memcpy(out, &h, sizeof(h));
    }
    void peer_connect(int rank, int world, size_t min_width, const uint8_t* handles) override {
        eng->use();
        if (!peer_fixed) throw Error(FSC_ERR_BAD_ARG, "call fsc_peer_pool_export before fsc_peer_pool_connect");
        if (exchange.peer) throw Error(FSC_ERR_BAD_ARG, "peer pool already connected");
        if (world > kMaxPeers) throw Error(FSC_ERR_BAD_ARG, "at most 8 ranks");
        FSC_CUDA_CHECK(cudaStreamSynchronize(eng->stream));
        uint64_t* pools[kMaxPeers] = {};
        for (int q = 0; q < world; ++q) {
            PeerHandle h;
            memcpy(&h, handles + (size_t)q * kPeerHandleBytes, sizeof(h));
            if (h.capacity != cap || h.words != words) throw Error(FSC_ERR_BAD_ARG, "peer pools must have the same capacity and block size on every rank");
            if (q == rank) {
                if (h.ptr != (uint64_t)(uintptr_t)pool) throw Error(FSC_ERR_BAD_ARG, "handle at index `rank` is not this context's own");
                pools[q] = pool;
            } else if (h.pid == (uint64_t)getpid()) {      // another context of this process: plain peer access
                int can = 0;
                FSC_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, eng->dev, h.device));
                if (!can) throw Error(FSC_ERR_COMM, "no peer access between the two devices");
                const cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) FSC_CUDA_CHECK(e);
                (void)cudaGetLastError();
                pools[q] = reinterpret_cast<uint64_t*>((uintptr_t)h.ptr);
            } else {
                void* p = nullptr;
                const cudaError_t e = cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {
                    for (void* o : peer_opened) cudaIpcCloseMemHandle(o);
                    peer_opened.clear();
                    throw Error(FSC_ERR_COMM, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
                }
                peer_opened.push_back(p);
                pools[q] = static_cast<uint64_t*>(p);
            }
        }
        // destination 0 is the local pool; the peers follow in ring order from this rank so that simultaneous epilogues
        // spread over different NVLink targets
        for (int i = 0; i < kMaxPeers; ++i) { peer_dests.base[i] = nullptr; peer_flags[i] = nullptr; }
        peer_dests.n = world;
        for (int i = 0; i < world; ++i) peer_dests.base[i] = pools[(rank + i) % world];
        for (int q = 0; q < world; ++q) peer_flags[q] = pools[q] + cap * words;
        exchange.rank = rank; exchange.world = world; exchange.min_width = min_width;
        exchange.all_gather = nullptr; exchange.buffer = nullptr; exchange.capacity = 0; exchange.peer = world > 1;
        peer_seq = 0; peer_dirty = true;
    }
    void peer_disconnect() override {
        if (!exchange.peer && peer_opened.empty()) { return; }
        eng->use();
        cudaStreamSynchronize(eng->stream);
        for (void* o : peer_opened) cudaIpcCloseMemHandle(o);
        peer_opened.clear();
        exchange = Exchange();
        // the pool may grow again - the caller guarantees that every rank has disconnected before any of them allocates
        // (fhe_sign_b200/distributed.py: a host barrier on both sides of the call)
        peer_fixed = false;
    }

    void run_linear(const std::vector<LinReq>& reqs) override {
        if (reqs.empty()) return;
        eng->use();
        peer_dirty = true;
        const Csr c = upload(reqs, false);
        launch_lincomb(pool, c.row_ptr, c.slot, c.coef, c.cst, delta, pool, c.dst, (int)c.count, (int)words, eng->stream);
        ++eng->launches;
        FSC_CUDA_CHECK(cudaGetLastError());
    }

    // ---- host transfer -----------------------------------------------------------------------------
    size_t words_per_block() const override { return words; }
    // One contiguous copy + one scatter / gather launch per radix value instead of one cudaMemcpyAsync per block
    // (a signature moves 385 blocks in and 272 out).
    const int32_t* upload_indices(const int32_t* idx, size_t n) {
        char* h = acquire_pinned(n * 4);
        ensure_dev_idx(n * 4);
        memcpy(h, idx, n * 4);
        FSC_CUDA_CHECK(cudaMemcpyAsync(dev_idx, h, n * 4, cudaMemcpyHostToDevice, eng->stream));
        FSC_CUDA_CHECK(cudaEventRecord(stage_ev[cur], eng->stream));
        return reinterpret_cast<const int32_t*>(dev_idx);
    }
    void import_blocks(const uint64_t* host, size_t n, const int32_t* slots) override {
        if (!n) return;
        eng->use();
        ensure_stage_big(n);
        FSC_CUDA_CHECK(cudaMemcpyAsync(stage_big, host, n * words * 8, cudaMemcpyHostToDevice, eng->stream));
        const int32_t* d = upload_indices(slots, n);
        launch_scatter(stage_big, d, pool, (int)n, (int)words, eng->stream);
        ++eng->launches;
        FSC_CUDA_CHECK(cudaGetLastError());
        FSC_CUDA_CHECK(cudaStreamSynchronize(eng->stream));      // the caller may release `host` on return
        peer_dirty = true;
    }
    void export_blocks(const int32_t* slots, size_t n, uint64_t* host) override {
        if (!n) return;
        eng->use();
        ensure_stage_big(n);
        const int32_t* d = upload_indices(slots, n);
        launch_gather(pool, d, stage_big, (int)n, (int)words, eng->stream);
        ++eng->launches;
        FSC_CUDA_CHECK(cudaGetLastError());
        FSC_CUDA_CHECK(cudaMemcpyAsync(host, stage_big, n * words * 8, cudaMemcpyDeviceToHost, eng->stream));
        FSC_CUDA_CHECK(cudaStreamSynchronize(eng->stream));
        check_peer_error();
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
    // peer-mapped pools (level sharding with the exchange fused into the blind rotation)
    bool peer_fixed = false;             // capacity frozen by fsc_peer_pool_export
    bool peer_dirty = true;              // local-only work enqueued since the last barrier
    uint64_t peer_seq = 0;               // barrier counter (identical on every rank: SPMD)
    OutDest peer_dests{};                // [0] own pool, then the peers' pools
    uint64_t* peer_flags[kMaxPeers] = {};
    std::vector<void*> peer_opened;      // cudaIpcOpenMemHandle mappings to close
};

}  // namespace

RadixBackend* make_cuda_backend(Engine* eng) { return new CudaBackend(eng); }

}  // namespace fsc
