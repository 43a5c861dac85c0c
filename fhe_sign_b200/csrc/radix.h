// radix.h — host-side scheduling of radix-integer operators over batched PBS levels.
//
// A radix integer is a little-endian vector of blocks, each an LWE ciphertext holding 2 message bits
// + 2 carry bits + 1 padding bit (PARAM_MESSAGE_2_CARRY_2): FheUint8/32/64 = 4/16/32 blocks, a
// 256-bit value = 128 blocks.  Every operator of the reference's hot path (src/biguint.rs:110-117,
// 135-143, 221-248; src/perf_test.rs:28-54) is expressed here as a short sequence of LEVELS; a level
// is a batch of independent "linear combination -> lookup table" requests and maps to one lincomb
// launch + one keyswitch launch + one PBS launch on the device (RadixBackend).  Nothing in this
// file touches ciphertext words: blocks are symbolic linear combinations of device slots, so
// additions, scalar multiplications, casts, block shifts and trivial constants cost no device work.
//
// The circuits are this engine's own; only decrypted results are contractual:
//   * carry propagation by parallel prefix - radix 2, or radix 3 (states as binary-adder digits, one lookup joins three
//     segments) wherever twice the bootstraps still fit one ciphertext per SM per GPU: a narrow level costs the same
//     whatever its width, so depth is what counts there;
//   * schoolbook partial products + carry-save column sums (an addend may ride in the sum: mul_add) + one propagation;
//   * comparison by a radix-3 tree of ordering codes, min / max selecting on the code directly;
//   * barrel shifter for encrypted amounts;
//   * division by a constant through a rounded-up magic multiplier (no fix-up when it fits the register width);
//   * remainder by 2^k - c with small c (the secp256k1 order) by folding hi 2^k + lo -> hi c + lo.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <array>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace fsc {

constexpr int kMsgBits = 2;
constexpr int kMsgMod = 4;          // message modulus
constexpr int kSpace = 16;          // message * carry modulus
constexpr int kMaxNoise = 5;        // max_noise_level of the parameter set
constexpr int kMaxVariance = 25;    // the same budget in variance units (nu^2): the default bookkeeping, see noise_in_variance_units()
constexpr size_t kBlocksPerGpuLevel = 148;   // widest PBS level that still runs at one ciphertext per SM (B200: 148 SMs)

struct RadixError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

using LutTable = std::array<uint8_t, kSpace>;

// One batched request: out = LUT[ sum_t coef_t * slot_t + cst ]   (cst in message units)
struct LevelReq {
    std::vector<std::pair<int32_t, int32_t>> terms;   // (slot, coefficient)
    int32_t cst = 0;
    int32_t lut = 0;                                    // index from RadixBackend::lut_id
    int32_t dst = -1;                                   // destination slot
};
// Pure linear materialisation: dst = sum_t coef_t * slot_t + cst   (no PBS)
struct LinReq {
    std::vector<std::pair<int32_t, int32_t>> terms;
    int32_t cst = 0;
    int32_t dst = -1;
};

// Multi-GPU: a level of at least min_width requests is cut into `world` contiguous slices of `per` requests;
// rank r bootstraps slice r into buffer[r * per ...] and the caller-provided all-gather completes the buffer on
// every rank (keys are replicated, so nothing else moves).  The callback enqueues the collective on the
// context's stream (NCCL over NVLink through torch.distributed in fhe_sign_b200/distributed.py).
typedef int32_t (*ExchangeFn)(void* user, void* buffer, size_t bytes_per_rank);
struct Exchange {
    int rank = 0, world = 1;
    size_t min_width = 0;
    void* buffer = nullptr;
    size_t capacity = 0;
    ExchangeFn all_gather = nullptr;
    void* user = nullptr;
    bool peer = false;                  // fused exchange: the PBS epilogue stores into every rank's peer-mapped pool (fsc_peer_*)
    bool enabled() const { return world > 1 && (all_gather || peer); }
    bool active(size_t count) const { return enabled() && count >= min_width; }
};
constexpr size_t kPeerHandleBytes = 128;
inline void shard_range(size_t count, int rank, int world, size_t* per, size_t* lo, size_t* hi) {
    *per = (count + world - 1) / world;
    *lo = std::min(count, (size_t)rank * *per);
    *hi = std::min(count, *lo + *per);
}

// Device (or mock) side of the radix layer.
class RadixBackend {
public:
    virtual ~RadixBackend() {}
    Exchange exchange;
    uint64_t sharded_levels = 0;
    virtual int32_t alloc_slot() = 0;
    virtual void free_slot(int32_t s) = 0;
    virtual int32_t lut_id(const LutTable& t) = 0;
    virtual void run_level(const std::vector<LevelReq>& reqs) = 0;
    virtual void run_linear(const std::vector<LinReq>& reqs) = 0;
    // host <-> slot transfer of whole LWE ciphertexts (big key, words_per_block() u64 words each)
    virtual size_t words_per_block() const = 0;
    virtual void import_blocks(const uint64_t* host, size_t n, const int32_t* slots) = 0;
    virtual void export_blocks(const int32_t* slots, size_t n, uint64_t* host) = 0;
    // peer-mapped block pools (CUDA backend only): see fsc_peer_pool_export / _connect in include/fhe_sign_cuda.h
    virtual void peer_export(size_t, uint8_t*) { throw RadixError("peer-mapped pools need the CUDA backend"); }
    virtual void peer_connect(int, int, size_t, const uint8_t*) { throw RadixError("peer-mapped pools need the CUDA backend"); }
    virtual void peer_disconnect() {}
    // statistics
    uint64_t pbs_count = 0, level_count = 0;
};

struct SlotRef {
    RadixBackend* be;
    int32_t idx;
    int32_t nv = 1;                     // noise variance of the ciphertext in this slot, in units of one fresh PBS output
    SlotRef(RadixBackend* b, int32_t i) : be(b), idx(i) {}
    ~SlotRef() { be->free_slot(idx); }
    SlotRef(const SlotRef&) = delete;
    SlotRef& operator=(const SlotRef&) = delete;
};
using SlotP = std::shared_ptr<SlotRef>;

// A block value = sum coef * slot + cst, with bookkeeping of the largest value it can take (deg)
// and of its noise in units of one fresh PBS output (nl).
struct Block {
    std::vector<std::pair<SlotP, int32_t>> terms;
    int32_t cst = 0;
    int32_t deg = 0;
    int32_t nl = 0;
    int32_t nv = 0;                     // noise variance in units of one fresh PBS output: sum over DISTINCT slots of coef^2 * slot variance
                                        // (recomputed from the merged coefficients: x + x is 4, not 2)

    bool trivial() const { return terms.empty(); }
    static Block constant(int v) { Block b; b.cst = v; b.deg = v; return b; }
    static Block from_slot(const SlotP& s, int deg, int nl = 1) {
        Block b; b.terms.emplace_back(s, 1); b.deg = deg; b.nl = nl; b.nv = nl; return b;
    }
};

Block operator+(const Block& a, const Block& b);
Block operator*(const Block& a, int c);            // c >= 0
Block complement(const Block& a, int top);         // top - a   (requires a <= top)
Block add_const(const Block& a, int c);
// Budget checks and column-sum chunking use sum c^2 <= 25 over independent fresh blocks (the parameter set's budget IS a variance:
// nu^2 = 25 fresh-PBS variances at the input of a lookup); FSC_RADIX_NOISE=linear falls back to sum |c| <= 5, the conservative rule
// of tfhe's NoiseLevel.  Measured on the GPU at the full 2_2 parameters with sum c^2 = 25 (4x + 3y) and with 7 unit terms: 0 decode
// failures in 102 400 lookups each, and the combination's noise is 0.7 % of what the lookup sees (keyswitch + modulus switch dominate).
bool noise_in_variance_units();
inline bool within_noise_budget(const Block& b) { return noise_in_variance_units() ? b.nv <= kMaxVariance : b.nl <= kMaxNoise; }

using Radix = std::vector<Block>;                   // little endian; blocks clean (deg <= 3) between operators

class Evaluator {
public:
    explicit Evaluator(RadixBackend* be) : be_(be) {}
    RadixBackend* backend() const { return be_; }

    // ---- level machinery ---------------------------------------------------------------
    struct Req { Block in; LutTable lut; };
    // Evaluates all requests as ONE level; trivial inputs are folded on the host.
    std::vector<Block> level(const std::vector<Req>& reqs);
    // Collapses every block to a single slot with coefficient 1 (for download).
    void materialize(Radix& r);
    // Identity-bootstraps every block that is not a single fresh slot (noise level back to 1).
    void clean(Radix& r);

    // ---- operators (inputs and outputs are clean radix integers) -------------------------
    Radix trivial_big(const std::vector<uint8_t>& blocks);
    Radix add(const Radix& a, const Radix& b, Block* carry_out = nullptr);
    Radix sub(const Radix& a, const Radix& b, Block* not_borrow = nullptr);
    Radix scalar_add(const Radix& a, const std::vector<uint8_t>& c);
    Radix mul(const Radix& a, const Radix& b, int out_blocks = -1);           // wrapping at out_blocks (default |a|)
    Radix mul_add(const Radix& a, const Radix& b, const Radix* addend, int out_blocks);   // a * b + addend, one propagation
    Radix scalar_mul(const Radix& a, const std::vector<uint8_t>& c, int out_blocks = -1);
    Radix scalar_mul_add(const Radix& a, const std::vector<uint8_t>& c, const Radix* addend, int out_blocks);
    Radix scalar_shr(const Radix& a, unsigned bits);
    Radix scalar_shl(const Radix& a, unsigned bits);
    Radix scalar_and(const Radix& a, const std::vector<uint8_t>& mask);
    Radix cast(const Radix& a, int n_blocks);
    Radix shr(const Radix& a, const Radix& amount);                            // amount taken mod bit width (power of two widths)
    Radix shl(const Radix& a, const Radix& amount);
    Block lt(const Radix& a, const Radix& b);                                  // encrypted bit a < b
    Block order_code(const Radix& a, const Radix& b);                          // 0 a < b, 1 equal, 2 a > b (radix-3 tree)
    Radix select_by_order(const Block& code, const Radix& if_lt, const Radix& otherwise);
    Block eq(const Radix& a, const Radix& b);
    Radix select(const Block& cond, const Radix& if_true, const Radix& if_false);
    Radix min(const Radix& a, const Radix& b);
    Radix max(const Radix& a, const Radix& b);
    Radix scalar_div(const Radix& a, const std::vector<uint8_t>& d, Radix* rem = nullptr);
    Radix scalar_rem(const Radix& a, const std::vector<uint8_t>& d);
    Radix bitop(const Radix& a, const Radix& b, int op);                       // 0 and, 1 or, 2 xor
    // multi-operand sum of clean radix integers, wrapping at n_blocks
    Radix sum(const std::vector<Radix>& operands, int n_blocks);

    // building blocks shared by the operators
    Radix sum_columns(std::vector<std::vector<Block>>& cols);                  // carry-save reduction + propagation
    Radix propagate(const std::vector<Block>& sums, Block* carry_out = nullptr);   // sums[i] <= 7 incl. carry-in
    Radix propagate_radix3(const std::vector<Block>& msg, std::vector<Block>& Y, std::vector<Block>& Z, Block* carry_out);
    int scan_world() const;

private:
    RadixBackend* be_;
    std::vector<Block> keep_;
};

// helpers on plain little-endian base-4 digit strings (host integers of any width)
std::vector<uint8_t> digits_from_u64(uint64_t v, int n_blocks);
std::vector<uint8_t> digits_from_bytes_le(const uint8_t* bytes, size_t n_bytes, int n_blocks);

}  // namespace fsc
