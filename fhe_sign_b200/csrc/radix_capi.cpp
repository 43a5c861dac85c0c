// radix_capi.cpp — C ABI of the radix-integer operators (include/fhe_sign_cuda.h, "radix" section).
// Backend agnostic: the same glue serves the CUDA engine and the CPU mock of the circuit tests.
#include <string.h>

#include <vector>

#include "ctx.h"

// NVTX ranges per radix operator (CUDA build only: the CPU mock / oracle builds of this file do not define FSC_NVTX), so that an
// ncu / nsys capture of an operator is self-describing: "radix mul 128x128" > "level 16384" > kernels.
#ifdef FSC_NVTX
#include <nvtx3/nvToolsExt.h>
#include <stdio.h>
namespace {
struct OpRange {
    OpRange(const char* what, size_t a, size_t b) {
        char name[96];
        snprintf(name, sizeof(name), "radix %s %zu x %zu blocks", what, a, b);
        nvtxRangePushA(name);
    }
    ~OpRange() { nvtxRangePop(); }
};
}  // namespace
#define FSC_OP_RANGE(what, a, b) OpRange fsc_op_range__(what, a, b)
#else
#define FSC_OP_RANGE(what, a, b) do { } while (0)
#endif

using fsc::Block;
using fsc::Radix;
using fsc::RadixError;

#define RX_BEGIN(ctx)                            \
    if (!(ctx)) return FSC_ERR_BAD_ARG;          \
    if (!(ctx)->ev) { (ctx)->err = "radix layer unavailable: server keys not uploaded"; return FSC_ERR_NO_KEYS; } \
    try {
#define RX_END(ctx)                                                                           \
    }                                                                                         \
    catch (const RadixError& e) { (ctx)->err = e.what(); return FSC_ERR_BAD_ARG; }            \
    catch (const std::bad_alloc&) { (ctx)->err = "host allocation failed"; return FSC_ERR_OOM; } \
    catch (const std::exception& e) { (ctx)->err = e.what(); return fsc_map_exception(e); }   \
    catch (...) { (ctx)->err = "unknown error"; return FSC_ERR_INTERNAL; }                    \
    return FSC_OK;

// engine errors carry their own status code; defined in fsc_api.cu (CUDA build) or mock_backend.cpp
extern "C" fsc_status fsc_map_exception(const std::exception& e);

static void need(bool c, const char* msg) { if (!c) throw RadixError(msg); }

static fsc_status emit(fsc_ctx* ctx, fsc_radix** out, Radix&& r) {
    fsc_radix* h = new fsc_radix();
    h->blocks = std::move(r);
    *out = h;
    (void)ctx;
    return FSC_OK;
}

static std::vector<uint8_t> scalar_digits(const uint8_t* bytes, size_t n_bytes, size_t n_blocks) {
    // all significant bits of the scalar are kept even when it is wider than the radix value
    size_t need_blocks = n_blocks;
    for (size_t i = n_bytes; i-- > 0;)
        if (bytes[i]) { need_blocks = std::max(need_blocks, (i + 1) * 4); break; }
    return fsc::digits_from_bytes_le(bytes, n_bytes, (int)need_blocks);
}

extern "C" {

fsc_status fsc_radix_from_lwe(fsc_ctx* ctx, const uint64_t* host_blocks, size_t n_blocks, fsc_radix** out) {
    RX_BEGIN(ctx)
    need(out && (host_blocks || !n_blocks), "null argument");
    std::vector<int32_t> slots(n_blocks);
    Radix r;
    for (size_t i = 0; i < n_blocks; ++i) {
        slots[i] = ctx->rb->alloc_slot();
        r.push_back(Block::from_slot(std::make_shared<fsc::SlotRef>(ctx->rb, slots[i]), 3, 1));
    }
    if (n_blocks) ctx->rb->import_blocks(host_blocks, n_blocks, slots.data());
    emit(ctx, out, std::move(r));
    RX_END(ctx)
}

fsc_status fsc_radix_to_lwe(fsc_ctx* ctx, fsc_radix* r, uint64_t* host_blocks) {
    RX_BEGIN(ctx)
    need(r && (host_blocks || r->blocks.empty()), "null argument");
    // trivial blocks become single-slot ciphertexts too (noiseless: zero mask, body = value * delta)
    ctx->ev->materialize(r->blocks);
    std::vector<int32_t> slots;
    for (auto& b : r->blocks) slots.push_back(b.terms[0].first->idx);
    if (!slots.empty()) ctx->rb->export_blocks(slots.data(), slots.size(), host_blocks);
    RX_END(ctx)
}

fsc_status fsc_radix_trivial(fsc_ctx* ctx, const uint8_t* value_le, size_t n_bytes, size_t n_blocks, fsc_radix** out) {
    RX_BEGIN(ctx)
    need(out && (value_le || !n_bytes), "null argument");
    emit(ctx, out, ctx->ev->trivial_big(fsc::digits_from_bytes_le(value_le, n_bytes, (int)n_blocks)));
    RX_END(ctx)
}

fsc_status fsc_radix_clone(fsc_ctx* ctx, const fsc_radix* a, fsc_radix** out) {
    RX_BEGIN(ctx)
    need(a && out, "null argument");
    Radix r = a->blocks;
    emit(ctx, out, std::move(r));
    RX_END(ctx)
}

fsc_status fsc_radix_free(fsc_ctx* ctx, fsc_radix* a) {
    if (!ctx) return FSC_ERR_BAD_ARG;
    try { delete a; } catch (...) {}
    return FSC_OK;
}

fsc_status fsc_radix_len(const fsc_radix* a, size_t* n_blocks) {
    if (!a || !n_blocks) return FSC_ERR_BAD_ARG;
    *n_blocks = a->blocks.size();
    return FSC_OK;
}

fsc_status fsc_radix_binary(fsc_ctx* ctx, uint32_t op, const fsc_radix* a, const fsc_radix* b, fsc_radix** out) {
    RX_BEGIN(ctx)
    need(a && b && out, "null argument");
#ifdef FSC_NVTX
    static const char* const kNames[] = {"add", "sub", "mul", "min", "max", "shr", "shl", "and", "or", "xor", "lt", "eq"};
    FSC_OP_RANGE(op < 12 ? kNames[op] : "?", a->blocks.size(), b->blocks.size());
#endif
    fsc::Evaluator& ev = *ctx->ev;
    Radix r;
    switch (op) {
        case FSC_OP_ADD: r = ev.add(a->blocks, b->blocks); break;
        case FSC_OP_SUB: r = ev.sub(a->blocks, b->blocks); break;
        case FSC_OP_MUL: r = ev.mul(a->blocks, b->blocks); break;
        case FSC_OP_MIN: r = ev.min(a->blocks, b->blocks); break;
        case FSC_OP_MAX: r = ev.max(a->blocks, b->blocks); break;
        case FSC_OP_SHR: r = ev.shr(a->blocks, b->blocks); break;
        case FSC_OP_SHL: r = ev.shl(a->blocks, b->blocks); break;
        case FSC_OP_AND: r = ev.bitop(a->blocks, b->blocks, 0); break;
        case FSC_OP_OR: r = ev.bitop(a->blocks, b->blocks, 1); break;
        case FSC_OP_XOR: r = ev.bitop(a->blocks, b->blocks, 2); break;
        case FSC_OP_LT: r = Radix{ev.lt(a->blocks, b->blocks)}; break;
        case FSC_OP_EQ: r = Radix{ev.eq(a->blocks, b->blocks)}; break;
        default: throw RadixError("unknown binary operator");
    }
    emit(ctx, out, std::move(r));
    RX_END(ctx)
}

fsc_status fsc_radix_scalar(fsc_ctx* ctx, uint32_t op, const fsc_radix* a, const uint8_t* scalar_le, size_t n_bytes, fsc_radix** out) {
    RX_BEGIN(ctx)
    need(a && out && (scalar_le || !n_bytes), "null argument");
#ifdef FSC_NVTX
    static const char* const kScalarNames[] = {"scalar add", "scalar sub", "scalar mul", "?", "?", "scalar shr", "scalar shl", "scalar and",
                                               "?", "?", "?", "?", "scalar div", "scalar rem"};
    FSC_OP_RANGE(op < 14 ? kScalarNames[op] : "?", a->blocks.size(), n_bytes * 4);
#endif
    fsc::Evaluator& ev = *ctx->ev;
    const size_t n = a->blocks.size();
    Radix r;
    switch (op) {
        case FSC_OP_ADD: r = ev.scalar_add(a->blocks, fsc::digits_from_bytes_le(scalar_le, n_bytes, (int)n)); break;
        case FSC_OP_MUL: r = ev.scalar_mul(a->blocks, fsc::digits_from_bytes_le(scalar_le, n_bytes, (int)n)); break;
        case FSC_OP_AND: r = ev.scalar_and(a->blocks, fsc::digits_from_bytes_le(scalar_le, n_bytes, (int)n)); break;
        case FSC_OP_DIV: r = ev.scalar_div(a->blocks, scalar_digits(scalar_le, n_bytes, n)); break;
        case FSC_OP_REM: r = ev.scalar_rem(a->blocks, scalar_digits(scalar_le, n_bytes, n)); break;
        case FSC_OP_SHR:
        case FSC_OP_SHL: {
            uint64_t bits = 0;
            for (size_t i = 0; i < n_bytes && i < 8; ++i) bits |= (uint64_t)scalar_le[i] << (8 * i);
            for (size_t i = 8; i < n_bytes; ++i) if (scalar_le[i]) bits = ~(uint64_t)0;
            // like the reference backend, a shift amount is taken modulo the bit width
            const uint64_t width = 2 * n;
            need(width > 0, "shift of an empty value");
            bits %= width;
            r = op == FSC_OP_SHR ? ev.scalar_shr(a->blocks, (unsigned)bits) : ev.scalar_shl(a->blocks, (unsigned)bits);
            break;
        }
        default: throw RadixError("unknown scalar operator");
    }
    emit(ctx, out, std::move(r));
    RX_END(ctx)
}

fsc_status fsc_radix_mul_wide(fsc_ctx* ctx, const fsc_radix* a, const fsc_radix* b, size_t out_blocks, fsc_radix** out) {
    RX_BEGIN(ctx)
    need(a && b && out, "null argument");
    FSC_OP_RANGE("mul_wide", a->blocks.size(), b->blocks.size());
    emit(ctx, out, ctx->ev->mul(a->blocks, b->blocks, (int)out_blocks));
    RX_END(ctx)
}

fsc_status fsc_radix_mul_add_wide(fsc_ctx* ctx, const fsc_radix* a, const fsc_radix* b, const fsc_radix* addend, size_t out_blocks,
                                  fsc_radix** out) {
    RX_BEGIN(ctx)
    need(a && b && addend && out, "null argument");
    FSC_OP_RANGE("mul_add_wide", a->blocks.size(), b->blocks.size());
    emit(ctx, out, ctx->ev->mul_add(a->blocks, b->blocks, &addend->blocks, (int)out_blocks));
    RX_END(ctx)
}

fsc_status fsc_radix_scalar_mul_add_wide(fsc_ctx* ctx, const fsc_radix* a, const uint8_t* scalar_le, size_t n_bytes, const fsc_radix* addend,
                                         size_t out_blocks, fsc_radix** out) {
    RX_BEGIN(ctx)
    need(a && out && (scalar_le || !n_bytes), "null argument");
    FSC_OP_RANGE("scalar_mul_add_wide", a->blocks.size(), n_bytes * 4);
    emit(ctx, out, ctx->ev->scalar_mul_add(a->blocks, scalar_digits(scalar_le, n_bytes, 0), addend ? &addend->blocks : nullptr, (int)out_blocks));
    RX_END(ctx)
}

fsc_status fsc_radix_cast(fsc_ctx* ctx, const fsc_radix* a, size_t n_blocks, fsc_radix** out) {
    RX_BEGIN(ctx)
    need(a && out, "null argument");
    emit(ctx, out, ctx->ev->cast(a->blocks, (int)n_blocks));
    RX_END(ctx)
}

fsc_status fsc_radix_slice(fsc_ctx* ctx, const fsc_radix* a, size_t first, size_t n_blocks, fsc_radix** out) {
    RX_BEGIN(ctx)
    need(a && out, "null argument");
    need(first <= a->blocks.size() && n_blocks <= a->blocks.size() - first, "slice exceeds the value");
    emit(ctx, out, Radix(a->blocks.begin() + first, a->blocks.begin() + first + n_blocks));
    RX_END(ctx)
}

fsc_status fsc_radix_concat(fsc_ctx* ctx, const fsc_radix* const* parts, size_t n_parts, fsc_radix** out) {
    RX_BEGIN(ctx)
    need(out && (parts || !n_parts), "null argument");
    Radix r;
    for (size_t i = 0; i < n_parts; ++i) {
        need(parts[i] != nullptr, "null part");
        r.insert(r.end(), parts[i]->blocks.begin(), parts[i]->blocks.end());
    }
    emit(ctx, out, std::move(r));
    RX_END(ctx)
}

fsc_status fsc_radix_sum(fsc_ctx* ctx, const fsc_radix* const* operands, size_t n_operands, size_t n_blocks, fsc_radix** out) {
    RX_BEGIN(ctx)
    need(out && (operands || !n_operands), "null argument");
    std::vector<Radix> ops;
    for (size_t i = 0; i < n_operands; ++i) {
        need(operands[i] != nullptr, "null operand");
        Radix r = operands[i]->blocks;
        ctx->ev->clean(r);
        ops.push_back(std::move(r));
    }
    emit(ctx, out, ctx->ev->sum(ops, (int)n_blocks));
    RX_END(ctx)
}

fsc_status fsc_radix_select(fsc_ctx* ctx, const fsc_radix* cond, const fsc_radix* if_true, const fsc_radix* if_false, fsc_radix** out) {
    RX_BEGIN(ctx)
    need(cond && if_true && if_false && out, "null argument");
    need(cond->blocks.size() == 1, "condition must be a single block");
    emit(ctx, out, ctx->ev->select(cond->blocks[0], if_true->blocks, if_false->blocks));
    RX_END(ctx)
}

fsc_status fsc_radix_stats(const fsc_ctx* ctx, uint64_t* pbs_count, uint64_t* level_count) {
    if (!ctx || !ctx->rb) return FSC_ERR_BAD_ARG;
    if (pbs_count) *pbs_count = ctx->rb->pbs_count;
    if (level_count) *level_count = ctx->rb->level_count;
    return FSC_OK;
}

fsc_status fsc_radix_stats2(const fsc_ctx* ctx, uint64_t* pbs_count, uint64_t* level_count, uint64_t* sharded_levels) {
    if (!ctx || !ctx->rb) return FSC_ERR_BAD_ARG;
    if (pbs_count) *pbs_count = ctx->rb->pbs_count;
    if (level_count) *level_count = ctx->rb->level_count;
    if (sharded_levels) *sharded_levels = ctx->rb->sharded_levels;
    return FSC_OK;
}

fsc_status fsc_set_level_exchange(fsc_ctx* ctx, int32_t rank, int32_t world, size_t min_width, void* buffer, size_t capacity_bytes,
                                  fsc_exchange_fn all_gather, void* user) {
    RX_BEGIN(ctx)
    need(world >= 1 && rank >= 0 && rank < world, "rank / world out of range");
    need(world == 1 || (buffer && all_gather && capacity_bytes), "exchange buffer and callback are required when world > 1");
    fsc::Exchange& x = ctx->rb->exchange;
    x.rank = rank; x.world = world; x.min_width = min_width; x.buffer = buffer; x.capacity = capacity_bytes;
    x.all_gather = all_gather; x.user = user;
    RX_END(ctx)
}

fsc_status fsc_peer_pool_export(fsc_ctx* ctx, size_t capacity_blocks, uint8_t* handle_out) {
    RX_BEGIN(ctx)
    need(handle_out && capacity_blocks, "null handle buffer or zero capacity");
    ctx->rb->peer_export(capacity_blocks, handle_out);
    RX_END(ctx)
}

fsc_status fsc_peer_pool_connect(fsc_ctx* ctx, int32_t rank, int32_t world, size_t min_width, const uint8_t* handles) {
    RX_BEGIN(ctx)
    need(world >= 1 && world <= 8 && rank >= 0 && rank < world, "rank / world out of range (at most 8 GPUs of one node)");
    need(handles, "null handle array");
    ctx->rb->peer_connect(rank, world, min_width, handles);
    RX_END(ctx)
}

fsc_status fsc_peer_pool_disconnect(fsc_ctx* ctx) {
    RX_BEGIN(ctx)
    ctx->rb->peer_disconnect();
    RX_END(ctx)
}

}  // extern "C"
