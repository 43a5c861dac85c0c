"""Multi-GPU level sharding: one process per GPU, replicated keys.

Two exchange paths behind the same radix operators:
  * `enable_peer_sharding` (default of bench.py / the tools): the library's own exchange - every rank's block pool is
    mapped into its peers (CUDA IPC) and the blind rotation's epilogue stores its outputs straight into every pool over
    NVLink (fsc_peer_pool_export / fsc_peer_pool_connect).  Python is only involved in swapping the 128-byte handles at
    set-up; nothing of it runs in the level loop.
  * `enable_level_sharding`: the callback form - an NCCL all-gather of PBS level outputs enqueued by a Python callback
    (kept as the baseline the fused path is measured against, and for the gloo CPU tests).

`enable_level_sharding(ctx, ...)` allocates the exchange buffer as a torch tensor on the context's device and
installs an all-gather callback (fsc_set_level_exchange) that runs `torch.distributed.all_gather_into_tensor`
over it — NCCL over NVLink 5 / NVSwitch on a B200 node, gloo in the CPU tests.  Every rank must issue the same
operator sequence (SPMD); levels narrower than `min_width` run replicated with no communication.
"""
import ctypes as C

EXCHANGE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t)


def make_all_gather(buffer_tensor, group=None, stream=None):
    """Returns the C callback all-gathering equal slices of `buffer_tensor` in place (on `stream` if given)."""
    import contextlib
    import torch
    import torch.distributed as dist

    flat = buffer_tensor.view(torch.uint8).reshape(-1)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)

    def _cb(user, buf, bytes_per_rank):
        try:
            total = bytes_per_rank * world
            out = flat[:total]
            mine = flat[rank * bytes_per_rank:(rank + 1) * bytes_per_rank]
            if flat.is_cuda:
                with (torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()):
                    dist.all_gather_into_tensor(out, mine, group=group)
            else:
                # gloo: gather into a list of views, then copy back into place
                parts = [torch.empty(bytes_per_rank, dtype=torch.uint8) for _ in range(world)]
                dist.all_gather(parts, mine.clone(), group=group)
                for r, p in enumerate(parts):
                    flat[r * bytes_per_rank:(r + 1) * bytes_per_rank].copy_(p)
            return 0
        except Exception as e:      # never let an exception cross the C boundary
            print("level all-gather failed:", repr(e))
            return 1

    return EXCHANGE_FN(_cb)


def enable_level_sharding(ctx, stream, min_width=149, capacity_blocks=1 << 16, group=None):
    """Shard every PBS level of >= min_width requests across the ranks of the default process group.

    `stream` is the torch.cuda.Stream the context was created on (Context(..., stream=stream.cuda_stream)); it must
    not be the legacy default stream (handle 0 asks the library for a stream of its own).  The collective is
    enqueued on that stream, so kernels and all-gathers stay ordered without host synchronisation."""
    import torch
    import torch.distributed as dist

    assert stream.cuda_stream != 0, "create the context on a dedicated torch.cuda.Stream()"
    words = ctx.params.glwe_dim * ctx.params.poly_size + 1
    with torch.cuda.stream(stream):
        buf = torch.empty(capacity_blocks * words, dtype=torch.int64, device="cuda")
    cb = make_all_gather(buf, group, stream)
    rc = ctx.L.fsc_set_level_exchange(ctx.h, dist.get_rank(group), dist.get_world_size(group), min_width,
                                      C.c_void_p(buf.data_ptr()), buf.numel() * 8, cb, None)
    ctx._check(rc)
    ctx._exchange_keepalive = (buf, cb)      # the library keeps raw pointers to both
    return buf


PEER_HANDLE_BYTES = 128


def enable_peer_sharding(ctx, min_width=149, capacity_blocks=1 << 17, group=None):
    """Library-owned exchange over peer-mapped block pools (include/fhe_sign_cuda.h, fsc_peer_*).

    Every rank exports its pool handle, the handles are all-gathered once through torch.distributed (any host channel
    would do), every rank connects.  From then on a sharded level is: lincomb + keyswitch + blind rotation whose epilogue
    writes into all pools + one flag-barrier kernel - no callback, no NCCL, no Python."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = (C.c_uint8 * PEER_HANDLE_BYTES)()
    ctx._check(ctx.L.fsc_peer_pool_export(ctx.h, capacity_blocks, mine))
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t_mine = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=dev)
    t_all = torch.empty(world * PEER_HANDLE_BYTES, dtype=torch.uint8, device=dev)
    if dev == "cuda":
        dist.all_gather_into_tensor(t_all, t_mine, group=group)
    else:
        parts = [torch.empty(PEER_HANDLE_BYTES, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(parts, t_mine, group=group)
        t_all = torch.cat(parts)
    blob = bytes(t_all.cpu().numpy().tobytes())
    handles = (C.c_uint8 * (world * PEER_HANDLE_BYTES)).from_buffer_copy(blob)
    ctx._check(ctx.L.fsc_peer_pool_connect(ctx.h, rank, world, min_width, handles))
    dist.barrier(group=group)      # nobody starts storing into a pool that a slower rank has not mapped yet (set-up only)


def disable_peer_sharding(ctx, group=None):
    import torch.distributed as dist
    ctx.sync()
    dist.barrier(group=group)      # every rank is done writing into every pool
    ctx._check(ctx.L.fsc_peer_pool_disconnect(ctx.h))
    dist.barrier(group=group)      # every rank has unmapped every pool: they may be reallocated from here on
