"""Row-sharded prediction over 1/2/4/8 GPUs, one process per GPU.

The reference has no multi-device path at all (``nt.batch(device_count=0)``, train.py:166-168); the only
stage that shards is prediction: test rows are independent given the fitted state {X, L, alpha, lambda}.
So the fit runs on ONE rank, its state is broadcast once (``torch.distributed`` -- NCCL over NVLink 5 /
NVSwitch on GPUs, gloo in the CPU tests), and every rank predicts a contiguous row range with no further
inter-GPU traffic; an all-gather of (mean, var) -- 16 bytes per query -- is optional.  There is no data-path
collective to fuse with a kernel here: the broadcast happens once per fit.

``engine`` is anything with the ``_lib.Handle`` interface (dims / get_state / set_state / predict); the CPU
tests inject an oracle-backed fake so the plumbing is covered without a GPU.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous row range [lo, hi) of ``rank``: [r*T/G, (r+1)*T/G)."""
    return rank * total // world, (rank + 1) * total // world


def _dist():
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("nngp_b200.dist: torch.distributed is not initialised")
    return dist


def _comm_device(group=None):
    import torch
    dist = _dist()
    backend = dist.get_backend(group)
    if "nccl" in str(backend):
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


# doubles per broadcast chunk of the packed state (256 MB): large enough for NCCL to run at NVLink bandwidth, small
# enough that the two staging buffers cost 0.5 GB instead of a second copy of the factor
CHUNK = 32 * 1024 * 1024


def broadcast_fit(engine, src: int = 0, group=None, chunk: int = CHUNK):
    """Rank ``src`` holds a fitted engine; on return every rank's engine holds the same state.

    The state travels PACKED -- X, alpha, the lower triangle of L by rows (N(N+1)/2, half of the square), plus M in
    'ntk' mode -- in chunks: ``nngp_state_pack`` fills a staging chunk straight from the handle's buffers,
    ``dist.broadcast`` ships it, ``nngp_state_unpack`` stores it into the receiving handle's buffers; the layer-0
    diagonal and the diagonal-block inverses are recomputed on arrival.  Two staging chunks alternate so that the
    packing / unpacking of one overlaps the broadcast of the other.  Every hand-over between the library's stream and
    torch's communication stream is a host-side wait (the C calls are complete on return; the broadcast handle is
    waited for and the device synchronised before a chunk is unpacked or reused) -- there is no ordering between the
    two streams to rely on.  Traffic per receiving rank: 8*(N*D + N + N(N+1)/2) bytes.  Returns (N, D, lambda)."""
    import torch
    dist = _dist()
    rank = dist.get_rank(group)
    dev = _comm_device(group)
    hdr = torch.zeros(4, dtype=torch.float64, device=dev)
    if rank == src:
        n, d, lam = engine.dims()
        hdr[0], hdr[1], hdr[2], hdr[3] = n, d, lam, float(bool(getattr(engine, "is_ntk", False)))
    dist.broadcast(hdr, src=src, group=group)
    hdr_h = hdr.cpu()
    n, d, lam, ntk = int(hdr_h[0]), int(hdr_h[1]), float(hdr_h[2]), bool(hdr_h[3])
    if not hasattr(engine, "state_pack"):          # engines without the packed interface (test doubles)
        return _broadcast_fit_dense(engine, src, group, n, d, lam, ntk)
    if rank != src:
        engine.state_import_begin(n, d)
    total = engine.packed_size()
    chunk = max(1, min(int(chunk), total))
    stage = [torch.empty(chunk, dtype=torch.float64, device=dev) for _ in range(2 if total > chunk else 1)]
    sync = (lambda: torch.cuda.synchronize(dev)) if dev.type == "cuda" else (lambda: None)
    offs = list(range(0, total, chunk))
    pending = None                                  # (work, buffer index, offset, count) of the chunk in flight
    for i, off in enumerate(offs):
        cnt = min(chunk, total - off)
        buf = stage[i % len(stage)][:cnt]
        if rank == src:
            engine.state_pack(off, cnt, buf)        # complete on return (library stream synchronised)
        work = dist.broadcast(buf, src=src, group=group, async_op=True)
        if pending is not None:                     # finish the previous chunk while this one travels
            pw, pbuf, poff, pcnt = pending
            pw.wait()
            sync()
            if rank != src:
                engine.state_unpack(poff, pcnt, pbuf)
        if len(stage) == 1:
            work.wait()
            sync()
            if rank != src:
                engine.state_unpack(off, cnt, buf)
            pending = None
        else:
            pending = (work, buf, off, cnt)
    if pending is not None:
        pw, pbuf, poff, pcnt = pending
        pw.wait()
        sync()
        if rank != src:
            engine.state_unpack(poff, pcnt, pbuf)
    if rank != src:
        engine.state_import_end(lam)
    del stage
    return n, d, lam


def _broadcast_fit_dense(engine, src, group, n, d, lam, ntk):
    """Dense variant over get_state / set_state (full N x N factor)."""
    import torch
    dist = _dist()
    rank = dist.get_rank(group)
    dev = _comm_device(group)
    x = torch.empty((n, d), dtype=torch.float64, device=dev)
    l = torch.empty((n, n), dtype=torch.float64, device=dev)
    alpha = torch.empty(n, dtype=torch.float64, device=dev)
    m = torch.empty((n, n), dtype=torch.float64, device=dev) if ntk else None
    if rank == src:
        engine.get_state(out={"x": x, "l": l, "alpha": alpha, **({"m": m} if ntk else {})})
    for t in (x, l, alpha) + ((m,) if ntk else ()):
        dist.broadcast(t, src=src, group=group)
    if dev.type == "cuda":
        # NCCL broadcasts are only ENQUEUED on torch's communication stream; the handle copies these buffers on its
        # own non-blocking stream, which has no ordering with torch's streams -> the host must wait here, or a
        # receiving rank could import a partially received state.
        torch.cuda.synchronize(dev)
    if rank != src:
        if ntk:
            engine.set_state(x, l, alpha, lam, m=m)
        else:
            engine.set_state(x, l, alpha, lam)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    return n, d, lam


def sharded_predict(engine, x_test, want_var: bool = True, gather: bool = True, group=None):
    """Every rank passes the SAME x_test [T, D]; rank r predicts rows shard_bounds(T, G, r).

    gather=True: all ranks return full-length (mean[T], var[T]) (var None when want_var is False);
    gather=False: each rank returns only its own rows.  A row's result is bitwise independent of G
    (fixed-order reductions; tests/test_gpu_parity.py::test_row_blocking_and_sharding_are_bitwise_invariant)."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    total = x_test.shape[0]
    lo, hi = shard_bounds(total, world, rank)
    if hi > lo:
        mean, var = engine.predict(x_test[lo:hi], want_var=want_var)
    else:
        mean, var = np.empty(0), (np.empty(0) if want_var else None)
    if not gather:
        return mean, var
    dev = _comm_device(group)
    width = max(shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world))
    cols = 2 if want_var else 1
    mine = torch.zeros((width, cols), dtype=torch.float64, device=dev)
    mine[: hi - lo, 0] = torch.as_tensor(np.asarray(mean), dtype=torch.float64).to(dev)
    if want_var:
        mine[: hi - lo, 1] = torch.as_tensor(np.asarray(var), dtype=torch.float64).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    full = np.empty((total, cols))
    for r, p in enumerate(parts):
        a, b = shard_bounds(total, world, r)
        full[a:b] = p[: b - a].cpu().numpy()
    return full[:, 0].copy(), (full[:, 1].copy() if want_var else None)
