"""CPU, world_size 2, gloo: the multi-GPU host logic (fit on rank 0 -> broadcast state -> row-sharded
predict -> gather).  The engine is an oracle-backed fake with the _lib.Handle interface, so the plumbing is
exercised without a GPU; the real engine runs the same code path under NCCL in bench.py."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class FakeEngine:
    """Handle-shaped object whose arithmetic is the CPU oracle (test double, never shipped)."""

    def __init__(self):
        self.state = None

    def fit(self, x, y):
        import nngp_oracle as oracle
        f = oracle.Fit(x, y)
        self.state = {"x": f.x, "l": f.c, "alpha": f.alpha, "lambda": f.lam}

    def dims(self):
        return self.state["x"].shape[0], self.state["x"].shape[1], self.state["lambda"]

    def get_state(self, out=None):
        import torch
        for k in ("x", "l", "alpha"):
            out[k].copy_(torch.from_numpy(np.ascontiguousarray(self.state[k])))
        return out

    def set_state(self, x, l, alpha, lam):
        self.state = {"x": x.numpy().copy(), "l": l.numpy().copy(), "alpha": alpha.numpy().copy(), "lambda": lam}

    def predict(self, xt, want_var=True):
        import scipy.linalg as sla
        import nngp_oracle as oracle
        ks = oracle.kernel_fn(xt, self.state["x"])
        mean = ks @ self.state["alpha"]
        if not want_var:
            return mean, None
        v = sla.solve_triangular(self.state["l"], ks.T, lower=True)
        return mean, oracle.final_diag(oracle.layer0_diag(xt)) - np.einsum("ij,ij->j", v, v)


class FakeNtkEngine(FakeEngine):
    """State-only double of a Handle(kernel_type='ntk'): {x, l, alpha, lambda, m}."""
    is_ntk = True

    def get_state(self, out=None):
        import torch
        for k in ("x", "l", "alpha", "m"):
            out[k].copy_(torch.from_numpy(np.ascontiguousarray(self.state[k])))
        return out

    def set_state(self, x, l, alpha, lam, m=None):
        self.state = {"x": x.numpy().copy(), "l": l.numpy().copy(), "alpha": alpha.numpy().copy(), "lambda": lam,
                      "m": m.numpy().copy()}


class FakePackedEngine(FakeEngine):
    """FakeEngine + the packed-state interface of _lib.Handle (nngp_state_pack / _unpack; include/nngp_b200.h):
    [ X (N*D) | alpha (N) | tril(L) by rows | ntk: M (N*N) ] -- so broadcast_fit's chunked pipeline runs on CPU."""

    def _flat(self):
        st = self.state
        n = st["x"].shape[0]
        parts = [st["x"].ravel(), st["alpha"].ravel(), st["l"][np.tril_indices(n)]]
        return np.concatenate(parts)

    def packed_size(self):
        n, d = (self.state["x"].shape if self.state else self._shape)
        return n * d + n + n * (n + 1) // 2

    def state_pack(self, off, cnt, dst):
        import torch
        dst.copy_(torch.from_numpy(self._flat()[off:off + cnt]))

    def state_import_begin(self, n, d):
        self.state, self._shape = None, (n, d)
        self._buf = np.full(n * d + n + n * (n + 1) // 2, np.nan)

    def state_unpack(self, off, cnt, src):
        self._buf[off:off + cnt] = src.numpy()

    def state_import_end(self, lam):
        n, d = self._shape
        l = np.zeros((n, n))
        l[np.tril_indices(n)] = self._buf[n * d + n:]
        self.state = {"x": self._buf[:n * d].reshape(n, d).copy(), "alpha": self._buf[n * d:n * d + n].copy(), "l": l,
                      "lambda": lam}


def _worker(rank, world, port, q):
    sys.path[:0] = [str(ROOT), str(ROOT / "nngp-src_b200"), str(ROOT / "oracle"), str(ROOT / "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from nngp_b200 import dist as ndist, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xtr, ytr, xte, _ = synth.make_problem(96, 37, 12)
    eng = FakeEngine()
    if rank == 0:
        eng.fit(xtr, ytr)
    n, d, lam = ndist.broadcast_fit(eng, src=0)
    mean, var = ndist.sharded_predict(eng, xte, want_var=True, gather=True)
    own_m, own_v = ndist.sharded_predict(eng, xte, want_var=True, gather=False)
    m_only, none = ndist.sharded_predict(eng, xte, want_var=False, gather=True)
    # 'ntk' engines carry a fourth state tensor (M): it must arrive on the other rank too
    ntk_eng = FakeNtkEngine()
    if rank == 0:
        ntk_eng.state = {"x": xtr, "l": np.tril(np.outer(ytr, ytr)), "alpha": ytr, "lambda": 0.25,
                         "m": np.arange(96.0 * 96.0).reshape(96, 96)}
    ndist.broadcast_fit(ntk_eng, src=0)
    ntk_ok = bool(np.array_equal(ntk_eng.state["m"], np.arange(96.0 * 96.0).reshape(96, 96))
                  and ntk_eng.state["lambda"] == 0.25 and np.array_equal(ntk_eng.state["alpha"], ytr))
    # the packed, chunked path (what the real handle uses): odd chunk sizes, one chunk, chunk > total
    for chunk in (1000, 777, 10 ** 9):
        pk = FakePackedEngine()
        if rank == 0:
            pk.state = dict(eng.state)
        ndist.broadcast_fit(pk, src=0, chunk=chunk)
        ntk_ok = ntk_ok and all(np.array_equal(pk.state[k], eng.state[k]) for k in ("x", "l", "alpha"))
        ntk_ok = ntk_ok and pk.state["lambda"] == eng.state["lambda"]
    q.put((rank, n, d, lam, mean, var, own_m, m_only, (none is None) and ntk_ok))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_partition():
    from nngp_b200.dist import shard_bounds
    for total in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 4, 8):
            b = [shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


@pytest.mark.timeout(300)
def test_broadcast_and_sharded_predict_world2():
    import torch.multiprocessing as mp
    import nngp_oracle as oracle
    from nngp_b200 import synth
    from nngp_b200.dist import shard_bounds
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    xtr, ytr, xte, _ = synth.make_problem(96, 37, 12)
    ref = oracle.Fit(xtr, ytr)
    rm, rv = ref.predict(xte)
    for rank, n, d, lam, mean, var, own_m, m_only, none_ok in got:
        assert (n, d) == (96, 12) and lam == ref.lam and none_ok
        assert np.allclose(mean, rm, rtol=1e-10, atol=0) and np.allclose(var, rv, rtol=1e-8, atol=0)
        lo, hi = shard_bounds(37, 2, rank)
        assert np.array_equal(own_m, mean[lo:hi]) and np.array_equal(m_only, mean)
    assert np.array_equal(got[0][4], got[1][4]) and np.array_equal(got[0][5], got[1][5])   # ranks agree bitwise
