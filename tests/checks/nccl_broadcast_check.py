"""torchrun worker (tests/test_gpu_configs.py::test_nccl_broadcast_fit_gives_every_rank_the_same_bits):
rank 0 fits, nngp_b200.dist.broadcast_fit ships the packed state over NCCL, every rank predicts the same rows and the
results are compared BITWISE against rank 0's; repeated with a second, larger fit on the same handles (buffer reuse)
and in 'ntk' mode (the state also carries M)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "nngp-src_b200"):
    sys.path.insert(0, str(p))

from nngp_b200 import _lib, synth  # noqa: E402
from nngp_b200 import dist as ndist  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for kt, cases in (("nngp", [(1500, 32, 2), (4200, 64, 3)]), ("ntk", [(700, 16, 2)])):
        h = _lib.Handle(depth=2, device=local, kernel_type=kt)
        for n, d, depth in cases:
            if h.cfg.depth != depth:
                h.close()
                h = _lib.Handle(depth=depth, device=local, kernel_type=kt)
            xtr, ytr, xte, _ = synth.make_problem(n, 1000, d)
            if rank == 0:
                h.fit(xtr, ytr)
            nn, dd, lam = ndist.broadcast_fit(h, src=0, chunk=300_000)      # many chunks: exercises the pipeline
            assert (nn, dd) == (n, d)
            mean, var = h.predict(xte)
            both = torch.from_numpy(np.stack([mean, var])).cuda()
            ref = both.clone()
            dist.broadcast(ref, src=0)
            same = bool(torch.equal(both.view(torch.int64), ref.view(torch.int64)))
            # sharded prediction with gather == rank 0's full prediction
            ms, vs = ndist.sharded_predict(h, xte, gather=True)
            same = same and np.array_equal(ms, ref[0].cpu().numpy()) and np.array_equal(vs, ref[1].cpu().numpy())
            flag = torch.tensor([int(same)], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = ok and bool(flag.item())
            if rank == 0:
                print(f"{kt} N={n} D={d} depth={depth}: all ranks bitwise equal: {bool(flag.item())}")
        h.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("NCCL_BROADCAST_OK" if ok else "NCCL_BROADCAST_MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
