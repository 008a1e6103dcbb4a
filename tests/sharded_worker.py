"""torchrun worker: time-sharded filter+smoother(+gradient) across ranks vs the unsharded result.
argv[1] = nccl (CUDA backend = product kernels) | gloo (CPU backend = oracle arithmetic; exercises the
exchange logic only)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import O, make_problem, pkg, rel_err  # noqa: E402


def main():
    backend = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    pkg()
    from pssgp_b200.dist import TimeShard
    T = 4001
    t, y, cov, ssm = make_problem("matern52", T, seed=4)
    bounds = np.linspace(0, T, world + 1).astype(int)
    lo, hi = bounds[rank], bounds[rank + 1]
    with torch.no_grad():
        rfm, rfP, rll = O.pkf(ssm, y[:, None], True)
        rsm, rsP = O.pks(ssm, rfm, rfP)
    xchg = None
    if backend in ("nccl", "peer"):
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=dev)
        ops = None
        if backend == "peer":
            from pssgp_b200.dist import PeerExchange
            xchg = PeerExchange(rank, world, dist, dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
        import cpu_shard_backend as ops
    to = lambda x: x.detach().to(dev).contiguous()
    P0, H, R = to(ssm.P0), to(ssm.H).reshape(-1), to(ssm.R).reshape(-1)
    Fs, Qs = to(ssm.Fs[lo:hi]), to(ssm.Qs[lo:hi])
    yd = torch.as_tensor(y[lo:hi]).to(dev)
    sh = TimeShard(rank, world, dist, backend=ops, exchange=xchg)
    fms, fPs, ll = sh.filter(P0, Fs, Qs, H, R, yd)
    g = torch.ones(1, dtype=torch.float64, device=dev)
    o = sh.smoother_and_grad(P0, Fs, Qs, H, R, yd, fms, fPs, g, want_grad=(backend != "gloo"))
    if backend == "peer":
        # several steps back to back: slots and sequence numbers wrap around
        for _ in range(5):
            ll5, sms5, sPs5, _ = sh.filter_smoother_grad(P0, Fs, Qs, H, R, yd, g)
        o["sms"], o["sPs"] = sms5, sPs5
        ll = ll5
    ok = (rel_err(fms.cpu(), rfm[lo:hi]) < 1e-9 and rel_err(fPs.cpu(), rfP[lo:hi]) < 1e-9
          and abs(float(ll) - float(rll)) <= 1e-9 * abs(float(rll))
          and rel_err(o["sms"].cpu(), rsm[lo:hi]) < 1e-9 and rel_err(o["sPs"].cpu(), rsP[lo:hi]) < 1e-9)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_OK" if float(flag) == 1.0 else "SHARDED_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
