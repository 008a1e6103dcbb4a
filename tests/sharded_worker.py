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
    if backend in ("nccl", "peer"):
        # host-buffer step (TimeShard.series_step): ll, gradient w.r.t. the SDE and the posterior of the latent function
        # against the oracle differentiated through its own discretisation (stationary-Q form, like the CUDA path)
        with torch.no_grad():
            sde = cov.get_sde()
        Fo, Po, Ho, Ro = [x.clone().requires_grad_(True) for x in (sde.F, sde.P0, sde.H, ssm.R)]
        ossm = O.get_ssm_stationary(sde._replace(F=Fo, P0=Po, H=Ho), t[:, None], Ro)
        ofm, ofP, oll = O.pkf(ossm, y[:, None], True)
        gF, gP, gH, gR = torch.autograd.grad(oll, (Fo, Po, Ho, Ro))
        with torch.no_grad():
            osm, osP = O.pks(ossm, ofm.detach(), ofP.detach())
            omean = (osm @ sde.H.reshape(-1))[lo:hi]
            ovar = torch.einsum("i,kij,j->k", sde.H.reshape(-1), osP, sde.H.reshape(-1))[lo:hi]
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        t_prev = 0.0 if lo == 0 else float(t[lo - 1])
        for _ in range(2):
            (ll_s, dF, dPinf, dH, dR), (mean, var) = sh.series_step(to(sde.F), to(sde.P0), to(sde.H), R, pin(t[lo:hi]),
                                                                    pin(y[lo:hi]), t_prev)
        if backend == "peer":
            # the same step captured as one CUDA graph and replayed: identical results
            mp, vp = torch.empty(hi - lo, dtype=torch.float64).pin_memory(), torch.empty(hi - lo, dtype=torch.float64).pin_memory()
            replay = sh.capture_series_step(to(sde.F), to(sde.P0), to(sde.H), R, pin(t[lo:hi]), pin(y[lo:hi]), t_prev, (mp, vp))
            for _ in range(3):
                (ll_g, dF_g, dP_g, dH_g, dR_g), (mean_g, var_g) = replay()
            if not (float(ll_g) == float(ll_s) and torch.equal(dF_g, dF) and torch.equal(mean_g, mean) and torch.equal(var_g, var)):
                print(f"[rank {rank}] graph replay differs from the eager series step", file=sys.stderr)
                ok = False
        errs = {"ll": abs(float(ll_s) - float(oll)) / abs(float(oll)), "dF": rel_err(dF, gF), "dPinf": rel_err(dPinf, 0.5 * (gP + gP.T)),  # gradient w.r.t. a symmetric matrix: symmetric part
               
                "dH": rel_err(dH, gH.reshape(-1)), "dR": rel_err(dR, gR.reshape(-1)), "mean": rel_err(mean, omean),
                "var": rel_err(var, ovar)}
        if max(errs.values()) >= 1e-8 or errs["ll"] > 1e-9:
            print(f"[rank {rank}] series_step errors: {errs}", file=sys.stderr)
            ok = False
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_OK" if float(flag) == 1.0 else "SHARDED_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
