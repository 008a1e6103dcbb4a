"""Scratch timing of the C-ABI entry points with device-resident inputs (not the bench)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.import_package()
from pssgp_b200 import _lib
from pssgp_b200.kalman.parallel import pkf, pks

def main():
    N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    chunks = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    # synthetic stable LGSSM: F = expm-like contraction, Q small SPD
    A = torch.randn(d, d, dtype=torch.float64) * 0.3
    Fc = -(A @ A.T) - 0.5 * torch.eye(d, dtype=torch.float64) + (A - A.T)
    dts = 0.004 * (0.5 + torch.rand(N, dtype=torch.float64))
    Fs = torch.linalg.matrix_exp(dts[:, None, None] * Fc[None]).to(dev)
    Pinf = torch.eye(d, dtype=torch.float64)
    Pinf = torch.linalg.solve(torch.kron(torch.eye(d, dtype=torch.float64), Fc) + torch.kron(Fc, torch.eye(d, dtype=torch.float64)), -torch.eye(d, dtype=torch.float64).reshape(-1, 1)).reshape(d, d)
    Pinf = 0.5 * (Pinf + Pinf.T)
    Pd = Pinf.to(dev)
    Qs = Pd[None] - Fs @ Pd[None] @ Fs.transpose(1, 2)
    H = torch.zeros(1, d, dtype=torch.float64, device=dev); H[0, 0] = 1
    R = torch.tensor([[0.1]], dtype=torch.float64, device=dev)
    y = torch.randn(N, 1, dtype=torch.float64, device=dev)
    lg = (Pd, Fs, Qs, H, R)
    h = _lib.handle(0)
    for c in chunks:
        h.set_option("chunk", c)
        for name, fn in (("pkf", lambda: pkf(lg, y, True)),):
            for _ in range(3): out = fn()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): out = fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"chunk={c} {name}: {ms*1e3:.1f} us  {N/ms/1e6:.2f} G steps/s  alg {8*(3*d*d+d+1)*N/ms/1e6:.0f} GB/s")
        fm, fP, ll = out
        for _ in range(3): o2 = pks(lg, fm, fP)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): o2 = pks(lg, fm, fP)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"chunk={c} pks: {ms*1e3:.1f} us  {N/ms/1e6:.2f} G steps/s  alg {8*(4*d*d+2*d)*N/ms/1e6:.0f} GB/s")

main()
