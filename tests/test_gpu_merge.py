"""pssgp_merge_queries (the data preparation of predict_f) against the oracle's restatement of the reference's
_merge_sorted (pssgp/model.py:15-55) with NaN observations at the queries (:99) and the time differencing of
kernels/base.py:31-35: bit-exact (values are moved, the differences are single subtractions)."""
import numpy as np
import pytest
import torch

from util import O, pkg

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _oracle(ts, ys, q):
    a, b = torch.as_tensor(ts), torch.as_tensor(q)
    ya, yb = torch.as_tensor(ys)[:, None], torch.full((len(q), 1), float("nan"), dtype=a.dtype)
    fa, fb = torch.zeros(len(ts), dtype=torch.bool), torch.ones(len(q), dtype=torch.bool)
    t_all, y_all, flags = O.merge_sorted(a, b, (ya, yb), (fa, fb))
    dts = t_all - torch.cat([torch.zeros(1, dtype=a.dtype), t_all[:-1]])
    return t_all.numpy(), y_all.numpy().reshape(-1), dts.numpy(), np.nonzero(flags.numpy())[0]


@pytest.mark.parametrize("n,K", [(1, 1), (5, 3), (3, 5), (1000, 1000), (1000, 17), (17, 1000), (100_000, 250_000)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_merge_matches_reference_semantics(n, K, dtype):
    pkg()
    from pssgp_b200 import ops
    rng = np.random.RandomState(n + K)
    # a coarse lattice so that ties between training and query times (and duplicates inside each) occur
    ts = np.sort(rng.randint(0, 4 * (n + K), size=n)).astype(dtype) / 8
    q = np.sort(rng.randint(0, 4 * (n + K), size=K)).astype(dtype) / 8
    ys = rng.randn(n).astype(dtype)
    rt, ry, rd, ridx = _oracle(ts, ys, q)
    t_all, y_all, dts, q_idx = ops.merge_queries(torch.as_tensor(ts).to(DEV), torch.as_tensor(ys).to(DEV),
                                                 torch.as_tensor(q).to(DEV))
    assert np.array_equal(t_all.cpu().numpy(), rt)
    assert np.array_equal(y_all.cpu().numpy(), ry, equal_nan=True)
    assert np.array_equal(dts.cpu().numpy(), rd)
    assert np.array_equal(q_idx.cpu().numpy(), ridx)
    assert np.all(dts.cpu().numpy() >= 0)
