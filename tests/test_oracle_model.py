"""CPU checks of the model half of the oracle (oracle/pyg_restated.py) against independent naive restatements of the
same published PyG 1.4.2 semantics: per-node Python loops instead of gather/scatter tensors.
The oracle itself stays PARITY UNPINNED (PyG is not available offline); this guards the restatement against slips."""
import numpy as np
import torch

from oracle import pyg_restated as pr


def _rand_graph(n=23, e=140, R=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    et = torch.randint(0, R, (e,), generator=g)
    x = torch.rand(n, 6, generator=g, dtype=torch.float64)
    return x, ei, et


def test_rgcnconv_mean_over_all_in_edges_naive():
    """A.1: out_v = mean over ALL incoming edges (any relation) of x_src W_type, 0 if none, + x_v root + bias"""
    torch.manual_seed(1)
    conv = pr.RGCNConvRef(6, 32, 5, 4).double()
    x, ei, et = _rand_graph()
    got = conv(x, ei, et)
    W = torch.einsum("rb,bio->rio", conv.att, conv.basis)
    want = torch.zeros(x.shape[0], 32, dtype=torch.float64)
    for v in range(x.shape[0]):
        msgs = [x[int(ei[0, k])] @ W[int(et[k])] for k in range(ei.shape[1]) if int(ei[1, k]) == v]
        if msgs:
            want[v] = torch.stack(msgs).mean(0)
        want[v] = want[v] + x[v] @ conv.root + conv.bias
    assert torch.allclose(got, want, atol=1e-12)
    bound = 1.0 / np.sqrt(4 * 6)          # uniform(1/sqrt(num_bases * in_channels)) for every parameter
    for p in conv.parameters():
        assert float(p.abs().max()) <= bound


def test_global_sort_pool_naive():
    """A.4: per graph, rows sorted by the last channel descending, first k kept, zero rows when the graph is smaller"""
    g = torch.Generator().manual_seed(3)
    sizes = [7, 3, 12, 5]
    x = torch.rand(sum(sizes), 9, generator=g, dtype=torch.float64) * 2 - 1
    x[2, -1] = x[4, -1]                    # a tie inside graph 0: the lower node index comes first
    batch = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes)])
    for k in (4, 7, 10):
        got = pr.global_sort_pool(x, batch, k).view(len(sizes), k, 9)
        off = 0
        for i, s in enumerate(sizes):
            rows = x[off:off + s]
            order = sorted(range(s), key=lambda t: (-float(rows[t, -1]), t))
            want = torch.zeros(k, 9, dtype=torch.float64)
            for pos, t in enumerate(order[:k]):
                want[pos] = rows[t]
            assert torch.equal(got[i], want), (k, i)
            off += s


def test_arr_regulariser_naive():
    """train_eval.py:167-174: sum over layers and adjacent rating pairs of ||W_{r+1} - W_r||^2"""
    torch.manual_seed(5)
    m = pr.IGMCRef(4, (32, 32, 32, 32), 5, 4, 0.0).double()
    want = 0.0
    for c in m.convs:
        W = torch.einsum("rb,bio->rio", c.att, c.basis)
        for r in range(4):
            want = want + float(((W[r + 1] - W[r]) ** 2).sum())
    assert abs(float(pr.arr_regulariser(m)) - want) <= 1e-10 * max(1.0, want)


def test_dgcnn_rs_shapes_follow_the_reference_constructor():
    """models.py:65-85: dense_dim = (int((k-2)/2+1) - 5 + 1) * 32, Conv1d(1,16,97,97), Conv1d(16,32,5,1)"""
    for k in (10, 30, 31, 61):
        m = pr.DGCNN_RSRef(4, (32, 32, 32, 1), k, 5, 2)
        assert m.total_latent_dim == 97
        assert m.dense_dim == (int((k - 2) / 2 + 1) - 5 + 1) * 32
        assert tuple(m.conv1d_params1.weight.shape) == (16, 1, 97) and m.conv1d_params1.stride == (97,)
        assert tuple(m.conv1d_params2.weight.shape) == (32, 16, 5)
        x, ei, et = _rand_graph(n=40, e=200)
        batch = torch.tensor([0] * 18 + [1] * 22)
        ei = torch.cat([ei[:, :100] % 18, 18 + ei[:, 100:] % 22], 1)
        out = m.eval()(x[:, :4].float(), ei, et, batch)
        assert out.shape == (2,)
