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


def test_transform_formulation_equals_bmm_formulation():
    """the memory-light message formulation used for BASELINE-size fp64 parity (x W_r first, then a row gather) is the
    same function as the reference-era per-edge weight gather + bmm: outputs and every gradient agree to 1e-12"""
    torch.manual_seed(2)
    x, ei, et = _rand_graph(n=40, e=400)
    x = x[:, :4]
    y = torch.rand(1, dtype=torch.float64)
    res = []
    for form in ("bmm", "transform"):
        torch.manual_seed(7)
        m = pr.set_formulation(pr.IGMCRef(4, (32, 32, 32, 32), 5, 4, 0.0).double().eval(), form)
        xx = torch.zeros_like(x)
        xx[0, 0] = 1.0
        xx[1, 1] = 1.0
        xx[2:, 2] = 1.0
        out = m(xx, ei, et)
        loss = ((out - y) ** 2).mean() + 0.001 * pr.arr_regulariser(m)
        loss.backward()
        res.append((out.detach(), [p.grad.clone() for p in m.parameters()]))
    assert torch.allclose(res[0][0], res[1][0], atol=1e-12)
    for a, b in zip(res[0][1], res[1][1]):
        assert float((a - b).abs().max()) <= 1e-12 * (1.0 + float(a.abs().max()))


def test_rgcnconv_backward_naive_per_edge():
    """gradients of one restated layer against a hand-written per-edge backward of A.1 (no autograd):
    with g = d out, c_v = 1 / max(in-degree, 1):  d W_r = sum_{e: type r} x_src^T g_dst c_dst,  d att[r,b] = <dW_r,
    basis_b>,  d basis_b = sum_r att[r,b] dW_r,  d root = x^T g,  d bias = sum g,  d x_u = sum_{e out of u} g_dst c_dst
    W_type^T + g_u root^T"""
    torch.manual_seed(3)
    conv = pr.RGCNConvRef(6, 32, 5, 4).double()
    x, ei, et = _rand_graph(n=19, e=120, seed=4)
    x = x.clone().requires_grad_(True)
    g = torch.rand(19, 32, dtype=torch.float64, generator=torch.Generator().manual_seed(9)) - 0.5
    (conv(x, ei, et) * g).sum().backward()
    n, E = x.shape[0], ei.shape[1]
    W = torch.einsum("rb,bio->rio", conv.att, conv.basis).detach()
    deg = torch.zeros(n, dtype=torch.float64)
    for k in range(E):
        deg[int(ei[1, k])] += 1
    c = 1.0 / deg.clamp(min=1)
    dW = torch.zeros_like(W)
    dx = torch.zeros(n, 6, dtype=torch.float64)
    xd = x.detach()
    for k in range(E):
        u, v, r = int(ei[0, k]), int(ei[1, k]), int(et[k])
        dW[r] += torch.outer(xd[u], g[v] * c[v])
        dx[u] += (g[v] * c[v]) @ W[r].T
    dx += g @ conv.root.detach().T
    datt = torch.einsum("rio,bio->rb", dW, conv.basis.detach())
    dbasis = torch.einsum("rb,rio->bio", conv.att.detach(), dW)
    assert torch.allclose(conv.att.grad, datt, atol=1e-11) and torch.allclose(conv.basis.grad, dbasis, atol=1e-11)
    assert torch.allclose(conv.root.grad, xd.T @ g, atol=1e-11) and torch.allclose(conv.bias.grad, g.sum(0), atol=1e-11)
    assert torch.allclose(x.grad, dx, atol=1e-11)


def test_dropout_adj_naive_and_directed_independence():
    """A.2: training -> keep exactly the edges whose Bernoulli(1-p) draw is 1, order preserved, attributes carried
    along, each DIRECTION of a rating drawn independently; eval or p = 0 -> identity.  The model applies it ONCE per
    forward (models.py:193-198): every layer sees the same kept edges."""
    x, ei, et = _rand_graph(n=12, e=60, seed=6)
    keep = torch.rand(60, generator=torch.Generator().manual_seed(1)) > 0.3
    ei2, et2 = pr.dropout_adj(ei, et, 0.3, True, keep)
    want = [k for k in range(60) if bool(keep[k])]
    assert ei2.shape[1] == len(want) and torch.equal(ei2, ei[:, want]) and torch.equal(et2, et[want])
    a, b = pr.dropout_adj(ei, et, 0.3, False, keep)
    assert a is ei and b is et
    a, b = pr.dropout_adj(ei, et, 0.0, True, keep)
    assert a is ei and b is et
    # random draws: a symmetrised edge list loses its two directions independently
    sym = torch.cat([ei, ei.flip(0)], 1)
    g = torch.Generator().manual_seed(5)
    e3, _ = pr.dropout_adj(sym, torch.cat([et, et]), 0.5, True, None, g)
    kept = {(int(e3[0, k]), int(e3[1, k])) for k in range(e3.shape[1])}
    one_way = sum(1 for (u, v) in kept if (v, u) not in kept)
    assert one_way > 0
    # the model forward with an injected mask == the forward on the explicitly filtered graph without dropout
    torch.manual_seed(0)
    m = pr.IGMCRef(4, (32, 32, 32, 32), 5, 4, 0.3).double().train()
    xx = torch.zeros(12, 4, dtype=torch.float64)
    xx[0, 0] = 1.0
    xx[1, 1] = 1.0
    xx[2:, 3] = 1.0
    hk = torch.ones(1, 128, dtype=torch.bool)
    out1 = m(xx, ei, et, keep, hk)
    m.adj_dropout = 0.0
    out2 = m(xx, ei[:, want], et[want], None, hk)
    assert torch.allclose(out1, out2, atol=1e-13)
