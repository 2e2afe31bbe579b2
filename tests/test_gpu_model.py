"""GPU parity of the fused RGCN forward / backward / train step against the restated PyG-1.4.2 oracle.

Tolerances: predicted ratings within 1e-4 RMSE of the oracle forward (north_star); gradients within
2e-4 relative (fp32 kernels vs fp64 oracle autograd)."""
import numpy as np
import pytest
import torch

from oracle import extract_np, pyg_restated
from igmc_b200.data import synth_ratings, build_adj

pytestmark = pytest.mark.gpu


def _oracle_batch(nu=120, nv=90, nnz=2500, R=5, B=16, mnph=30, seed=3):
    u, v, lab = synth_ratings(nu, nv, nnz, R, seed)
    A = build_adj(u, v, lab, nu, nv)
    g = extract_np.RatingCSR(A)
    cv = np.arange(1, R + 1, dtype=np.float32)
    ob = extract_np.extract_batch(g, u[:B], v[:B], lab[:B], cv, 1, 1.0, mnph, seed=5)
    return A, (u, v, lab), cv, ob


def _models(R=5, NB=4, adj_dropout=0.2, seed=0, multiply_by=1, plan="auto"):
    from igmc_b200.models import IGMC
    torch.manual_seed(seed)
    ref = pyg_restated.IGMCRef(4, (32, 32, 32, 32), R, NB, adj_dropout, multiply_by).double()
    m = IGMC(4, latent_dim=[32, 32, 32, 32], num_relations=R, num_bases=NB, regression=True,
             adj_dropout=adj_dropout, multiply_by=multiply_by).cuda()
    m.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    m.kernel_plan = plan   # 0: generic basis-space kernels; 1/2/4: relation-space kernels, CTAs per subgraph
    return ref, m


def _gpu_batch(ob):
    from igmc_b200.util_functions import Batch
    return Batch.from_arrays(ob["x"], ob["edge_index"], ob["edge_type"], ob["batch"], ob["y"], ob["num_graphs"])


def _rmse(a, b):
    return float(torch.sqrt(torch.mean((a.double().cpu() - b.double().cpu()) ** 2)))


@pytest.mark.parametrize("R,NB,mult,plan", [(5, 4, 1, 0), (5, 4, 1, 1), (5, 4, 1, 2), (5, 4, 1, 3), (5, 4, 1, 4), (10, 2, 1, 0),
                                            (10, 2, 1, 2), (5, 4, 2, "auto"), (30, 4, 1, "auto")])
def test_forward_eval_parity(R, NB, mult, plan):
    A, links, cv, ob = _oracle_batch(R=R)
    ref, m = _models(R, NB, multiply_by=mult, plan=plan)
    ref.eval(); m.eval()
    tb = pyg_restated.to_torch_batch(ob, torch.float64)
    with torch.no_grad():
        want, cs = ref(tb["x"], tb["edge_index"], tb["edge_type"], return_states=True)
        got = m(_gpu_batch(ob))
    assert _rmse(got, want) <= 1e-4, _rmse(got, want)
    ws = next(iter(m._ws.values()))
    assert _rmse(ws["states"][:cs.shape[0]], cs) <= 1e-5


def test_forward_on_extracted_batch_matches_foreign_batch():
    """same subgraphs through the CUDA extractor (symmetric adjacency) and through from_arrays"""
    from igmc_b200.util_functions import MyDynamicDataset
    A, (u, v, lab), cv, ob = _oracle_batch()
    ref, m = _models()
    m.eval()
    ds = MyDynamicDataset(None, A, (u, v), lab, 1, 1.0, 30, None, None, cv, seed=5)
    b = ds.extract_batch(np.arange(16))
    with torch.no_grad():
        p1 = m(b).clone()
        p2 = m(_gpu_batch(ob)).clone()
    assert torch.equal(p1, p2)


@pytest.mark.parametrize("R,NB,symmetric,plan", [(5, 4, True, 0), (5, 4, True, 1), (5, 4, True, 2), (5, 4, True, 3), (5, 4, True, 4),
                                                 (10, 2, True, 0), (10, 2, True, 2), (5, 4, False, 0),
                                                 (5, 4, False, 2), (30, 4, True, "auto")])
def test_train_forward_backward_parity(R, NB, symmetric, plan):
    """injected edge/hidden dropout draws; loss, predictions and every parameter gradient vs autograd"""
    A, links, cv, ob = _oracle_batch(R=R)
    if not symmetric:   # make the message graph structurally asymmetric (foreign batch path)
        rng = np.random.default_rng(1)
        keep = rng.random(ob["edge_index"].shape[1]) > 0.3
        # keep edges grouped by graph: boolean mask preserves order
        ob = dict(ob, edge_index=ob["edge_index"][:, keep], edge_type=ob["edge_type"][keep])
    ref, m = _models(R, NB, adj_dropout=0.2, plan=plan)
    ref.train(); m.train()
    E, B = ob["edge_index"].shape[1], ob["num_graphs"]
    gen = torch.Generator().manual_seed(11)
    ek = torch.rand(E, generator=gen) > 0.2
    hk = torch.rand(B, 128, generator=gen) > 0.5
    tb = pyg_restated.to_torch_batch(ob, torch.float64)
    ARR = 0.001
    loss_ref, out_ref = pyg_restated.train_loss(ref, tb, ARR, ek, hk)
    loss_ref.backward()
    b = _gpu_batch(ob)
    loss = m.fused_step(b, ARR=ARR, edge_keep=ek, hidden_keep=hk)
    b.check()
    ws = next(iter(v for k, v in m._ws.items() if k[2]))
    assert _rmse(ws["pred"], out_ref.detach()) <= 1e-4
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * max(1.0, abs(float(loss_ref)))
    sd_ref = dict(ref.named_parameters())
    worst = 0.0
    names = {id(p): name for name, p in m.named_parameters()}
    for (o, n, s), (_, _, p) in zip(m._layout, m._named_order()):
        name = names[id(p)]
        gref = sd_ref[name].grad.reshape(-1)
        ggpu = m.flat_grad[o:o + n].double().cpu()
        denom = float(gref.abs().max()) + 1e-12
        err = float((ggpu - gref).abs().max()) / denom
        worst = max(worst, err)
        assert err <= 2e-4, (name, err)
    # autograd path (IGMC.forward + loss.backward) gives the same gradients
    m.zero_grad()
    out = m(b, edge_keep=ek, hidden_keep=hk)
    l2 = torch.nn.functional.mse_loss(out, b.y)
    l2.backward()
    for name, p in m.named_parameters():
        gref = sd_ref[name].grad.float()
        if name.endswith("att") or name.endswith("basis"):
            continue  # reference grads include the ARR term; compared above through fused_step
        assert float((p.grad.cpu() - gref).abs().max()) <= 2e-4 * (float(gref.abs().max()) + 1e-12), name


def test_hash_dropout_matches_host_twin():
    """without injected draws the kernels use the counter hash; the host twin reproduces the edge mask"""
    from igmc_b200.models import edge_keep_reference, splitmix64
    A, links, cv, ob = _oracle_batch(B=6)
    ref, m = _models(adj_dropout=0.3)
    ref.train(); m.train()
    b = _gpu_batch(ob)
    E = ob["edge_index"].shape[1]
    gen = torch.Generator().manual_seed(3)
    hk = torch.rand(6, 128, generator=gen) > 0.5
    m._step = 41
    seed = splitmix64(m.drop_seed + 42)      # fused_step increments _step first
    ek = edge_keep_reference(seed, E, 0.3)
    assert 0.55 < float(ek.float().mean()) < 0.85
    tb = pyg_restated.to_torch_batch(ob, torch.float64)
    loss_ref, out_ref = pyg_restated.train_loss(ref, tb, 0.0, ek, hk)
    m.fused_step(b, ARR=0.0, hidden_keep=hk)
    ws = next(iter(v for k, v in m._ws.items() if k[2]))
    assert _rmse(ws["pred"], out_ref.detach()) <= 1e-4


def test_adam_matches_torch():
    from igmc_b200.models import FusedAdam
    A, links, cv, ob = _oracle_batch()
    ref, m = _models(adj_dropout=0.0)
    ref = ref.float()
    ref.train(); m.train()
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=0)
    opt = FusedAdam(m, lr=1e-3)
    tb = pyg_restated.to_torch_batch(ob, torch.float32)
    B = ob["num_graphs"]
    b = _gpu_batch(ob)
    for it in range(3):
        hk = torch.rand(B, 128, generator=torch.Generator().manual_seed(it)) > 0.5
        opt_ref.zero_grad()
        loss_ref, _ = pyg_restated.train_loss(ref, tb, 0.001, None, hk)
        loss_ref.backward()
        opt_ref.step()
        m.fused_step(b, ARR=0.001, hidden_keep=hk)
        opt.step()
    for name, p in m.named_parameters():
        want = dict(ref.named_parameters())[name]
        assert float((p.detach().cpu() - want.detach()).abs().max()) <= 5e-5, name
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) >= {"step", "exp_avg", "exp_avg_sq"} and int(sd["state"][0]["step"]) == 3


def test_determinism_and_graph_replay():
    """two runs of the same steps give bit-identical parameters; CUDA-graph replay == eager launches"""
    from igmc_b200.data import make_synthetic_dataset
    from igmc_b200.models import IGMC, FusedAdam
    from igmc_b200.train_eval import TrainEngine
    from igmc_b200.util_functions import MyDynamicDataset
    ds = make_synthetic_dataset("tiny", seed=0)
    tu, tv, tl = ds["train"]
    outs = []
    for use_graph in (False, True, True):
        d = MyDynamicDataset(None, ds["adj_train"], (tu, tv), tl, 1, 1.0, 10, None, None, ds["class_values"])
        torch.manual_seed(0)
        m = IGMC(d, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.2).cuda()
        opt = FusedAdam(m, lr=1e-3)
        eng = TrainEngine(d, m, opt, 8, ARR=0.001, use_graph=use_graph)
        for s in range(6):
            eng.step(np.arange(s * 8, s * 8 + 8), epoch=1)
        eng.check()
        outs.append(m.flat_params.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    assert torch.isfinite(outs[0]).all()
    # pipelined engine (extraction of batch k+1 on a side stream / second graph branch): same parameters
    for use_graph in (False, True):
        d = MyDynamicDataset(None, ds["adj_train"], (tu, tv), tl, 1, 1.0, 10, None, None, ds["class_values"])
        torch.manual_seed(0)
        m = IGMC(d, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.2).cuda()
        opt = FusedAdam(m, lr=1e-3)
        eng = TrainEngine(d, m, opt, 8, ARR=0.001, use_graph=use_graph)
        eng.prime(np.arange(0, 8), epoch=1)
        for s in range(6):
            eng.step_pipe(np.arange((s + 1) * 8, (s + 1) * 8 + 8) if s < 5 else None, epoch=1)
        eng.check()
        torch.cuda.synchronize()
        assert torch.equal(m.flat_params, outs[0]), "pipelined engine diverged (graph=%s)" % use_graph


@pytest.mark.parametrize("name,mnph,plan", [("ml_100k", 200, 2), ("ml_100k", 200, 4), ("ml_1m", 100, 1), ("ml_1m", 100, 2)])
def test_full_size_train_step_parity(name, mnph, plan):
    """BASELINE-size subgraphs (up to 402 nodes: chunked staging rows, long segmented lists, hash edge dropout):
    extracted on the GPU, checked against the oracle on the same subgraphs — predictions, loss and gradients."""
    from igmc_b200.data import make_synthetic_dataset
    from igmc_b200.models import IGMC, edge_keep_reference, splitmix64
    from igmc_b200.util_functions import MyDynamicDataset
    ds = make_synthetic_dataset(name, seed=0)
    tu, tv, tl = ds["train"]
    d = MyDynamicDataset(None, ds["adj_train"], (tu, tv), tl, 1, 1.0, mnph, None, None, ds["class_values"], seed=3)
    idx = np.array([0, 1, 2, 3, 4, 5])
    b = d.extract_batch(idx)
    g = extract_np.RatingCSR(ds["adj_train"])
    ob = extract_np.extract_batch(g, tu[idx], tv[idx], tl[idx], ds["class_values"], 1, 1.0, mnph, seed=3, pair_ids=idx)
    assert np.array_equal(b.edge_index.cpu().numpy(), ob["edge_index"])
    torch.manual_seed(0)
    ref = pyg_restated.IGMCRef(4, (32, 32, 32, 32), 5, 4, 0.2).double().train()
    m = IGMC(d, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.2).cuda().train()
    m.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    m.kernel_plan = plan
    E = ob["edge_index"].shape[1]
    m._step = 10
    ek = edge_keep_reference(splitmix64(m.drop_seed + 11), E, 0.2)     # the kernels' own hash draw
    hk = torch.rand(6, 128, generator=torch.Generator().manual_seed(2)) > 0.5
    tb = pyg_restated.to_torch_batch(ob, torch.float64)
    loss_ref, out_ref = pyg_restated.train_loss(ref, tb, 0.001, ek, hk)
    loss_ref.backward()
    loss = m.fused_step(b, ARR=0.001, hidden_keep=hk)
    b.check()
    ws = next(iter(v for k, v in m._ws.items() if k[2]))
    assert _rmse(ws["pred"], out_ref.detach()) <= 1e-4
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * max(1.0, abs(float(loss_ref)))
    names = {id(p): n for n, p in m.named_parameters()}
    sd_ref = dict(ref.named_parameters())
    for (o, n, s), (_, _, p) in zip(m._layout, m._named_order()):
        gref = sd_ref[names[id(p)]].grad.reshape(-1)
        err = float((m.flat_grad[o:o + n].double().cpu() - gref).abs().max()) / (float(gref.abs().max()) + 1e-12)
        assert err <= 2e-4, (names[id(p)], err)


@pytest.mark.parametrize("plan,adj_dropout,symmetric", [(1, 0.0, True), (2, 0.0, True), (2, 0.2, True), (3, 0.2, True),
                                                        (4, 0.2, True), (2, 0.2, False), (1, 0.3, False)])
def test_staged_list_images_equal_self_staged(plan, adj_dropout, symmetric):
    """igmc_stage_lists + the kernels' bulk (TMA) loads of the list images give BIT-identical predictions, loss and
    gradients to the kernels staging their own lists (same hash dropout draws), for extracted (symmetric) batches
    and for foreign batches with separate out-lists."""
    from igmc_b200.util_functions import MyDynamicDataset
    A, (u, v, lab), cv, ob = _oracle_batch(B=12)
    ref, m = _models(adj_dropout=adj_dropout, plan=plan)
    m.train()
    if symmetric:
        ds = MyDynamicDataset(None, A, (u, v), lab, 1, 1.0, 30, None, None, cv, seed=5)
        b = ds.extract_batch(np.arange(12))
    else:
        keep = np.random.default_rng(1).random(ob["edge_index"].shape[1]) > 0.3
        b = _gpu_batch(dict(ob, edge_index=ob["edge_index"][:, keep], edge_type=ob["edge_type"][keep]))
    hk = torch.rand(12, 128, generator=torch.Generator().manual_seed(2)) > 0.5
    m._step = 20
    loss0 = float(m.fused_step(b, ARR=0.001, hidden_keep=hk))
    b.check()
    ws = next(iter(v for k, v in m._ws.items() if k[2]))
    g0, p0 = m.flat_grad.clone(), ws["pred"].clone()
    m._step = 21                                   # fused_step draws with the seed of step 21
    st = m.stage_batch(b, True, m.make_dropout(True))
    assert st is not None and int(st["fwd"][1]["tab"][:, 3].sum()) == int(b._priv["node_ptr"][-1])   # own nodes add up
    m._step = 20
    m.flat_grad.zero_()
    loss1 = float(m.fused_step(b, ARR=0.001, hidden_keep=hk))
    b.check()
    assert torch.equal(ws["pred"], p0) and loss1 == loss0
    assert torch.equal(m.flat_grad, g0)
    assert float(g0.abs().max()) > 0
    b._stage = None


@pytest.mark.parametrize("plan,staged", [(1, False), (2, False), (2, True), (3, True), (4, True), (4, False)])
def test_one_launch_forward_backward_equals_two_launches(plan, staged, monkeypatch):
    """igmc_forward_backward (a cluster runs its subgraph's forward, loss and backward in ONE launch) against
    igmc_forward + igmc_backward on the same batch and dropout draws: predictions, raw gradient rows, readout
    factors and the assembled gradient are BIT-identical."""
    from igmc_b200.util_functions import MyDynamicDataset
    A, (u, v, lab), cv, ob = _oracle_batch(B=12)
    ref, m = _models(adj_dropout=0.2, plan=plan)
    m.train()
    ds = MyDynamicDataset(None, A, (u, v), lab, 1, 1.0, 30, None, None, cv, seed=5)
    b = ds.extract_batch(np.arange(12))
    hk = torch.rand(12, 128, generator=torch.Generator().manual_seed(2)) > 0.5
    got = []
    for fused in ("0", "1", "1"):
        monkeypatch.setenv("IGMC_FUSED_FB", fused)
        m._step = 30
        if staged:
            m._step = 31
            assert m.stage_batch(b, True, m.make_dropout(True)) is not None
            m._step = 30
        m.flat_grad.zero_()
        saved = m.forward_backward(b, hidden_keep=hk)
        ws = saved["ws"]
        ws["gpart_copy"] = ws["gpart"].clone()
        m._launch_grad_reduce(b, saved, loss_scale=1.0 / 12, arr=0.001)
        b.check()
        got.append((ws["pred"].clone(), ws["gpart_copy"], ws["dhid"].clone(), ws["dpred"].clone(),
                    m.flat_grad.clone(), float(ws["loss"])))
        b._stage = None
    for a, c in ((got[0], got[1]), (got[1], got[2])):
        for x, y in zip(a[:5], c[:5]):
            assert torch.equal(x, y)
        assert a[5] == c[5]
    assert float(got[0][4].abs().max()) > 0
