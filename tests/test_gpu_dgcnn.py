"""GPU parity of DGCNN_RS (R-GCN stack -> SortPooling -> 1-D convolutions -> dense head; reference models.py:123-167)
against the restated oracle (oracle/pyg_restated.DGCNN_RSRef).

Tolerances: predictions within 1e-4 RMSE (north_star); gradients within 2e-4 relative of fp64 autograd.
The sort order is compared exactly (ties by node index on both sides)."""
import numpy as np
import pytest
import torch

from oracle import pyg_restated
from tests.test_gpu_model import _gpu_batch, _oracle_batch, _rmse

pytestmark = pytest.mark.gpu


def _models(R=5, NB=4, k=30, adj_dropout=0.2, seed=0, plan="auto", last=1):
    from igmc_b200.models import DGCNN_RS
    torch.manual_seed(seed)
    ref = pyg_restated.DGCNN_RSRef(4, (32, 32, 32, last), k, R, NB, adj_dropout).double()
    m = DGCNN_RS(4, latent_dim=[32, 32, 32, last], k=k, num_relations=R, num_bases=NB, regression=True,
                 adj_dropout=adj_dropout).cuda()
    m.load_state_dict({k_: v.float() for k_, v in ref.state_dict().items()})
    m.kernel_plan = plan
    return ref, m


@pytest.mark.parametrize("R,NB,k,plan,last", [(5, 4, 30, 0, 1), (5, 4, 30, 2, 1), (5, 4, 30, 4, 1), (5, 4, 80, 1, 1),
                                              (10, 2, 25, 2, 1), (5, 4, 31, "auto", 8)])
def test_forward_eval_parity(R, NB, k, plan, last):
    """k = 80 exceeds every subgraph of the batch (zero padding rows); k = 31 is odd (MaxPool drops the last row)"""
    A, links, cv, ob = _oracle_batch(R=R)
    ref, m = _models(R, NB, k, plan=plan, last=last)
    ref.eval(); m.eval()
    tb = pyg_restated.to_torch_batch(ob, torch.float64)
    with torch.no_grad():
        want, cs = ref(tb["x"], tb["edge_index"], tb["edge_type"], tb["batch"], num_graphs=tb["num_graphs"],
                       return_states=True)
        got = m(_gpu_batch(ob))
    ws = next(iter(m._ws.values()))
    W = ref.total_latent_dim
    assert _rmse(ws["states"][:cs.shape[0], :W], cs) <= 1e-5
    assert float(ws["states"][:cs.shape[0], W:].abs().max()) == 0.0     # padded channels of the narrow layer
    # the pooled order: same nodes at the same positions wherever the oracle's keys are separated
    perm = ws["sp"]["perm"].cpu().numpy()
    key = cs[:, -1].numpy()
    node_ptr = np.concatenate([[0], np.cumsum(np.bincount(ob["batch"], minlength=ob["num_graphs"]))])
    for g in range(ob["num_graphs"]):
        lo, hi = node_ptr[g], node_ptr[g + 1]
        order = lo + np.argsort(-key[lo:hi], kind="stable")
        kk = min(k, hi - lo)
        sep = np.abs(np.diff(key[order])) > 1e-6
        same = perm[g, :kk] == order[:kk]
        ok = same | ~np.concatenate([[True], sep])[:kk] | ~np.concatenate([sep, [True]])[:kk]
        assert ok.all(), g
        assert (perm[g, kk:] == -1).all()
    assert _rmse(got, want) <= 1e-4, _rmse(got, want)


@pytest.mark.parametrize("R,NB,k,plan", [(5, 4, 30, 0), (5, 4, 30, 1), (5, 4, 30, 2), (5, 4, 80, 2), (10, 2, 25, 2)])
def test_train_forward_backward_parity(R, NB, k, plan):
    A, links, cv, ob = _oracle_batch(R=R)
    ref, m = _models(R, NB, k, adj_dropout=0.2, plan=plan)
    ref.train(); m.train()
    E, B = ob["edge_index"].shape[1], ob["num_graphs"]
    gen = torch.Generator().manual_seed(11)
    ek = torch.rand(E, generator=gen) > 0.2
    hk = torch.rand(B, 128, generator=gen) > 0.5
    tb = pyg_restated.to_torch_batch(ob, torch.float64)
    ARR = 0.001
    out_ref = ref(tb["x"], tb["edge_index"], tb["edge_type"], tb["batch"], ek, hk, tb["num_graphs"])
    loss_ref = torch.nn.functional.mse_loss(out_ref, tb["y"].view(-1)) + ARR * pyg_restated.arr_regulariser(ref)
    loss_ref.backward()
    b = _gpu_batch(ob)
    loss = m.fused_step(b, ARR=ARR, edge_keep=ek, hidden_keep=hk)
    b.check()
    ws = next(iter(v for k_, v in m._ws.items() if k_[2]))
    assert _rmse(ws["sp"]["pred"], out_ref.detach()) <= 1e-4
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * max(1.0, abs(float(loss_ref)))
    sd_ref = dict(ref.named_parameters())
    names = {id(p): name for name, p in m.named_parameters()}
    for e, (_, _, p) in zip(m._layout, m._named_order()):
        name = names[id(p)]
        gref = sd_ref[name].grad
        ggpu = m._pview(m.flat_grad, e).double().cpu()
        err = float((ggpu - gref).abs().max()) / (float(gref.abs().max()) + 1e-12)
        assert err <= 2e-4, (name, err)
        if len(e) == 4:   # padded slot: the unused columns have exactly zero gradient
            assert float(m.flat_grad[e[0]:e[0] + e[1]].view(e[3])[..., e[2][-1]:].abs().max()) == 0.0
    # autograd path
    m.zero_grad()
    out = m(b, edge_keep=ek, hidden_keep=hk)
    torch.nn.functional.mse_loss(out, b.y).backward()
    for name, p in m.named_parameters():
        if name.endswith("att") or name.endswith("basis"):
            continue   # reference grads include the ARR term; compared above
        gref = sd_ref[name].grad.float()
        assert float((p.grad.cpu() - gref).abs().max()) <= 2e-4 * (float(gref.abs().max()) + 1e-12), name


def test_training_reduces_loss_and_keeps_padding_zero():
    """a few fused Adam steps through train_multiple_epochs' engine: loss goes down, padded weights stay zero"""
    from igmc_b200.models import FusedAdam
    A, links, cv, ob = _oracle_batch()
    ref, m = _models(k=30, adj_dropout=0.0)
    m.train()
    opt = FusedAdam(m, lr=1e-3)
    b = _gpu_batch(ob)
    losses = []
    for _ in range(30):
        losses.append(m.fused_step(b, ARR=0.001).clone())
        opt.step()
    losses = [float(x) for x in losses]
    assert losses[-1] < 0.7 * losses[0], losses[::5]
    for e in m._layout:
        if len(e) == 4:
            assert float(m.flat_params[e[0]:e[0] + e[1]].view(e[3])[..., e[2][-1]:].abs().max()) == 0.0


def test_percentile_k_and_engine():
    """k < 1 (models.py:69-73) from a dataset; one epoch through TrainEngine (CUDA graph, pipelined extraction)"""
    from igmc_b200.data import make_synthetic_dataset
    from igmc_b200.models import DGCNN_RS
    from igmc_b200.train_eval import train_multiple_epochs
    from igmc_b200.util_functions import MyDynamicDataset
    ds = make_synthetic_dataset("tiny", seed=0)
    tu, tv, tl = ds["train"]
    eu, ev, el = ds["test"]
    train = MyDynamicDataset(None, ds["adj_train"], (tu, tv), tl, 1, 1.0, 10, None, None, ds["class_values"])
    test = MyDynamicDataset(None, ds["adj_train"], (eu[:200], ev[:200]), el[:200], 1, 1.0, 10, None, None,
                            ds["class_values"])
    torch.manual_seed(1)
    m = DGCNN_RS(train, latent_dim=[32, 32, 32, 1], k=0.6, num_relations=5, num_bases=4, regression=True,
                 adj_dropout=0.0)
    nn_ = sorted(int(train[i].num_nodes) for i in range(len(train)))
    assert m.k == max(10, nn_[int(np.ceil(0.6 * len(nn_))) - 1])
    log = []
    rmse = train_multiple_epochs(train, test, m, epochs=3, batch_size=50, lr=1e-3, lr_decay_factor=0.1,
                                 lr_decay_step_size=50, weight_decay=0, ARR=0.001,
                                 logger=lambda info, mm, opt: log.append(dict(info)))
    assert np.isfinite(rmse) and len(log) == 3
    assert log[-1]["train_loss"] < log[0]["train_loss"]
