"""GPU end-to-end checks of the reference-facing API: train_multiple_epochs, eval, static dataset."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import extract_np, pyg_restated

pytestmark = pytest.mark.gpu


def _tiny():
    from igmc_b200.data import make_synthetic_dataset
    return make_synthetic_dataset("tiny", seed=0)


def test_train_multiple_epochs_runs_and_learns(tmp_path):
    from igmc_b200.models import IGMC
    from igmc_b200.train_eval import train_multiple_epochs
    from igmc_b200.util_functions import MyDynamicDataset
    ds = _tiny()
    tu, tv, tl = ds["train"]
    eu, ev, el = ds["test"]
    train = MyDynamicDataset(None, ds["adj_train"], (tu, tv), tl, 1, 1.0, 10, None, None, ds["class_values"])
    test = MyDynamicDataset(None, ds["adj_train"], (eu[:200], ev[:200]), el[:200], 1, 1.0, 10, None, None,
                            ds["class_values"])
    torch.manual_seed(1)
    model = IGMC(train, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.2)
    log = []

    def logger(info, m, opt):
        log.append(dict(info))
        if m is not None:
            torch.save(m.state_dict(), tmp_path / "model_checkpoint{}.pth".format(info["epoch"]))
            torch.save(opt.state_dict(), tmp_path / "optimizer_checkpoint{}.pth".format(info["epoch"]))

    rmse = train_multiple_epochs(train, test, model, epochs=4, batch_size=50, lr=1e-3, lr_decay_factor=0.1,
                                 lr_decay_step_size=3, weight_decay=0, ARR=0.001, logger=logger)
    assert len(log) == 4 and [l["epoch"] for l in log] == [1, 2, 3, 4]
    assert all(math.isfinite(l["train_loss"]) and math.isfinite(l["test_rmse"]) for l in log)
    assert log[-1]["train_loss"] < log[0]["train_loss"]          # it learns
    assert rmse == log[-1]["test_rmse"]
    sd = torch.load(tmp_path / "model_checkpoint4.pth")
    assert sorted(sd.keys()) == sorted(pyg_restated.IGMCRef().state_dict().keys())   # reference checkpoint names
    # resume path (train_eval.py:55-64)
    model2 = IGMC(train, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.2)
    r2 = train_multiple_epochs(train, test, model2, epochs=5, batch_size=50, lr=1e-3, lr_decay_factor=0.1,
                               lr_decay_step_size=3, weight_decay=0, ARR=0.001, continue_from=4,
                               res_dir=str(tmp_path))
    assert math.isfinite(r2)


def test_eval_rmse_matches_oracle():
    from igmc_b200.models import IGMC
    from igmc_b200.train_eval import eval_rmse
    from igmc_b200.util_functions import MyDynamicDataset
    ds = _tiny()
    eu, ev, el = ds["test"]
    n = 120
    test = MyDynamicDataset(None, ds["adj_train"], (eu[:n], ev[:n]), el[:n], 1, 1.0, 10, None, None,
                            ds["class_values"], seed=9)
    torch.manual_seed(3)
    ref = pyg_restated.IGMCRef(4, (32, 32, 32, 32), 5, 4, 0.2).eval()
    model = IGMC(test, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.2).cuda()
    model.load_state_dict(ref.state_dict())
    got = eval_rmse(model, test, None, batch_size=50)
    g = extract_np.RatingCSR(ds["adj_train"])
    idx = np.arange(n)
    ob = extract_np.extract_batch(g, eu[:n], ev[:n], el[:n], ds["class_values"], 1, 1.0, 10, seed=9, pair_ids=idx)
    tb = pyg_restated.to_torch_batch(ob)
    with torch.no_grad():
        pred = ref(tb["x"], tb["edge_index"], tb["edge_type"])
    want = math.sqrt(float(((pred - tb["y"]) ** 2).mean()))
    assert abs(got - want) <= 1e-4, (got, want)


def test_static_dataset_equals_dynamic():
    """MyDataset (extract once, device-resident (data, slices)) gives the same graphs / predictions as the
    dynamic path when no sampling is involved (config 3 style)."""
    from igmc_b200.models import IGMC
    from igmc_b200.util_functions import MyDataset, MyDynamicDataset
    ds = _tiny()
    tu, tv, tl = ds["train"]
    n = 150
    args = (None, ds["adj_train"], (tu[:n], tv[:n]), tl[:n], 1, 1.0, None, None, None, ds["class_values"])
    dyn = MyDynamicDataset(*args)
    sta = MyDataset(*args)
    assert len(sta) == n and sta.num_features == 4
    for k in (0, 7, n - 1):
        a, b = dyn[k], sta[k]
        assert torch.equal(a.x, b.x) and torch.equal(a.edge_index, b.edge_index)
        assert torch.equal(a.edge_type, b.edge_type) and torch.equal(a.y, b.y)
    torch.manual_seed(0)
    m = IGMC(dyn, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True).cuda().eval()
    idx = np.array([3, 50, 51, 100, 149, 0])
    with torch.no_grad():
        p1 = m(dyn.extract_batch(idx)).clone()
        p2 = m(sta.extract_batch(idx)).clone()
    assert float((p1 - p2).abs().max()) <= 1e-5


def test_static_dataset_trains_like_dynamic():
    """config-3 style: training from the device-resident store (batch assembly kernel, CUDA-graph replay) gives
    bit-identical parameters to training with on-the-fly extraction when no sampling is involved."""
    from igmc_b200.models import IGMC, FusedAdam
    from igmc_b200.train_eval import TrainEngine, train
    from igmc_b200.util_functions import MyDataset, MyDynamicDataset
    ds = _tiny()
    tu, tv, tl = ds["train"]
    n = 200
    args = (None, ds["adj_train"], (tu[:n], tv[:n]), tl[:n], 1, 1.0, None, None, None, ds["class_values"])
    outs = []
    for cls in (MyDynamicDataset, MyDataset):
        d = cls(*args)
        torch.manual_seed(0)
        m = IGMC(d, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.0).cuda()
        opt = FusedAdam(m, lr=1e-3)
        eng = TrainEngine(d, m, opt, 25, ARR=0.001)
        gen = torch.Generator().manual_seed(5)
        losses = [train(m, opt, d, None, regression=True, ARR=0.001, epoch=e, engine=eng, generator=gen) for e in (1, 2)]
        assert all(math.isfinite(l) for l in losses)
        outs.append(m.flat_params.clone())
    assert float((outs[0] - outs[1]).abs().max()) <= 1e-6


def test_ensemble_eval_matches_oracle(tmp_path):
    """test_once(..., ensemble=True, checkpoints=[...]) (reference train_eval.py:114-139,208-245): predictions of the
    checkpoints are averaged, then the RMSE is taken; compared with the same average through the oracle models"""
    from igmc_b200.models import IGMC
    from igmc_b200.train_eval import test_once
    from igmc_b200.util_functions import MyDynamicDataset
    ds = _tiny()
    eu, ev, el = ds["test"]
    n = 110
    test = MyDynamicDataset(None, ds["adj_train"], (eu[:n], ev[:n]), el[:n], 1, 1.0, 10, None, None,
                            ds["class_values"], seed=4)
    model = IGMC(test, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.2).cuda()
    g = extract_np.RatingCSR(ds["adj_train"])
    ob = extract_np.extract_batch(g, eu[:n], ev[:n], el[:n], ds["class_values"], 1, 1.0, 10, seed=4,
                                  pair_ids=np.arange(n))
    tb = pyg_restated.to_torch_batch(ob)
    ckpts, preds = [], []
    for k in range(3):
        torch.manual_seed(100 + k)
        ref = pyg_restated.IGMCRef(4, (32, 32, 32, 32), 5, 4, 0.2).eval()
        path = tmp_path / "model_checkpoint{}.pth".format(k)
        torch.save(ref.state_dict(), path)          # a reference-format checkpoint (same keys and shapes)
        ckpts.append(str(path))
        with torch.no_grad():
            preds.append(ref(tb["x"], tb["edge_index"], tb["edge_type"]))
    want = math.sqrt(float(((torch.stack(preds, 1).mean(1) - tb["y"]) ** 2).mean()))
    log = []
    got = test_once(test, model, 50, logger=lambda info, m, o: log.append(info), ensemble=True, checkpoints=ckpts)
    assert abs(got - want) <= 1e-4, (got, want)
    assert log and log[0]["epoch"] == "ensemble"


def test_continue_from_checkpoint(tmp_path):
    """train_multiple_epochs(..., continue_from=k, res_dir=...) (reference train_eval.py:56-64): model and optimizer
    state are restored from the reference-format checkpoints, training resumes at epoch k+1 and runs epochs-k epochs"""
    from igmc_b200.models import IGMC
    from igmc_b200.train_eval import train_multiple_epochs
    from igmc_b200.util_functions import MyDynamicDataset
    ds = _tiny()
    tu, tv, tl = ds["train"]
    eu, ev, el = ds["test"]
    train = MyDynamicDataset(None, ds["adj_train"], (tu, tv), tl, 1, 1.0, 10, None, None, ds["class_values"])
    test = MyDynamicDataset(None, ds["adj_train"], (eu[:100], ev[:100]), el[:100], 1, 1.0, 10, None, None,
                            ds["class_values"])
    steps_per_epoch = math.ceil(len(train) / 50)
    state = {}

    def logger(info, m, opt):
        state.setdefault("log", []).append(dict(info))
        state["steps"] = int(opt.step_count[0])
        torch.save(m.state_dict(), tmp_path / "model_checkpoint{}.pth".format(info["epoch"]))
        torch.save(opt.state_dict(), tmp_path / "optimizer_checkpoint{}.pth".format(info["epoch"]))

    torch.manual_seed(1)
    m1 = IGMC(train, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.2)
    train_multiple_epochs(train, test, m1, epochs=2, batch_size=50, lr=1e-3, lr_decay_factor=0.1,
                          lr_decay_step_size=50, weight_decay=0, ARR=0.001, logger=logger)
    assert state["steps"] == 2 * steps_per_epoch
    saved = {k: v.clone() for k, v in torch.load(tmp_path / "model_checkpoint2.pth").items()}
    state["log"] = []
    m2 = IGMC(train, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.2)
    train_multiple_epochs(train, test, m2, epochs=3, batch_size=50, lr=1e-3, lr_decay_factor=0.1,
                          lr_decay_step_size=50, weight_decay=0, ARR=0.001, logger=logger, continue_from=2,
                          res_dir=str(tmp_path))
    assert [l["epoch"] for l in state["log"]] == [3]
    assert state["steps"] == 3 * steps_per_epoch                     # Adam's step count went on from the checkpoint
    moved = max(float((m2.state_dict()[k].cpu() - saved[k].cpu()).abs().max()) for k in saved)
    assert 0.0 < moved < 0.5                                          # started from the checkpoint, not from a fresh init


def _engine(ds, B, fused, seed=0, adj_dropout=0.2, hidden_p=0.5, eps=1e-8):
    from igmc_b200.models import IGMC, FusedAdam
    from igmc_b200.train_eval import TrainEngine
    from igmc_b200.util_functions import MyDynamicDataset
    tu, tv, tl = ds["train"]
    d = MyDynamicDataset(None, ds["adj_train"], (tu, tv), tl, 1, 1.0, 10, None, None, ds["class_values"])
    torch.manual_seed(seed)
    m = IGMC(d, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=adj_dropout).cuda()
    m.hidden_dropout_p = hidden_p
    opt = FusedAdam(m, lr=1e-3, eps=eps)
    eng = TrainEngine(d, m, opt, B, ARR=0.001)
    eng.fused_update = fused
    return eng, m, opt


def test_fused_update_kernel_equals_separate_kernels():
    """igmc_reduce_update (gradient assembly + ARR -> exchange -> Adam in ONE launch, world 1) leaves bit-identical
    parameters, Adam moments, step count and epoch-loss accumulator to igmc_grad_reduce + igmc_adam_step"""
    ds = _tiny()
    res = []
    for fused in (False, True):
        eng, m, opt = _engine(ds, 8, fused)
        eng.prime(np.arange(0, 8), epoch=1)
        for s in range(7):
            eng.step_pipe(np.arange((s + 1) * 8, (s + 1) * 8 + 8) if s < 6 else None, epoch=1)
        eng.check()
        torch.cuda.synchronize()
        assert (eng.exchange is not None) == fused
        res.append((m.flat_params.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone(), int(opt.step_count[0]),
                    float(eng.loss_acc), float(eng.last_loss)))
        if eng.exchange is not None:
            eng.exchange.close()
    a, b = res
    assert a[3] == b[3] == 7
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert a[4] == b[4] and a[5] == b[5] and math.isfinite(a[4])


def _dp_rank(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # carries the IPC handles only
    ds = _tiny()
    B = 8
    eng, m, opt = _engine(ds, B, True, adj_dropout=0.0, hidden_p=0.0, eps=1.0)
    G = B * world
    sl = lambda s: np.arange(s * G + rank * B, s * G + rank * B + B)   # noqa: E731
    eng.prime(sl(0), epoch=1, G=G)
    for s in range(5):
        eng.step_pipe(sl(s + 1) if s < 4 else None, epoch=1, next_G=G)
    eng.check()
    torch.cuda.synchronize()
    q.put((rank, m.flat_params.cpu(), float(eng.loss_acc)))
    dist.barrier()
    eng.exchange.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp2_peer_exchange_equals_single_gpu_global_batch():
    """2 ranks x batch 8 through the NVLink peer exchange inside igmc_reduce_update == 1 GPU x batch 16 on the same
    pairs (dropout off so that both see the same function): parameters within 1e-6, both ranks bit-identical, and the
    summed epoch-loss accumulators agree.  Adam runs with eps = 1 here: with the default 1e-8 the first updates are
    lr * sign(g) and entries whose gradient is near zero amplify summation-order noise (8 + 8 vs 16 partial rows) far
    beyond what the exchange could be blamed for; with eps = 1 the update is ~linear in g, so the bound on the
    parameters is a bound on the all-reduced gradients."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert torch.equal(got[0][1], got[1][1])                 # ranks stay bit-identical
    ds = _tiny()
    eng, m, opt = _engine(ds, 16, True, adj_dropout=0.0, hidden_p=0.0, eps=1.0)
    eng.prime(np.arange(0, 16), epoch=1)
    for s in range(5):
        eng.step_pipe(np.arange((s + 1) * 16, (s + 1) * 16 + 16) if s < 4 else None, epoch=1)
    eng.check()
    torch.cuda.synchronize()
    diff = float((m.flat_params.cpu() - got[0][1]).abs().max())
    moved = float((m.flat_params.cpu() - _engine(ds, 16, True, adj_dropout=0.0, hidden_p=0.0)[1].flat_params.cpu()).abs().max())
    assert diff <= 1e-6 and moved > 1e-5, (diff, moved)
    assert abs(float(eng.loss_acc) - (got[0][2] + got[1][2])) <= 1e-3 * abs(float(eng.loss_acc))
    eng.exchange.close()


def test_static_dataset_cache_file_interchange(tmp_path):
    """MyDataset(root=...) keeps the reference's cache contract (util_functions.py:91-110): the first construction
    extracts and writes <root>/processed/data.pt as (data, slices); the second loads the FILE (no extraction) and
    serves bit-identical batches - message-passing lists, predictions and gradients included - and a file written the
    reference's way (plain Data + slices pickled by hand) loads too."""
    from igmc_b200 import pyg_cache
    from igmc_b200.models import IGMC
    from igmc_b200.util_functions import MyDataset
    ds = _tiny()
    tu, tv, tl = ds["train"]
    n = 120
    args = (ds["adj_train"], (tu[:n], tv[:n]), tl[:n], 1, 1.0, 10, None, None, ds["class_values"])
    root = str(tmp_path / "data" / "tiny" / "train")
    a = MyDataset(root, *args)
    assert a.loaded_from is None and os.path.isfile(os.path.join(root, "processed", "data.pt"))
    b = MyDataset(root, *args)
    assert b.loaded_from is not None
    idx = np.array([5, 0, 77, 119, 33, 34, 35])
    ba, bb = a.extract_batch(idx), b.extract_batch(idx)
    for k in ("x", "edge_index", "edge_type", "batch", "y"):
        assert torch.equal(getattr(ba, k), getattr(bb, k)), k
    E = int(ba._priv["counts"][1])
    N = int(ba._priv["counts"][0])
    for k in ("adj_in_ptr", "adj_in", "adj_eid"):
        lim = N + 1 if k == "adj_in_ptr" else E
        assert torch.equal(ba._adj[1][k][:lim], bb._adj[1][k][:lim]), k
    torch.manual_seed(0)
    m = IGMC(a, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.2).cuda().train()
    hk = torch.rand(len(idx), 128, generator=torch.Generator().manual_seed(1)) > 0.5
    m._step = 3
    la = float(m.fused_step(ba, ARR=0.001, hidden_keep=hk)); ga = m.flat_grad.clone()
    m._step = 3
    lb = float(m.fused_step(bb, ARR=0.001, hidden_keep=hk))
    assert la == lb and torch.equal(ga, m.flat_grad)
    # single graphs through the reference's indexing idiom
    assert torch.equal(a[7].edge_index, b[7].edge_index) and torch.equal(a[7].x, b[7].x)
    # a cache pickled "by the reference": same arrays, its own Data object
    x, ei, et, y, noff, eoff = [t.cpu() for t in a.store.arrays()]
    root2 = str(tmp_path / "data" / "tiny" / "ref_written")
    os.makedirs(os.path.join(root2, "processed"))
    Data = pyg_cache._data_class()
    torch.save((Data(x=x, edge_index=ei, y=y, edge_type=et),
                {"x": noff, "edge_index": eoff, "y": torch.arange(n + 1), "edge_type": eoff}),
               os.path.join(root2, "processed", "data.pt"))
    c = MyDataset(root2, *args)
    assert c.loaded_from is not None
    bc = c.extract_batch(idx)
    assert torch.equal(bc.edge_index, ba.edge_index) and torch.equal(bc._adj[1]["adj_in"][:E], ba._adj[1]["adj_in"][:E])
    # a stale cache (different number of pairs) is refused, not silently used
    with pytest.raises(ValueError):
        MyDataset(root, ds["adj_train"], (tu[:50], tv[:50]), tl[:50], 1, 1.0, 10, None, None, ds["class_values"])
