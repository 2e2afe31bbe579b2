"""Where does a graph-replayed step go?  Times (CUDA events, L2 flushed between replays) three captured graphs on the
headline workload: the model branch alone, the extraction branch alone, and the full two-branch step."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from igmc_b200.data import make_synthetic_dataset
from igmc_b200.models import IGMC, FusedAdam
from igmc_b200.train_eval import TrainEngine
from igmc_b200.util_functions import MyDynamicDataset

wl = sys.argv[1] if len(sys.argv) > 1 else "ml_1m"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ds = make_synthetic_dataset(wl, seed=0)
tu, tv, tl = ds["train"]
d = MyDynamicDataset(None, ds["adj_train"], (tu, tv), tl, 1, 1.0, ds["max_nodes_per_hop"], None, None, ds["class_values"])
torch.manual_seed(1)
m = IGMC(d, latent_dim=[32] * 4, num_relations=ds["num_relations"], num_bases=4, regression=True,
         adj_dropout=ds["adj_dropout"]).cuda()
opt = FusedAdam(m, lr=1e-3)
eng = TrainEngine(d, m, opt, B, ARR=0.001)
rng = np.random.default_rng(0)
idx = [rng.choice(len(tu), B, replace=False) for _ in range(64)]
eng.prime(idx[0], epoch=1)
for s in range(8):
    eng.step_pipe(idx[s + 1], epoch=1)
torch.cuda.synchronize()
buf, side = eng.stepbuf_dev, eng.side
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def model_only():
    eng._model_step(eng.batches[0], B, B, buf[B + 1:B + 2])


def extract_only():
    b = d.extractor.extract(idx=buf[:B], seed_dev=buf[B:B + 1], reuse=True, slot=1)
    m.stage_batch(b, True, m.make_dropout(True, seed_dev=buf[B + 3:B + 4]), slot=1)


def both():
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        extract_only()
    model_only()
    main.wait_stream(side)


lo_side = torch.cuda.Stream(priority=0)
hi_main = torch.cuda.Stream(priority=-1)


def variant(delay_after, prio):
    """delay_after: None | 'prep' | 'fwd'  (the extraction branch starts after that model kernel has finished);
    prio: run the model branch on a high-priority stream"""
    def fn():
        outer = torch.cuda.current_stream()
        ms = hi_main if prio else outer
        ss = lo_side if prio else side
        if prio:
            ms.wait_stream(outer)
        with torch.cuda.stream(ms):
            b = eng.batches[0]
            m._step += 1
            drop = m.make_dropout(True, seed_dev=buf[B + 1:B + 2])
            if delay_after is None:
                ss.wait_stream(ms)
            import igmc_b200._lib as L, ctypes as C
            from igmc_b200.util_functions import _stream_ptr
            ws = m._workspace(b, True)
            L.check(L.load().igmc_prep_weights(C.byref(m._cmodel), m.flat_params.data_ptr(), m._wprep_buf().data_ptr(),
                                               _stream_ptr()), "prep")
            if delay_after == "prep":
                ss.wait_stream(ms)
            _, saved = m._launch_forward(b, True, drop, y=b.y, loss_scale=1.0 / B)
            if delay_after == "fwd":
                ss.wait_stream(ms)
            with torch.cuda.stream(ss):
                extract_only()
            m._launch_backward(b, drop, saved, saved["ws"]["dpred"])
            opt.reduce_update(eng.exchange, saved["ws"], B, B * saved["ws"]["cluster"], 1.0 / B, 0.001,
                              lr_dev=eng.lr_dev, loss_acc=eng.loss_acc, loss_weight=float(B))
            ms.wait_stream(ss)
        if prio:
            outer.wait_stream(ms)
    return fn


def parts():
    out = {}
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
    acc = np.zeros(7)
    for it in range(20):
        flush.fill_(1)
        b = eng.batches[0]
        evs[0].record()
        m._step += 1
        drop = m.make_dropout(True, seed_dev=buf[B + 1:B + 2])
        _, saved = m._launch_forward(b, True, drop, y=b.y, loss_scale=1.0 / B)
        evs[1].record()
        m._launch_backward(b, drop, saved, saved["ws"]["dpred"])
        evs[2].record()
        m._launch_grad_reduce(b, saved, 1.0 / B, 0.001, True)
        evs[3].record()
        opt.step(lr_dev=eng.lr_dev)
        evs[4].record()
        bb = d.extractor.extract(idx=buf[:B], seed_dev=buf[B:B + 1], reuse=True, slot=1)
        evs[5].record()
        m.stage_batch(bb, True, m.make_dropout(True, seed_dev=buf[B + 3:B + 4]), slot=1)
        evs[6].record()
        torch.cuda.synchronize()
        for i in range(6):
            acc[i] += evs[i].elapsed_time(evs[i + 1])
    names = ["prep+forward", "backward", "grad_reduce", "adam", "extract(2 kernels)", "stage_lists"]
    return {n: round(1000 * acc[i] / 20, 1) for i, n in enumerate(names)}


def timed(fn, reps=40, flushed=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    for k in range(reps):
        if flushed:
            flush.fill_(k & 0xff)
        e0[k].record()
        g.replay()
        e1[k].record()
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in zip(e0, e1)]) * 1000
    return round(float(np.median(t)), 1), round(float(t.min()), 1)


print("workload", wl, "B", B, "plan", m._plan(eng.batches[0]))
print("eager per-launch (us, L2 flushed):", parts())
cases = [("model branch", model_only), ("extraction branch", extract_only), ("both branches", both)]
for d_ in (None, "prep", "fwd"):
    for pr in (False, True):
        cases.append(("both, extraction after %s%s" % (d_ or "start", ", model high-priority" if pr else ""), variant(d_, pr)))
for name, fn in cases:
    print("%-58s graph replay us  flushed median/min %s   warm median/min %s" % (name, timed(fn), timed(fn, flushed=False)))
