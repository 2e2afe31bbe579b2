"""In-kernel phase timeline (clock64 stamps) of the cluster-plan forward / backward kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from igmc_b200.data import make_synthetic_dataset
from igmc_b200.models import IGMC
from igmc_b200.util_functions import MyDynamicDataset

ds = make_synthetic_dataset("ml_1m", seed=0)
tu, tv, tl = ds["train"]
d = MyDynamicDataset(None, ds["adj_train"], (tu, tv), tl, 1, 1.0, 100, None, None, ds["class_values"])
torch.manual_seed(1)
m = IGMC(d, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.0).cuda()
m.train()
B = 50
rng = np.random.default_rng(0)
for plan in (2, 1):
    m.kernel_plan = plan
    m._plans.clear(); m._ws.clear()
    grid = B * plan
    m._prof_buf = torch.zeros(grid * 64, dtype=torch.int64, device="cuda")
    for it in range(3):
        b = d.extract_batch(rng.choice(len(tu), B, replace=False))
        m._step += 1
        drop = m.make_dropout(True)
        if os.environ.get("IGMC_NO_STAGE") != "1":
            m.stage_batch(b, True, drop)     # list images, as the pipelined engine does on its side stream
        m._prof_buf.zero_()
        _, saved = m._launch_forward(b, True, drop, y=b.y, loss_scale=1.0 / B)
        torch.cuda.synchronize()
        f = m._prof_buf.view(grid, 64).cpu().numpy().copy()
        m._prof_buf.zero_()
        m._launch_backward(b, drop, saved, saved["ws"]["dpred"])
        torch.cuda.synchronize()
        bw = m._prof_buf.view(grid, 64).cpu().numpy().copy()
    def show(name, a, labels):
        a = a.astype(np.float64)
        t0 = a[:, 0:1]
        rel = (a - t0)
        print("== %s plan=%d (cycles since kernel start, mean over CTAs that stamped; us at 1.965 GHz)" % (name, plan))
        prev = 0.0
        for i, lab in enumerate(labels):
            col = rel[:, i]
            ok = a[:, i] > 0
            if not ok.any():
                continue
            mean = col[ok].mean()
            print("  %-28s t=%9.0f cyc (%6.1f us)  +%8.0f   max=%9.0f" % (lab, mean, mean / 1965.0, mean - prev, col[ok].max()))
            prev = mean
    def show2(name, a, labels):
        a = a.astype(np.float64)
        print("== %s plan=%d (us since kernel start at 1.965 GHz, mean / max over CTAs)" % (name, plan))
        order = sorted(labels.items(), key=lambda kv: a[:, kv[0]][a[:, kv[0]] > 0].mean() if (a[:, kv[0]] > 0).any() else 1e30)
        prev = 0.0
        for i, lab in order:
            ok = a[:, i] > 0
            if not ok.any():
                continue
            col = (a[:, i] - a[:, 0])[ok]
            print("  %-34s t=%7.1f us  +%6.1f   max=%7.1f" % (lab, col.mean() / 1965.0, (col.mean() - prev) / 1965.0, col.max() / 1965.0))
            prev = col.mean()
    fl = {0: "start", 1: "init+lists"}
    for l in range(4):
        fl.update({2 + 6 * l: "L%d layer start" % l, 3 + 6 * l: "L%d gather done (warp0)" % l, 4 + 6 * l: "L%d gather synced" % l,
                   5 + 6 * l: "L%d mma done (warp0)" % l, 6 + 6 * l: "L%d mma synced" % l, 7 + 6 * l: "L%d cluster synced" % l})
    fl.update({39: "L1 w0 gather start", 49: "L1 w31 gather start", 40: "L1 w0 gather_segments done", 43: "L1 w31 gather_segments done", 46: "L1 w15 gather_segments done",
               41: "L1 w0 after sync", 44: "L1 w31 after sync", 42: "L1 w0 fold done", 45: "L1 w31 fold done"})
    fl[26] = "readout done"
    fl.update({32: "L0 gather_segments done (warp0)", 33: "L0 after sync", 27: "readout start", 28: "readout lin1 done (warp0)",
               29: "readout lin1 synced", 34: "L0 zsave copied (barrier window)"})
    show2("forward", f, fl)
    print("  staging facts (fwd): staged=%s entries=%s segs=%s lcap=%s chunk=%s n_own=%s" % tuple(
        sorted(set(f[:, c].tolist()))[:6] for c in (60, 61, 62, 63, 59, 58)))
    bl = {0: "start", 1: "lists+readout bwd"}
    for i, l in enumerate((3, 2, 1, 0)):
        sb = 2 + 8 * i
        bl.update({sb: "L%d dpre ready" % l, sb + 5: "L%d dgrad gather+fold done (warp0)" % l, sb + 6: "L%d dgrad gather synced" % l,
                   sb + 1: "L%d dgrad mma done" % l, sb + 7: "L%d wgrad done (warp0)" % l, sb + 2: "L%d layer-0 wgrad done" % l,
                   sb + 3: "L%d layer synced" % l, sb + 4: "L%d cluster synced" % l})
    show2("backward", bw, bl)
