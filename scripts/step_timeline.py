"""Wall-clock timeline of one graph-replayed step: when and on which SM every CTA of the forward, backward and
one-launch extraction kernels ran (globaltimer stamps written through the kernels' debug hooks)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from igmc_b200.data import make_synthetic_dataset
from igmc_b200.models import IGMC, FusedAdam
from igmc_b200.train_eval import TrainEngine
from igmc_b200.util_functions import MyDynamicDataset

B = 50
ds = make_synthetic_dataset("ml_1m", seed=0)
tu, tv, tl = ds["train"]
d = MyDynamicDataset(None, ds["adj_train"], (tu, tv), tl, 1, 1.0, 100, None, None, ds["class_values"])
torch.manual_seed(1)
m = IGMC(d, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.0).cuda()
opt = FusedAdam(m, lr=1e-3)
eng = TrainEngine(d, m, opt, B, ARR=0.001)
rng = np.random.default_rng(0)
idx = [rng.choice(len(tu), B, replace=False) for _ in range(40)]
# separate stamp rows for the forward (rows 0..99) and backward (rows 100..199): the kernels index by blockIdx
prof = torch.zeros(2 * 100 * 64, dtype=torch.int64, device="cuda")
m._prof_buf = prof          # forward and backward share the pointer: the backward overwrites -> read between them is
                            # impossible inside a graph, so run two replays and keep whichever kernel wrote last
eng.prime(idx[0], epoch=1)
for s in range(12):
    eng.step_pipe(idx[s + 1], epoch=1)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rep in range(3):
    flush.fill_(rep)
    prof.zero_()
    eng.step_pipe(idx[20 + rep], epoch=1)
    torch.cuda.synchronize()
    p = prof.view(-1, 64)[:100].cpu().numpy()
    ho = d.extractor.ws["hop_off"].view(-1, 8)[:B].cpu().numpy()
    # the backward wrote last into the shared rows: its start/end; the forward's are lost in this replay
    bs, be, bsm = p[:, 50] & 0x7fffffff, p[:, 51] & 0x7fffffff, p[:, 52]
    es, ee, esm = ho[:, 5].astype(np.int64), ho[:, 6].astype(np.int64), ho[:, 7]
    t0 = min(bs.min(), es.min())
    print("replay %d  (us relative to the earliest stamp)" % rep)
    print("  backward   CTA start min/median/max %.1f %.1f %.1f   end max %.1f   SMs used %d" % (
        (bs.min() - t0) / 1e3, (np.median(bs) - t0) / 1e3, (bs.max() - t0) / 1e3, (be.max() - t0) / 1e3, len(set(bsm.tolist()))))
    print("  extraction CTA start min/median/max %.1f %.1f %.1f   end min/median/max %.1f %.1f %.1f   SMs used %d   "
          "per-CTA duration median %.1f max %.1f" % (
              (es.min() - t0) / 1e3, (np.median(es) - t0) / 1e3, (es.max() - t0) / 1e3, (ee.min() - t0) / 1e3,
              (np.median(ee) - t0) / 1e3, (ee.max() - t0) / 1e3, len(set(esm.tolist())), np.median(ee - es) / 1e3,
              (ee - es).max() / 1e3))
    late = np.sort(bs - bs.min())[-6:] / 1e3
    print("  backward latest CTA starts (us after the first):", np.round(late, 1), " shared SMs:", len(set(bsm.tolist()) & set(esm.tolist())))
