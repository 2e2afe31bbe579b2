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
M31 = 0x7fffffff
for mode in ("flushed", "back-to-back"):
    for rep in range(4):
        if mode == "flushed":
            flush.fill_(rep)
        prof.zero_()
        for _ in range(1 if mode == "flushed" else 3):      # back-to-back: stamps of the LAST of three replays
            eng.step_pipe(idx[20 + rep], epoch=1)
        torch.cuda.synchronize()
        p = prof.view(-1, 64)[:100].cpu().numpy()
        ho = d.extractor.ws["hop_off"].view(-1, 8)[:B].cpu().numpy()
        fs, fe, fsm = p[:, 50] & M31, p[:, 51] & M31, p[:, 52]
        bs, be, bsm = p[:, 53] & M31, p[:, 54] & M31, p[:, 55]
        es, ee, esm = ho[:, 5].astype(np.int64), ho[:, 6].astype(np.int64), ho[:, 7]
        t0 = fs.min()
        u = lambda a: np.round((a - t0) / 1e3, 1)
        print("%s replay %d (us after the first forward CTA started)" % (mode, rep))
        print("  forward    start min/med/max %s %s %s  end med/max %s %s  SMs %d" % (u(fs.min()), u(np.median(fs)), u(fs.max()), u(np.median(fe)), u(fe.max()), len(set(fsm.tolist()))))
        print("  extraction start min/med/max %s %s %s  end med/max %s %s  SMs %d" % (u(es.min()), u(np.median(es)), u(es.max()), u(np.median(ee)), u(ee.max()), len(set(esm.tolist()))))
        print("  backward   start min/med/max %s %s %s  end med/max %s %s  SMs %d  (SMs shared with extraction %d)" % (u(bs.min()), u(np.median(bs)), u(bs.max()), u(np.median(be)), u(be.max()), len(set(bsm.tolist())), len(set(bsm.tolist()) & set(esm.tolist()))))
        late = np.sort(fs)[-4:]
        print("  latest forward CTA starts:", u(late), " latest backward CTA starts:", u(np.sort(bs)[-4:]))
