"""Host-side cost of one pipelined step (perf_counter around the pieces of TrainEngine.step_pipe)."""
import os, sys, time
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
idx = [rng.choice(len(tu), B, replace=False) for _ in range(64)]
eng.prime(idx[0], epoch=1)
for s in range(12):
    eng.step_pipe(idx[s + 1], epoch=1)
torch.cuda.synchronize()
K = 300
for mode in ("full step_pipe", "replay only", "replay + event record", "fill + replay"):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K):
        if mode == "full step_pipe":
            eng.step_pipe(idx[k % 64], epoch=1)
        else:
            key = (B, B, eng.slot, B, True)
            if mode == "fill + replay":
                eng._fill(eng.hostbuf_np[eng.slot], idx[k % 64], 1, B)
            eng.graphs[key].replay()
            if mode == "replay + event record":
                eng.slot_ev[eng.slot].record()
            eng.slot ^= 1
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-24s host issue %.1f us/step   incl. drain %.1f us/step" % (mode, 1e6 * (t1 - t0) / K, 1e6 * (t2 - t0) / K))
# GPU-only rate of the same graphs (events)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(K):
    eng.graphs[(B, B, eng.slot, B, True)].replay(); eng.slot ^= 1
e1.record(); torch.cuda.synchronize()
print("device time per replay %.1f us" % (1000 * e0.elapsed_time(e1) / K))
