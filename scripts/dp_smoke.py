"""2-rank NCCL smoke of the pipelined engine with progress logging (debug aid; short, bounded)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.makedirs("gpurun_out", exist_ok=True)
logf = open("gpurun_out/dp_rank%d.log" % rank, "w")
t0 = time.time()
def log(*a):
    logf.write("[%6.2f] " % (time.time() - t0) + " ".join(str(x) for x in a) + "\n"); logf.flush()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
log("pg up")
from igmc_b200.data import make_synthetic_dataset
from igmc_b200.models import IGMC, FusedAdam
from igmc_b200.train_eval import TrainEngine
from igmc_b200.util_functions import MyDynamicDataset
ds = make_synthetic_dataset("tiny", seed=0)
tu, tv, tl = ds["train"]
d = MyDynamicDataset(None, ds["adj_train"], (tu, tv), tl, 1, 1.0, 10, None, None, ds["class_values"])
torch.manual_seed(1)
m = IGMC(d, latent_dim=[32] * 4, num_relations=5, num_bases=4, regression=True, adj_dropout=0.0).cuda()
dist.broadcast(m.flat_params, 0)
log("broadcast done")
use_graph = os.environ.get("DP_GRAPH", "1") == "1"
opt = FusedAdam(m, lr=1e-3)
eng = TrainEngine(d, m, opt, 8, ARR=0.001, use_graph=use_graph)
G = 16
idx = lambda s: np.arange(s * G + rank * 8, s * G + rank * 8 + 8)
eng.prime(idx(0), epoch=1, G=G)
for s in range(12):
    eng.step_pipe(idx(s + 1), epoch=1, next_G=G)
    torch.cuda.synchronize()
    log("step", s, "done loss", float(eng.last_loss.item()), "graphs", len(eng.graphs))
chk = m.flat_params.clone()
dist.all_reduce(chk)
log("param checksum equal across ranks:", bool(torch.allclose(chk, m.flat_params * world)))
dist.barrier()
log("barrier passed")
import threading
def bail():
    log("destroy_process_group did not return in 15 s -> os._exit")
    os._exit(0)
tm = threading.Timer(15.0, bail); tm.daemon = True; tm.start()
eng.graphs.clear()
torch.cuda.synchronize()
log("graphs released")
dist.destroy_process_group()
log("clean exit")
os._exit(0)
