"""Drop-in for the reference's ``train_eval.py`` (train loop / eval) on the fused kernels.

``train_multiple_epochs`` / ``test_once`` / ``train`` / ``eval_loss`` / ``eval_rmse`` /
``eval_rmse_ensemble`` keep the reference signatures (train_eval.py:23-36,114-119,149-245) and
semantics (MSE mean + ARR regulariser, Adam, LR x factor every ``lr_decay_step_size`` epochs,
``logger(eval_info, model, optimizer)`` once per epoch, ``continue_from`` checkpoints).

What changed underneath:
* no DataLoader workers: a step is  H2D(indices) -> extract -> adjacency -> fused forward+loss ->
  backward -> gradient assembly (+ARR) -> [NCCL all-reduce] -> fused Adam, captured ONCE as a CUDA graph
  per batch size and replayed; the host never synchronises inside an epoch;
* data parallel: every rank holds the rating CSR, draws the same epoch permutation and takes its
  slice of every global batch; the only collective is one all-reduce(SUM) of the flat gradient
  (49 k floats) per step, plus one scalar all-reduce per epoch for the reported loss / RMSE.
"""
import math
import os
import time

import numpy as np
import torch

from .models import FusedAdam, splitmix64
from .util_functions import MyDynamicDataset  # noqa: F401  (re-export, Main.py star-imports)

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


def _dist_info():
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def deal_balanced(chunk, cost, world):
    """Partition one global batch over ``world`` ranks so that every rank gets the same number of pairs (+-1) and a
    similar mix of expensive and cheap ones: pairs sorted by ``cost`` (descending, ties by position) are dealt in
    snake order 0..w-1, w-1..0, ...  A step ends when the slowest rank's largest subgraph is done, so what matters
    is that no rank collects the big ones.  Returns a list of index arrays (positions into ``chunk`` kept in their
    original relative order).  Pure host logic."""
    g = len(chunk)
    order = np.argsort(-np.asarray(cost, dtype=np.float64), kind="stable")
    owner = np.empty(g, dtype=np.int64)
    pos = np.arange(g)
    lap, k = pos // world, pos % world
    owner[order] = np.where(lap % 2 == 0, k, world - 1 - k)
    return [np.asarray(chunk)[owner == r] for r in range(world)]


def shard_batches(perm, batch_size, rank, world, cost=None):
    """Split an epoch permutation into global batches of ``batch_size*world`` and return, per step,
    (this rank's indices, global batch size).  The last global batch may be short (the reference's
    DataLoader keeps the partial batch, drop_last=False); it is split as evenly as possible and a rank
    may then receive an empty slice.  ``cost`` (optional, one value per dataset index): deal every global batch
    balanced by it (``deal_balanced``) instead of in contiguous slices - the global batches, and therefore the
    optimisation trajectory up to summation order, are the same.  Pure host logic (covered by CPU tests)."""
    n = len(perm)
    gb = batch_size * world
    out = []
    for s in range(0, n, gb):
        chunk = perm[s:s + gb]
        g = len(chunk)
        if cost is not None and world > 1:
            mine = deal_balanced(chunk, np.asarray(cost)[chunk], world)[rank]
        elif g == gb:
            mine = chunk[rank * batch_size:(rank + 1) * batch_size]
        else:
            base, extra = divmod(g, world)
            lo = rank * base + min(rank, extra)
            mine = chunk[lo:lo + base + (1 if rank < extra else 0)]
        out.append((mine, g))
    return out


def _u64_as_i64(x):
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


class TrainEngine(object):
    """One training step as a replayable CUDA graph (see module docstring)."""

    RING = 16
    LOSS_RING = 1024          # per-step losses the host can read back late (power of two)

    def __init__(self, dataset, model, optimizer, batch_size, ARR=0.0, use_graph=True, sample_seed=0):
        self.dataset, self.model, self.opt = dataset, model, optimizer
        self.B, self.ARR = int(batch_size), float(ARR)
        self.rank, self.world = _dist_info()
        self.dev = model.flat_params.device
        self.use_graph = bool(use_graph) and hasattr(dataset, "extractor") and not hasattr(dataset, "slices")
        self.sample_seed = int(sample_seed)
        self.graphs = {}
        self.eager_done = set()
        # [idx(B) | sample_seed | drop_seed | global batch size | drop_seed of the NEXT step]  per step, pinned ring +
        # one device copy (the next step's seed is what the list images of the batch being extracted are built with)
        self.stepbuf_dev = torch.zeros(self.B + 4, dtype=torch.int64, device=self.dev)
        self.ring = [torch.zeros(self.B + 4, dtype=torch.int64).pin_memory() for _ in range(self.RING)]
        self.ring_np = [t.numpy() for t in self.ring]          # host-side views: filling a slot is a plain memcpy
        self.ring_ev = [torch.cuda.Event() for _ in range(self.RING)]
        self.ring_used = [False] * self.RING
        self.ring_pos = 0
        self._seed_cache = {}
        self.lr_dev = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.loss_acc = torch.zeros(1, dtype=torch.float32, device=self.dev)   # sum_steps loss*G (this rank)
        self.last_loss = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.steps = 0
        # zero-copy step I/O (pipelined engine): the kernels of a step read its inputs straight out of one of two
        # pinned host buffers and the update kernel stores the step's loss into a pinned host ring - no copy-engine
        # operation between two graph replays (the two per-step memcpys cost ~45 us of serialisation per 180 us step)
        self.zero_copy = self.dev.type == "cuda" and os.environ.get("IGMC_ZERO_COPY", "1") != "0"
        if self.zero_copy:
            from .util_functions import mapped_view
            self.hostbuf = [torch.zeros(self.B + 4, dtype=torch.int64).pin_memory() for _ in range(2)]
            self.hostbuf_np = [t.numpy() for t in self.hostbuf]
            self.hostview = [mapped_view(t) for t in self.hostbuf]
            self.slot_ev = [torch.cuda.Event() for _ in range(2)]
            self.slot_used = [False, False]
            self.loss_ring = torch.zeros(self.LOSS_RING, dtype=torch.float32).pin_memory()
            self.loss_ring_view = mapped_view(self.loss_ring)
        self._adam_base = None    # Adam step number before this engine's first update (index base of the loss ring)
        self._updates = 0
        self.exchange = None      # peer-mapped gradient buffers of the fused update kernel (created on first use)
        self.fused_update = os.environ.get("IGMC_FUSED_UPDATE", "1") != "0"
        self.sync_lr()

    def sync_lr(self):
        self.lr_dev.fill_(float(self.opt.param_groups[0]["lr"]))

    def drop_seed(self, step):
        """dropout stream of optimisation step ``step`` on this rank (ranks draw independent masks)."""
        return splitmix64(self.model.drop_seed + step * self.world + self.rank)

    @property
    def arr_local(self):
        """the regulariser is added once per global batch: rank 0 carries it (every global batch gives rank 0 at
        least one graph), so that the all-reduced gradient and the summed loss equal the single-process ones."""
        return self.ARR if self.rank == 0 else 0.0

    # ---- the launches of one step, reading everything from stepbuf_dev ----------------------------
    def _model_step(self, batch, nb, G, seed_dev):
        """model half of a step on an extracted batch: [prep, forward+loss, backward] then either ONE fused kernel
        (gradient assembly + ARR -> all-reduce over NVLink peer memory -> Adam) or, for plans without raw partial
        rows / external readouts, gradient assembly -> NCCL all-reduce -> Adam."""
        m = self.model
        if self.fused_update and (m.fused_update_ok(batch) if nb > 0 else self.exchange is not None):
            if self.exchange is None:
                from .exchange import Exchange
                self.exchange = Exchange(m.flat_params.numel(), self.dev, self.rank, self.world)
            if nb > 0:
                ws = m.forward_backward(batch, global_num_graphs=G, seed_dev=seed_dev)["ws"]
                self._last_ws = ws
                rows = nb * ws["cluster"]
            else:   # no graph of a short tail batch on this rank: it still takes part in the exchange (zero rows)
                ws, rows = self._last_ws, 0
            loss = self.opt.reduce_update(self.exchange, ws, nb, rows, 1.0 / max(G, 1), self.arr_local,
                                          lr_dev=self.lr_dev, loss_acc=self.loss_acc, loss_weight=float(G),
                                          loss_ring=self.loss_ring_view if self.zero_copy else None)
        else:
            if nb > 0:
                loss = m.fused_step(batch, ARR=self.arr_local, global_num_graphs=G, seed_dev=seed_dev)
            else:
                m.flat_grad.zero_()
                loss = torch.zeros(1, dtype=torch.float32, device=self.dev)
            if self.world > 1:
                dist.all_reduce(m.flat_grad)
            self.opt.step(grad_mul=1.0, lr_dev=self.lr_dev, loss_in=loss, loss_acc=self.loss_acc, loss_weight=float(G))
        self.last_loss = loss

    def _launch(self, nb, G):
        B = self.B
        buf = self.stepbuf_dev
        batch = None
        if nb > 0:
            batch = self.dataset.extractor.extract(idx=buf[:nb], seed_dev=buf[B:B + 1], reuse=True)
        self._model_step(batch, nb, G, buf[B + 1:B + 2])

    def _launch_static(self, idx, G):
        batch = self.dataset.extract_batch(idx)
        loss = self.model.fused_step(batch, ARR=self.arr_local, global_num_graphs=G)
        if self.world > 1:
            dist.all_reduce(self.model.flat_grad)
        self.opt.step(grad_mul=1.0, lr_dev=self.lr_dev, loss_in=loss, loss_acc=self.loss_acc, loss_weight=float(G))
        self.last_loss = loss

    def stage(self, idx, epoch, G):
        """fill the next pinned slot with this step's inputs and enqueue its H2D copy."""
        # (the host side of a step is on the critical path of the end-to-end rate: one memcpy into pinned memory,
        # one async H2D copy, one event record)
        slot = self.ring_pos
        self.ring_pos = (slot + 1) % self.RING
        ev = self.ring_ev[slot]
        if self.ring_used[slot]:
            ev.synchronize()                   # the copy that last used this slot has finished
        nb = self._fill(self.ring_np[slot], idx, epoch, G)
        self.stepbuf_dev.copy_(self.ring[slot], non_blocking=True)
        ev.record()
        self.ring_used[slot] = True
        return nb

    def step(self, idx, epoch=0, G=None, staged=False):
        """One optimisation step on this rank's ``idx`` (host int array).  H2D copy + graph replay."""
        G = len(idx) * self.world if G is None else int(G)
        self.steps += 1
        if not hasattr(self.dataset, "extractor") or hasattr(self.dataset, "slices"):
            return self._launch_static(np.asarray(idx), G)
        nb = len(idx) if staged else self.stage(idx, epoch, G)
        key = (nb, G)
        if self.use_graph and key in self.graphs:
            self.graphs[key].replay()
        elif self.use_graph and key in self.eager_done and nb > 0:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch(nb, G)
            self.graphs[key] = g
            g.replay()
        else:
            self._launch(nb, G)
            self.eager_done.add(key)

    # ---- software pipeline: extract batch k+1 on a side stream while the model trains on batch k ------------
    # (what the reference's DataLoader workers do, train_eval.py:40-45, as two branches of one CUDA graph)
    def pipelined(self):
        return hasattr(self.dataset, "extractor") and not hasattr(self.dataset, "slices")

    def loss_of_update(self, k):
        """loss of this engine's k-th update (0-based) read from the pinned ring; valid after the step has finished
        (any stream / device sync) and for ``LOSS_RING`` further steps."""
        return float(self.loss_ring[(self._adam_base + 1 + k) & (self.LOSS_RING - 1)])

    def _fill(self, h, idx, epoch, G):
        """[idx | sample seed | dropout seed of this step | global batch | dropout seed of the next step] -> ``h``"""
        nb = len(idx)
        h[:nb] = idx
        B = self.B
        ss = self._seed_cache.get(epoch)
        if ss is None:
            ss = self._seed_cache[epoch] = _u64_as_i64(splitmix64(self.sample_seed + epoch))
        h[B] = ss
        nxt = _u64_as_i64(self.drop_seed(self.steps + 1))
        cur = self._next_seed if getattr(self, "_next_step", None) == self.steps else _u64_as_i64(self.drop_seed(self.steps))
        self._next_seed, self._next_step = nxt, self.steps + 1
        h[B + 1] = cur
        h[B + 2] = G
        h[B + 3] = nxt
        return nb

    def _launch_pipe(self, nb, G, slot, nb_next, buf=None):
        B = self.B
        buf = self.stepbuf_dev if buf is None else buf
        main = torch.cuda.current_stream()
        if nb_next > 0:
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                if nb > 0:
                    # hold the extraction back until every cluster of this step's forward is resident: the forward
                    # needs 100 SMs to itself (one CTA per SM); extraction CTAs that start first spread one per SM over
                    # 50 SMs, leave room for 49 of the 50 clusters and cost the forward a second wave (measured: 235
                    # us per step instead of 180, profiles/README.md).  Started behind it they pack two per SM on the
                    # 48 SMs that are left.
                    self.model.gate_wait(self.batches[slot])
                self.batches[slot ^ 1] = self.dataset.extractor.extract(idx=buf[:nb_next], seed_dev=buf[B:B + 1],
                                                                        reuse=True, slot=slot ^ 1)
                # the next step's edge lists (after its dropout draws), staged for the model kernels' bulk loads
                self.model.stage_batch(self.batches[slot ^ 1], True,
                                       self.model.make_dropout(True, seed_dev=buf[B + 3:B + 4]), slot=slot ^ 1)
        self._model_step(self.batches[slot] if nb > 0 else None, nb, G, buf[B + 1:B + 2])
        if nb_next > 0:
            main.wait_stream(self.side)

    def prime(self, idx, epoch=0, G=None):
        """extract the first batch of a pipelined sequence (no model work)."""
        if not hasattr(self, "side"):
            self.side = torch.cuda.Stream()
            self.batches = [None, None]
        G = len(idx) * self.world if G is None else int(G)
        nb = self.stage(idx, epoch, G)
        self.slot = 0
        if nb > 0:
            self.batches[0] = self.dataset.extractor.extract(idx=self.stepbuf_dev[:nb],
                                                             seed_dev=self.stepbuf_dev[self.B:self.B + 1], reuse=True,
                                                             slot=0)
            self.model.stage_batch(self.batches[0], True,
                                   self.model.make_dropout(True, seed_dev=self.stepbuf_dev[self.B + 3:self.B + 4]),
                                   slot=0)
        self.cur = (nb, G)

    def step_pipe(self, next_idx=None, epoch=0, next_G=None, staged=False):
        """train on the batch extracted by the previous call (or prime) while extracting ``next_idx``.
        One H2D copy of the next indices + one graph replay."""
        nb, G = self.cur
        self.steps += 1
        if self._adam_base is None:
            self._adam_base = int(self.opt.step_count[0].item())      # one sync, before the first step only
        nb_next = 0
        if next_idx is not None:
            next_G = len(next_idx) * self.world if next_G is None else int(next_G)
        zc = self.zero_copy and not staged
        buf = None
        if staged:                 # bench.py's device-resident inputs: stepbuf_dev was filled by the caller
            nb_next = len(next_idx) if next_idx is not None else 0
        elif zc:                   # this step's inputs go into the pinned buffer its graph variant reads
            if self.slot_used[self.slot]:
                self.slot_ev[self.slot].synchronize()    # the replay that last read this buffer has finished
            nb_next = self._fill(self.hostbuf_np[self.slot], next_idx if next_idx is not None else (), epoch,
                                 next_G if next_idx is not None else 0)
            buf = self.hostview[self.slot]
        elif next_idx is not None:
            nb_next = self.stage(next_idx, epoch, next_G)
        else:
            self.stage(np.zeros(0, np.int64), epoch, 0)    # still refresh the dropout seed of this step
        key = (nb, G, self.slot, nb_next, zc)
        if self.use_graph and key in self.graphs:
            self.graphs[key].replay()
        elif self.use_graph and key in self.eager_done:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch_pipe(nb, G, self.slot, nb_next, buf)
            self.graphs[key] = g
            g.replay()
        else:
            self._launch_pipe(nb, G, self.slot, nb_next, buf)
            self.eager_done.add(key)
        if zc:
            self.slot_ev[self.slot].record()
            self.slot_used[self.slot] = True
        self._updates += 1
        self.slot ^= 1
        self.cur = (nb_next, next_G if next_idx is not None else 0)

    def close(self):
        """drop the captured graphs and release the peer-mapped exchange buffers (collective under data parallelism:
        every rank calls it, nobody may still be reading a buffer that is being freed)"""
        torch.cuda.synchronize()
        self.graphs.clear()
        if self.exchange is not None:
            self.exchange.close()
            self.exchange = None

    def check(self):
        code = int(self.dataset.extractor.err.item()) if hasattr(self.dataset, "extractor") else 0
        if code:
            from . import _lib
            raise RuntimeError("igmc_b200 kernel error %d: %s" % (code, _lib.ERR_NAMES.get(code, "?")))


def train_multiple_epochs(train_dataset, test_dataset, model, epochs, batch_size, lr, lr_decay_factor,
                          lr_decay_step_size, weight_decay, ARR=0, test_freq=1, logger=None,
                          continue_from=None, res_dir=None, use_graph=True, seed=1):
    """Reference train_eval.py:23-111.  Returns the last test RMSE."""
    rank, world = _dist_info()
    rmses = []
    model.to(device).reset_parameters()
    if world > 1:  # identical initial weights on every rank
        dist.broadcast(model.flat_params, 0)
    optimizer = FusedAdam(model, lr=lr, weight_decay=weight_decay)
    start_epoch = 1
    if continue_from is not None:
        model.load_state_dict(torch.load(os.path.join(res_dir, "model_checkpoint{}.pth".format(continue_from))))
        optimizer.load_state_dict(
            torch.load(os.path.join(res_dir, "optimizer_checkpoint{}.pth".format(continue_from))))
        start_epoch = continue_from + 1
        epochs -= continue_from
    engine = TrainEngine(train_dataset, model, optimizer, batch_size, ARR, use_graph, sample_seed=seed)
    torch.cuda.synchronize()
    t_start = time.perf_counter()
    gen = torch.Generator().manual_seed(int(seed))
    if start_epoch > 1:
        # resumed run: continue the dropout / permutation streams where the checkpointed run left them instead of
        # replaying epoch 1's (the sampling stream is keyed by the epoch number already)
        engine.steps = int(optimizer.step_count[0].item())
        for _ in range(start_epoch - 1):
            torch.randperm(len(train_dataset), generator=gen)
    for epoch in range(start_epoch, epochs + start_epoch):
        train_loss = train(model, optimizer, train_dataset, device, regression=True, ARR=ARR, epoch=epoch,
                           engine=engine, generator=gen)
        if epoch % test_freq == 0:
            rmses.append(eval_rmse(model, test_dataset, device, batch_size=batch_size))
        else:
            rmses.append(np.nan)
        eval_info = {"epoch": epoch, "train_loss": train_loss, "test_rmse": rmses[-1]}
        if rank == 0:
            print("Epoch {}, train loss {:.6f}, test rmse {:.6f}".format(*eval_info.values()))
        if epoch % lr_decay_step_size == 0:
            for param_group in optimizer.param_groups:
                param_group["lr"] = lr_decay_factor * param_group["lr"]
            engine.sync_lr()
        if logger is not None and rank == 0:
            logger(eval_info, model, optimizer)
    torch.cuda.synchronize()
    duration = time.perf_counter() - t_start
    if rank == 0:
        print("Final Test RMSE: {:.6f}, Duration: {:.6f}".format(rmses[-1], duration))
    engine.close()
    return rmses[-1]


def train(model, optimizer, loader, device, regression=False, ARR=0, show_progress=False, epoch=None,
          engine=None, generator=None, batch_size=50):
    """One epoch (reference train_eval.py:149-179).  ``loader`` is the train dataset; mini-batches are
    drawn from a seeded permutation identical on every rank.  Returns sum(loss*num_graphs)/len(dataset)
    (global under data parallelism)."""
    dataset = loader
    model.train()
    if engine is None:
        engine = TrainEngine(dataset, model, optimizer, batch_size, ARR)
    rank, world = _dist_info()
    perm = torch.randperm(len(dataset), generator=generator).numpy()
    engine.loss_acc.zero_()
    ep = 0 if epoch is None else int(epoch)
    cost = dataset.pair_cost() if world > 1 and hasattr(dataset, "pair_cost") else None
    batches = shard_batches(perm, engine.B, rank, world, cost)
    if engine.pipelined() and len(batches) > 0:
        engine.prime(batches[0][0], ep, batches[0][1])
        for k in range(len(batches)):
            if k + 1 < len(batches):
                engine.step_pipe(batches[k + 1][0], ep, batches[k + 1][1])
            else:
                engine.step_pipe(None, ep)
    else:
        for idx, G in batches:
            engine.step(idx, epoch=ep, G=G)
    acc = engine.loss_acc.clone()
    if world > 1:
        dist.all_reduce(acc)
    engine.check()
    return float(acc.item()) / len(dataset)


def _eval_sqerr_sum(model, dataset, batch_size):
    """sum of squared errors of this rank's share of ``dataset`` (device scalar), eval mode, no grad."""
    rank, world = _dist_info()
    dev = model.flat_params.device
    acc = torch.zeros(1, dtype=torch.float32, device=dev)
    order = np.arange(len(dataset), dtype=np.int64)
    drop = model.make_dropout(False)
    for idx, _ in shard_batches(order, batch_size, rank, world):
        if len(idx) == 0:
            continue
        batch = dataset.extract_batch(idx)
        _, saved = model._launch_forward(batch, False, drop, y=batch.y, loss_scale=0.0)
        acc += saved["ws"]["sqerr"].sum()
    _check_dataset(dataset)      # a kernel error during evaluation must not turn into a silent garbage RMSE
    if world > 1:
        dist.all_reduce(acc)
    return acc


def _check_dataset(dataset):
    ex = getattr(dataset, "extractor", None)
    err = getattr(ex, "err", None)
    if err is not None:
        code = int(err.item())
        if code:
            from . import _lib
            raise RuntimeError("igmc_b200 kernel error %d: %s" % (code, _lib.ERR_NAMES.get(code, "?")))


def eval_loss(model, loader, device, regression=False, show_progress=False, batch_size=50):
    """Reference train_eval.py:182-199: mean squared error over the dataset (sum reduction / N)."""
    model.eval()
    with torch.no_grad():
        s = _eval_sqerr_sum(model, loader, batch_size)
    return float(s.item()) / len(loader)


def eval_rmse(model, loader, device, show_progress=False, batch_size=50):
    return math.sqrt(eval_loss(model, loader, device, True, show_progress, batch_size))


def eval_loss_ensemble(model, checkpoints, loader, device, regression=False, show_progress=False, batch_size=50):
    """Reference train_eval.py:208-238: average the predictions of several checkpoints, then MSE."""
    dataset = loader
    rank, world = _dist_info()
    dev = model.flat_params.device
    order = np.arange(len(dataset), dtype=np.int64)
    shards = [idx for idx, _ in shard_batches(order, batch_size, rank, world) if len(idx)]   # this rank's share
    outs, ys = [], None
    for i, checkpoint in enumerate(checkpoints):
        model.load_state_dict(torch.load(checkpoint))
        model.eval()
        o, y = [], []
        drop = model.make_dropout(False)
        with torch.no_grad():
            for idx in shards:
                batch = dataset.extract_batch(idx)
                pred, _ = model._launch_forward(batch, False, drop)
                o.append(pred.clone())
                if i == 0:
                    y.append(batch.y.clone())
        outs.append(torch.cat(o).view(-1, 1) if o else torch.zeros(0, 1, device=dev))
        if i == 0:
            ys = torch.cat(y) if y else torch.zeros(0, device=dev)
    _check_dataset(dataset)
    mean = torch.cat(outs, 1).mean(1)
    sq = ((mean - ys) ** 2).sum().view(1)
    if world > 1:
        dist.all_reduce(sq)
    return float(sq.item()) / len(dataset)


def eval_rmse_ensemble(model, checkpoints, loader, device, show_progress=False, batch_size=50):
    return math.sqrt(eval_loss_ensemble(model, checkpoints, loader, device, True, show_progress, batch_size))


def test_once(test_dataset, model, batch_size, logger=None, ensemble=False, checkpoints=None):
    """Reference train_eval.py:114-139."""
    model.to(device)
    t_start = time.perf_counter()
    if ensemble and checkpoints:
        rmse = eval_rmse_ensemble(model, checkpoints, test_dataset, device, batch_size=batch_size)
    else:
        rmse = eval_rmse(model, test_dataset, device, batch_size=batch_size)
    duration = time.perf_counter() - t_start
    rank, _ = _dist_info()
    if rank == 0:
        print("Test Once RMSE: {:.6f}, Duration: {:.6f}".format(rmse, duration))
    eval_info = {"epoch": "test_once" if not ensemble else "ensemble", "train_loss": 0, "test_rmse": rmse}
    if logger is not None and rank == 0:
        logger(eval_info, None, None)
    return rmse


def visualize(*args, **kwargs):
    raise NotImplementedError("visualisation (train_eval.py:248-322) is reporting code outside the hot path")
