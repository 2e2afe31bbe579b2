"""Pure-torch CPU restatement of the model half of the IGMC hot path (TEST ORACLE).

PARITY UNPINNED: the arithmetic lives in third-party PyTorch-Geometric 1.4.2
(pinned only in prose, reference README.md:26), which is neither under
/root/reference nor installable offline, and the reference ships no tests, golden
outputs or checkpoints.  This file restates the published PyG 1.4.2 algorithms
 and anchors on the reference's own call sites:

* ``RGCNConv(in, out, num_relations, num_bases)``  built at models.py:182-184,
  called at models.py:201; parameters ``att [R,B]``, ``basis [B,in,out]`` confirmed
  by train_eval.py:168-172; ``root [in,out]``, ``bias [out]``; ``aggr='mean'`` over
  ALL incoming edges; init ``U(-1/sqrt(B*in), 1/sqrt(B*in))`` for every parameter.
* ``dropout_adj``  call site models.py:193-198 (Bernoulli(1-p) per directed edge).
* ``IGMC.forward``  models.py:190-217 (concat of tanh layer outputs, target-row
  readout, lin1 -> relu -> dropout(0.5) -> lin2).
* train-step loss  train_eval.py:157-175 (MSE mean + ARR * sum_r ||W_{r+1}-W_r||^2).
* ``global_sort_pool``  call site models.py:155 and the ``DGCNN_RS`` readout models.py:156-167.

The message function deliberately uses the reference-era formulation
(``index_select`` of a per-edge weight + ``bmm`` + scatter-mean): it is the CPU
baseline that ``bench.py --impl reference`` times.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def scatter_mean(src, index, dim_size):
    out = torch.zeros(dim_size, src.shape[1], dtype=src.dtype, device=src.device).index_add_(0, index, src)
    cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device).index_add_(
        0, index, torch.ones_like(index, dtype=src.dtype))
    return out / cnt.clamp(min=1).unsqueeze(1)


def dropout_adj(edge_index, edge_attr, p, training=True, keep_mask=None, generator=None):
    """PyG 1.4.2 ``dropout_adj`` with force_undirected=False.  ``keep_mask`` injects the
    Bernoulli draw (bool [E]) so that two implementations can share it."""
    if not training or p == 0.0:
        return edge_index, edge_attr
    if keep_mask is None:
        keep_mask = torch.bernoulli(torch.full((edge_index.shape[1],), 1 - p, device=edge_index.device),
                                    generator=generator).bool()
    return edge_index[:, keep_mask], edge_attr[keep_mask]


class RGCNConvRef(nn.Module):
    def __init__(self, in_channels, out_channels, num_relations, num_bases, aggr="mean_all"):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_relations, self.num_bases = num_relations, num_bases
        self.aggr = aggr
        self.basis = nn.Parameter(torch.empty(num_bases, in_channels, out_channels))
        self.att = nn.Parameter(torch.empty(num_relations, num_bases))
        self.root = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.num_bases * self.in_channels)
        for p in (self.basis, self.att, self.root, self.bias):
            p.data.uniform_(-bound, bound)

    # "bmm": the reference-era message function (per-edge weight gather + bmm) - the formulation bench.py times as the
    # baseline.  "transform": the same sum re-associated (x W_r for every relation first, then a row gather); it avoids
    # the [E, in, out] tensor so that fp64 parity tests can run BASELINE-size batches (tests/test_oracle_model.py pins
    # the two against each other).
    formulation = "bmm"

    def forward(self, x, edge_index, edge_type):
        src, dst = edge_index[0], edge_index[1]
        w = torch.matmul(self.att, self.basis.view(self.num_bases, -1))
        w = w.view(self.num_relations, self.in_channels, self.out_channels)
        if self.formulation == "transform":
            y = torch.einsum("ni,rio->nro", x, w)                        # [N, R, out]
            msg = y[src, edge_type]                                       # [E, out]
        else:
            w_e = torch.index_select(w, 0, edge_type)                     # [E, in, out]
            msg = torch.bmm(x[src].unsqueeze(1), w_e).squeeze(1)          # [E, out]
        n = x.shape[0]
        if self.aggr == "mean_all":
            agg = scatter_mean(msg, dst, n)
        elif self.aggr == "add":
            agg = torch.zeros(n, msg.shape[1], dtype=msg.dtype, device=msg.device).index_add_(0, dst, msg)
        else:
            raise ValueError(self.aggr)
        return agg + x @ self.root + self.bias


class IGMCRef(nn.Module):
    """Restated ``IGMC`` (models.py:170-217); state_dict keys equal the reference's."""

    def __init__(self, num_features=4, latent_dim=(32, 32, 32, 32), num_relations=5, num_bases=4,
                 adj_dropout=0.2, multiply_by=1, aggr="mean_all"):
        super().__init__()
        self.adj_dropout, self.multiply_by = adj_dropout, multiply_by
        dims = [num_features] + list(latent_dim)
        self.convs = nn.ModuleList(
            [RGCNConvRef(dims[l], dims[l + 1], num_relations, num_bases, aggr) for l in range(len(latent_dim))])
        self.lin1 = nn.Linear(2 * sum(latent_dim), 128)
        self.lin2 = nn.Linear(128, 1)

    def reset_parameters(self):
        for c in self.convs:
            c.reset_parameters()
        self.lin1.reset_parameters()
        self.lin2.reset_parameters()

    def forward(self, x, edge_index, edge_type, edge_keep=None, hidden_keep=None, return_states=False):
        """``edge_keep`` bool [E] / ``hidden_keep`` bool [B,128]: injected dropout draws (training
        semantics).  With ``self.training`` False no dropout is applied at all."""
        x0 = x
        if self.training and self.adj_dropout > 0:
            edge_index, edge_type = dropout_adj(edge_index, edge_type, self.adj_dropout, True, edge_keep)
        states = []
        for conv in self.convs:
            x = torch.tanh(conv(x, edge_index, edge_type))
            states.append(x)
        cs = torch.cat(states, 1)
        users, items = x0[:, 0] == 1, x0[:, 1] == 1
        z = torch.cat([cs[users], cs[items]], 1)
        z = F.relu(self.lin1(z))
        if self.training:
            if hidden_keep is not None:
                z = z * hidden_keep.to(z.dtype) * 2.0
            else:
                z = F.dropout(z, p=0.5, training=True)
        out = self.lin2(z)[:, 0] * self.multiply_by
        return (out, cs) if return_states else out


def global_sort_pool(x, batch, k, num_graphs=None):
    """PyG 1.4.2 ``global_sort_pool`` (call site models.py:108,155): dense batch padded with
    ``x.min()-1``, rows of every graph sorted by the LAST channel descending, first k rows kept (padded if the
    graph is smaller), padding set to 0, flattened to [B, k*D].  torch's sort is not stable; ties are resolved
    here in favour of the lower node index (``stable=True``), which is what the CUDA path does too."""
    B = int(batch.max()) + 1 if num_graphs is None else int(num_graphs)
    D = x.shape[1]
    fill = float(x.detach().min()) - 1.0
    counts = torch.bincount(batch, minlength=B)
    nmax = int(counts.max())
    start = torch.cumsum(counts, 0) - counts
    pos = torch.arange(x.shape[0], device=x.device) - start[batch]
    dense = torch.full((B, nmax, D), fill, dtype=x.dtype, device=x.device)
    dense[batch, pos] = x
    _, perm = torch.sort(dense[:, :, -1], dim=-1, descending=True, stable=True)
    dense = torch.gather(dense, 1, perm.unsqueeze(-1).expand(-1, -1, D))
    if nmax >= k:
        dense = dense[:, :k]
    else:
        dense = torch.cat([dense, torch.full((B, k - nmax, D), fill, dtype=x.dtype, device=x.device)], 1)
    dense = torch.where(dense == fill, torch.zeros_like(dense), dense)
    return dense.reshape(B, k * D)


class DGCNN_RSRef(nn.Module):
    """Restated ``DGCNN_RS`` (models.py:123-167 on top of ``DGCNN.__init__`` models.py:65-85): R-GCN layers with
    latent_dim [32,32,32,1], SortPooling, Conv1d(1,16,97,97) / MaxPool1d(2,2) / Conv1d(16,32,5,1), dense head."""

    def __init__(self, num_features=4, latent_dim=(32, 32, 32, 1), k=30, num_relations=5, num_bases=2,
                 adj_dropout=0.2, aggr="mean_all"):
        super().__init__()
        self.adj_dropout, self.k = adj_dropout, int(k)
        dims = [num_features] + list(latent_dim)
        self.convs = nn.ModuleList(
            [RGCNConvRef(dims[l], dims[l + 1], num_relations, num_bases, aggr) for l in range(len(latent_dim))])
        self.total_latent_dim = sum(latent_dim)
        self.conv1d_params1 = nn.Conv1d(1, 16, self.total_latent_dim, self.total_latent_dim)
        self.maxpool1d = nn.MaxPool1d(2, 2)
        self.conv1d_params2 = nn.Conv1d(16, 32, 5, 1)
        dense_dim = int((self.k - 2) / 2 + 1)
        self.dense_dim = (dense_dim - 5 + 1) * 32
        self.lin1 = nn.Linear(self.dense_dim, 128)
        self.lin2 = nn.Linear(128, 1)

    def forward(self, x, edge_index, edge_type, batch, edge_keep=None, hidden_keep=None, num_graphs=None,
                return_states=False):
        if self.training and self.adj_dropout > 0:
            edge_index, edge_type = dropout_adj(edge_index, edge_type, self.adj_dropout, True, edge_keep)
        states = []
        for conv in self.convs:
            x = torch.tanh(conv(x, edge_index, edge_type))
            states.append(x)
        cs = torch.cat(states, 1)
        z = global_sort_pool(cs, batch, self.k, num_graphs).unsqueeze(1)
        z = F.relu(self.conv1d_params1(z))
        z = self.maxpool1d(z)
        z = F.relu(self.conv1d_params2(z))
        z = z.view(len(z), -1)
        z = F.relu(self.lin1(z))
        if self.training:
            if hidden_keep is not None:
                z = z * hidden_keep.to(z.dtype) * 2.0
            else:
                z = F.dropout(z, p=0.5, training=True)
        out = self.lin2(z)[:, 0]
        return (out, cs) if return_states else out


def set_formulation(model, formulation):
    """switch every conv of a restated model between the "bmm" and "transform" message formulations"""
    for c in model.convs:
        c.formulation = formulation
    return model


def arr_regulariser(model):
    """train_eval.py:167-174."""
    reg = 0.0
    for g in model.convs:
        w = torch.matmul(g.att, g.basis.view(g.num_bases, -1)).view(g.num_relations, g.in_channels, g.out_channels)
        reg = reg + torch.sum((w[1:] - w[:-1]) ** 2)
    return reg


def train_loss(model, batch, ARR=0.001, edge_keep=None, hidden_keep=None):
    """MSE (mean over graphs) + ARR term (train_eval.py:162,167-174)."""
    out = model(batch["x"], batch["edge_index"], batch["edge_type"], edge_keep, hidden_keep)
    loss = F.mse_loss(out, batch["y"].view(-1))
    if ARR != 0:
        loss = loss + ARR * arr_regulariser(model)
    return loss, out


def to_torch_batch(np_batch, dtype=torch.float32, device=None):
    """``device``: the restatement is device-agnostic; bench.py's ``gpu_baseline`` runs it with torch CUDA ops (the
    reference itself puts its model on CUDA when one exists, train_eval.py:20)."""
    d = dict(x=torch.from_numpy(np_batch["x"]).to(dtype),
             edge_index=torch.from_numpy(np_batch["edge_index"]),
             edge_type=torch.from_numpy(np_batch["edge_type"]),
             y=torch.from_numpy(np_batch["y"]).to(dtype),
             batch=torch.from_numpy(np_batch["batch"]), num_graphs=np_batch["num_graphs"])
    if device is not None:
        d = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in d.items()}
    return d
