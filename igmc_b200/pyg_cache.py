"""The reference's on-disk cache of the static dataset: ``<root>/processed/data.pt`` (or ``data_<max_num>.pt``),
written by ``MyDataset.process`` as ``torch.save((data, slices), path)`` (util_functions.py:92-110) with PyG 1.4.2's
``InMemoryDataset.collate`` semantics:

* ``data``   one ``torch_geometric.data.Data`` holding every graph's tensors concatenated - ``x [sum n, 2h+2] f32``,
  ``edge_index [2, sum e] i64`` with graph-LOCAL node ids (no offsets), ``edge_type [sum e] i64``, ``y [G] f32``;
* ``slices`` ``{key: LongTensor [G+1]}`` boundaries of every key (x: nodes, edge_index / edge_type: edges, y: 0..G).

``save_processed`` writes exactly that, so a reference installation opens the file with its own ``MyDataset``;
``load_processed`` reads files written by either side.  PyG itself is not needed: the pickle only names the class
``torch_geometric.data.data.Data`` and restores its ``__dict__``; when PyG is absent a stand-in class of that name is
registered for the duration of the (un)pickling.
"""
import os
import sys
import types

import torch

KEYS = ("x", "edge_index", "y", "edge_type")      # order of Data.keys for construct_pyg_graph's Data (util_functions.py:287)


def _data_class():
    """``torch_geometric.data.data.Data``: the real one if PyG is importable, else a stand-in of the same name whose
    attribute set is PyG 1.4.2's ``Data.__init__`` (x, edge_index, edge_attr, y, pos, norm, face + keyword extras)."""
    mod = sys.modules.get("torch_geometric.data.data")
    if mod is not None and hasattr(mod, "Data"):
        return mod.Data
    try:
        from torch_geometric.data.data import Data     # noqa: F401  (a real installation)
        return Data
    except Exception:
        pass

    class Data(object):
        def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, norm=None, face=None, **kwargs):
            self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
            self.pos, self.norm, self.face = pos, norm, face
            for k, v in kwargs.items():
                setattr(self, k, v)

        @property
        def keys(self):
            return [k for k, v in self.__dict__.items() if v is not None]

        def __getitem__(self, key):
            return getattr(self, key)

    Data.__module__ = "torch_geometric.data.data"
    Data.__qualname__ = "Data"
    tg = sys.modules.setdefault("torch_geometric", types.ModuleType("torch_geometric"))
    tgd = sys.modules.get("torch_geometric.data")
    if tgd is None:
        tgd = sys.modules["torch_geometric.data"] = types.ModuleType("torch_geometric.data")
        tg.data = tgd
    tgdd = types.ModuleType("torch_geometric.data.data")
    tgdd.Data = Data
    sys.modules["torch_geometric.data.data"] = tgdd
    tgd.data = tgdd
    if not hasattr(tgd, "Data"):
        tgd.Data = Data
    return Data


def processed_path(root, max_num=None):
    """reference util_functions.py:94-99"""
    name = "data.pt" if max_num is None else "data_{}.pt".format(max_num)
    return os.path.join(root, "processed", name)


def save_processed(path, x, edge_index, edge_type, y, node_off, edge_off):
    """``node_off`` / ``edge_off``: [G+1] boundaries; ``edge_index`` holds graph-local node ids."""
    Data = _data_class()
    G = int(y.numel())
    data = Data(x=x.detach().cpu().float().contiguous(), edge_index=edge_index.detach().cpu().long().contiguous(),
                y=y.detach().cpu().float().contiguous(), edge_type=edge_type.detach().cpu().long().contiguous())
    node_off = torch.as_tensor(node_off).detach().cpu().long()
    edge_off = torch.as_tensor(edge_off).detach().cpu().long()
    slices = {"x": node_off.clone(), "edge_index": edge_off.clone(), "y": torch.arange(G + 1, dtype=torch.long),
              "edge_type": edge_off.clone()}
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save((data, slices), path)


def load_processed(path):
    """-> dict(x, edge_index, edge_type, y, node_off, edge_off) of CPU tensors; raises ValueError on files that are not
    the reference's (data, slices) layout."""
    _data_class()
    try:
        data, slices = torch.load(path, map_location="cpu", weights_only=False)
    except TypeError:      # torch without the weights_only keyword
        data, slices = torch.load(path, map_location="cpu")
    get = (lambda k: data[k]) if hasattr(data, "__getitem__") else (lambda k: getattr(data, k))
    try:
        x, ei, et, y = get("x"), get("edge_index"), get("edge_type"), get("y")
        node_off, edge_off = slices["x"].long(), slices["edge_index"].long()
    except Exception as e:
        raise ValueError("%s is not a (data, slices) cache of IGMC subgraphs: %s" % (path, e))
    G = int(y.numel())
    if node_off.numel() != G + 1 or edge_off.numel() != G + 1 or ei.dim() != 2 or ei.shape[0] != 2 \
            or int(node_off[-1]) != x.shape[0] or int(edge_off[-1]) != ei.shape[1] or et.numel() != ei.shape[1]:
        raise ValueError("%s: inconsistent (data, slices)" % path)
    if not torch.equal(slices["edge_type"].long(), edge_off):
        raise ValueError("%s: edge_type and edge_index slices differ" % path)
    return dict(x=x.float(), edge_index=ei.long(), edge_type=et.long(), y=y.float().view(-1), node_off=node_off,
                edge_off=edge_off)
