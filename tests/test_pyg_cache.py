"""The reference's ``processed/data.pt`` interchange (util_functions.py:92-110), CPU part: files
written by igmc_b200 open with PyG-1.4.2-style ``InMemoryDataset`` slicing under a foreign ``Data`` class, and files
written the reference's way load back.  (The GPU part - MyDataset(root=...) - is in tests/test_gpu_train.py.)"""
import numpy as np
import torch

from igmc_b200 import pyg_cache
from igmc_b200.data import build_adj, synth_ratings
from oracle import extract_np


def _graphs(n=9):
    u, v, lab = synth_ratings(40, 30, 300, 5, 2)
    g = extract_np.RatingCSR(build_adj(u, v, lab, 40, 30))
    cv = np.arange(1, 6, dtype=np.float32)
    out = []
    for k in range(n):
        sub = extract_np.extract_subgraph(g, u[k], v[k], 1, 1.0, 8, seed=0, pair_id=k)
        out.append(extract_np.construct_graph(sub, cv[lab[k]], 1))
    return out


def _collate_like_inmemorydataset(graphs):
    """PyG 1.4.2 InMemoryDataset.collate: concatenate every key, NO node offsets, slices = boundaries"""
    xs = torch.cat([torch.from_numpy(g["x"]) for g in graphs])
    ei = torch.cat([torch.from_numpy(g["edge_index"]) for g in graphs], 1)
    et = torch.cat([torch.from_numpy(g["edge_type"]) for g in graphs])
    y = torch.cat([torch.from_numpy(np.asarray(g["y"], np.float32)).view(-1) for g in graphs])
    noff = torch.tensor([0] + list(np.cumsum([g["x"].shape[0] for g in graphs])))
    eoff = torch.tensor([0] + list(np.cumsum([g["edge_index"].shape[1] for g in graphs])))
    return xs, ei, et, y, noff, eoff


def test_written_file_opens_with_inmemorydataset_slicing(tmp_path):
    graphs = _graphs()
    xs, ei, et, y, noff, eoff = _collate_like_inmemorydataset(graphs)
    path = pyg_cache.processed_path(str(tmp_path), None)
    assert path.endswith("processed/data.pt") and pyg_cache.processed_path("r", 500).endswith("processed/data_500.pt")
    pyg_cache.save_processed(path, xs, ei, et, y, noff, eoff)
    data, slices = torch.load(path, weights_only=False)
    assert type(data).__module__ == "torch_geometric.data.data" and type(data).__name__ == "Data"
    assert sorted(slices) == ["edge_index", "edge_type", "x", "y"]
    for i, g in enumerate(graphs):        # InMemoryDataset.get(i): slice every key along its cat dim
        s = {k: (int(slices[k][i]), int(slices[k][i + 1])) for k in slices}
        assert np.array_equal(data.x[s["x"][0]:s["x"][1]].numpy(), g["x"])
        assert np.array_equal(data.edge_index[:, s["edge_index"][0]:s["edge_index"][1]].numpy(), g["edge_index"])
        assert np.array_equal(data.edge_type[s["edge_type"][0]:s["edge_type"][1]].numpy(), g["edge_type"])
        assert float(data.y[s["y"][0]:s["y"][1]]) == float(np.float32(np.asarray(g["y"]).reshape(-1)[0]))
    assert int(data.edge_index.max()) < max(g["x"].shape[0] for g in graphs)      # graph-LOCAL node ids


def test_reference_style_file_loads(tmp_path):
    """a cache pickled the reference's way (its own Data object with extra None attributes, long slices)"""
    graphs = _graphs(5)
    xs, ei, et, y, noff, eoff = _collate_like_inmemorydataset(graphs)
    Data = pyg_cache._data_class()
    d = Data(x=xs, edge_index=ei, y=y, edge_type=et)
    sl = {"x": noff, "edge_index": eoff, "y": torch.arange(6), "edge_type": eoff}
    p = tmp_path / "processed" / "data.pt"
    p.parent.mkdir()
    torch.save((d, sl), p)
    got = pyg_cache.load_processed(str(p))
    assert torch.equal(got["x"], xs) and torch.equal(got["edge_index"], ei) and torch.equal(got["edge_type"], et)
    assert torch.equal(got["y"], y) and torch.equal(got["node_off"], noff) and torch.equal(got["edge_off"], eoff)
    torch.save((d, {"x": noff[:-1], "edge_index": eoff, "y": torch.arange(6), "edge_type": eoff}), p)
    try:
        pyg_cache.load_processed(str(p))
        assert False, "inconsistent slices must be refused"
    except ValueError:
        pass
