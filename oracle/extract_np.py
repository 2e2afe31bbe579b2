"""numpy restatement of the reference's enclosing-subgraph extraction (TEST ORACLE).

Follows, in its own words and in *canonical form* (SURVEY.md §8c):

* ``subgraph_extraction_labeling``  reference util_functions.py:208-277
* ``neighbors``                      reference util_functions.py:300-304
* ``construct_pyg_graph`` / ``one_hot``  reference util_functions.py:280-297, 307-311
* PyG ``Batch.from_data_list`` collate (third-party)

Canonical form: the reference's node order inside a hop is CPython ``set``
iteration order and therefore not a contract; we order every hop's fringe by
ascending global id, and undirected edges by (u_local, v_local).  The reference
output relabelled into that form must equal this module's output bit for bit
(``tests/test_oracle_vs_reference.py`` pins that against the real reference,
``tests/golden`` holds vectors it produced).

Sampling.  The reference draws ``random.sample`` from Python's Mersenne Twister
inside forked DataLoader workers (util_functions.py:222-229); that stream cannot
be reproduced on a GPU.  The sampler is therefore injectable:

* ``sampler=None``       -> counter-hash sampler below (the one the CUDA kernels
                            implement; ``hash_keys`` is its bit-exact restatement)
* ``sampler=callable``   -> ``callable(sorted_candidates, k, side, hop) -> subset``
                            (used to inject the reference's own draws)

Two successive uniform draws without replacement (sample_ratio, then
max_nodes_per_hop) are one uniform draw of the smaller size, so a single
selection of ``k = min(int(ratio*len), mnph)`` is distribution-identical.
"""
import numpy as np

MASK64 = (1 << 64) - 1


def splitmix64(x):
    """Vectorised splitmix64 finaliser on uint64 arrays (wraps mod 2^64)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def hash_keys(seed, pair_id, side, hop, nodes):
    """32-bit sampling keys; bit-exact twin of ``sample_key`` in csrc/common.cuh.

    stream = splitmix64(seed ^ splitmix64(pair_id*4 + side*2 ... )) is folded
    into one 64-bit state per (seed, pair, side, hop); the key of a candidate
    node is the high 32 bits of splitmix64(state + node).
    """
    with np.errstate(over="ignore"):
        tag = np.uint64((int(pair_id) * 16 + int(hop) * 2 + int(side)) & MASK64)
        state = splitmix64(np.uint64(int(seed) & MASK64) ^ splitmix64(tag))
        k = splitmix64(state + np.asarray(nodes, dtype=np.uint64))
    return (k >> np.uint64(32)).astype(np.uint32)


def hash_sample(cands, k, seed, pair_id, side, hop):
    """The k candidates with the smallest (key, node id); returned sorted by id."""
    cands = np.asarray(cands, dtype=np.int64)
    keys = hash_keys(seed, pair_id, side, hop, cands)
    order = np.lexsort((cands, keys))
    return np.sort(cands[order[:k]])


class RatingCSR(object):
    """Flat CSR + CSC of the train rating matrix (values = label + 1).

    Mirrors the *contract* of the reference's SparseRowIndexer/SparseColIndexer
    (util_functions.py:20-66): row/column slices of ``adj_train``.
    """

    def __init__(self, A):
        import scipy.sparse as ssp
        A = ssp.csr_matrix(A)
        A.sum_duplicates()
        A.sort_indices()
        self.shape = A.shape
        self.indptr = A.indptr.astype(np.int64)
        self.indices = A.indices.astype(np.int64)
        self.rating = (np.rint(A.data).astype(np.int64) - 1)  # label = value - 1
        C = A.tocsc()
        C.sort_indices()
        self.cindptr = C.indptr.astype(np.int64)
        self.cindices = C.indices.astype(np.int64)

    def row(self, u):
        s, e = self.indptr[u], self.indptr[u + 1]
        return self.indices[s:e], self.rating[s:e]

    def col(self, v):
        s, e = self.cindptr[v], self.cindptr[v + 1]
        return self.cindices[s:e]


def _union_rows(g, rows):
    if len(rows) == 0:
        return np.zeros(0, np.int64)
    return np.unique(np.concatenate([g.row(int(u))[0] for u in rows]))


def _union_cols(g, cols):
    if len(cols) == 0:
        return np.zeros(0, np.int64)
    return np.unique(np.concatenate([g.col(int(v)) for v in cols]))


def extract_subgraph(g, i, j, h=1, sample_ratio=1.0, max_nodes_per_hop=None,
                     sampler=None, seed=0, pair_id=0):
    """One enclosing subgraph in canonical form.

    Returns dict(u_nodes, v_nodes [global ids, target first], u_dist, v_dist,
    u, v [local ids, v offset by n_u], r [rating labels], node_labels).
    Reference: util_functions.py:208-247.
    """
    i, j = int(i), int(j)
    u_nodes, v_nodes = [i], [j]
    u_dist, v_dist = [0], [0]
    u_vis, v_vis = np.array([i], np.int64), np.array([j], np.int64)
    u_fr, v_fr = np.array([i], np.int64), np.array([j], np.int64)
    for dist in range(1, h + 1):
        # both new fringes come from the *previous* fringes (tuple assignment, ref :217)
        v_new, u_new = _union_rows(g, u_fr), _union_cols(g, v_fr)
        u_new = np.setdiff1d(u_new, u_vis, assume_unique=True)
        v_new = np.setdiff1d(v_new, v_vis, assume_unique=True)
        # visited grows by the UNSAMPLED fringe (ref :220-221 precede sampling)
        u_vis = np.union1d(u_vis, u_new)
        v_vis = np.union1d(v_vis, v_new)
        for side in (0, 1):
            cand = u_new if side == 0 else v_new
            k = len(cand)
            if sample_ratio < 1.0:
                k = int(sample_ratio * len(cand))
            if max_nodes_per_hop is not None and max_nodes_per_hop < k:  # strict '<' (ref :226,:228)
                k = int(max_nodes_per_hop)
            if k < len(cand):
                if sampler is None:
                    cand = hash_sample(cand, k, seed, pair_id, side, dist)
                else:
                    cand = np.sort(np.asarray(sampler(cand, k, side, dist), dtype=np.int64))
                    assert len(cand) == k
            if side == 0:
                u_fr = cand
            else:
                v_fr = cand
        if len(u_fr) == 0 and len(v_fr) == 0:
            break
        u_nodes += u_fr.tolist()
        v_nodes += v_fr.tolist()
        u_dist += [dist] * len(u_fr)
        v_dist += [dist] * len(v_fr)
    n_u = len(u_nodes)
    # induced sub-matrix Arow[u_nodes][:, v_nodes] minus the target edge (ref :236-243)
    lut = np.full(g.shape[1], -1, np.int64)
    lut[np.asarray(v_nodes, np.int64)] = np.arange(len(v_nodes))
    us, vs, rs = [], [], []
    for a, unode in enumerate(u_nodes):
        cols, rat = g.row(unode)
        b = lut[cols]
        keep = b >= 0
        if a == 0:
            keep &= b != 0
        b, rr = b[keep], rat[keep]
        o = np.argsort(b, kind="stable")
        us.append(np.full(len(b), a, np.int64))
        vs.append(b[o])
        rs.append(rr[o])
    u = np.concatenate(us) if us else np.zeros(0, np.int64)
    v = np.concatenate(vs) if vs else np.zeros(0, np.int64)
    r = np.concatenate(rs) if rs else np.zeros(0, np.int64)
    node_labels = np.array([2 * d for d in u_dist] + [2 * d + 1 for d in v_dist], np.int64)  # ref :245
    return dict(u_nodes=np.asarray(u_nodes, np.int64), v_nodes=np.asarray(v_nodes, np.int64),
                u_dist=np.asarray(u_dist, np.int64), v_dist=np.asarray(v_dist, np.int64),
                u=u, v=v + n_u, r=r, node_labels=node_labels)


def construct_graph(sub, y, h=1):
    """(u,v,r,labels) -> PyG-layout arrays.  Reference util_functions.py:280-297."""
    u, v, r = sub["u"], sub["v"], sub["r"]
    edge_index = np.stack([np.concatenate([u, v]), np.concatenate([v, u])], 0).astype(np.int64)
    edge_type = np.concatenate([r, r]).astype(np.int64)
    n = len(sub["node_labels"])
    x = np.zeros((n, 2 * h + 2), np.float32)
    x[np.arange(n), sub["node_labels"]] = 1.0
    return dict(x=x, edge_index=edge_index, edge_type=edge_type, y=np.array([y], np.float32),
                node_labels=sub["node_labels"].astype(np.int64))


def collate(graphs):
    """PyG ``Batch.from_data_list``: concat x / edge_type / y,
    offset-concat edge_index by the running node count, build ``batch``."""
    xs, eis, ets, ys, bs, labs = [], [], [], [], [], []
    off = 0
    for gi, d in enumerate(graphs):
        n = d["x"].shape[0]
        xs.append(d["x"])
        eis.append(d["edge_index"] + off)
        ets.append(d["edge_type"])
        ys.append(d["y"])
        labs.append(d["node_labels"])
        bs.append(np.full(n, gi, np.int64))
        off += n
    return dict(x=np.concatenate(xs, 0), edge_index=np.concatenate(eis, 1),
                edge_type=np.concatenate(ets), y=np.concatenate(ys), batch=np.concatenate(bs),
                node_labels=np.concatenate(labs), num_graphs=len(graphs))


def extract_batch(g, pair_u, pair_v, labels, class_values, h=1, sample_ratio=1.0,
                  max_nodes_per_hop=None, sampler=None, seed=0, pair_ids=None):
    """MyDynamicDataset.get for every pair (ref :138-145) + collate; also returns the
    per-graph node lists so tests can compare them."""
    graphs, subs = [], []
    for k in range(len(pair_u)):
        pid = int(pair_ids[k]) if pair_ids is not None else k
        smp = None
        if sampler is not None:
            smp = (lambda c, kk, side, hop, _k=k: sampler(_k, c, kk, side, hop))
        sub = extract_subgraph(g, pair_u[k], pair_v[k], h, sample_ratio, max_nodes_per_hop,
                               smp, seed, pid)
        subs.append(sub)
        graphs.append(construct_graph(sub, class_values[int(labels[k])], h))
    out = collate(graphs)
    out["subs"] = subs
    return out


def canonicalize_reference(u, v, r, node_labels, u_nodes_ref, v_nodes_ref):
    """Relabel a *reference* subgraph (node order = CPython set order) into canonical
    form.  ``u_nodes_ref``/``v_nodes_ref`` are the global ids in the reference's order
    (target first).  Nodes are re-ordered by (distance, global id) within each side."""
    u, v, r = np.asarray(u, np.int64), np.asarray(v, np.int64), np.asarray(r, np.int64)
    n_u, n_v = len(u_nodes_ref), len(v_nodes_ref)
    labels = np.asarray(node_labels, np.int64)
    ud, vd = labels[:n_u] // 2, (labels[n_u:] - 1) // 2
    uo = np.lexsort((np.asarray(u_nodes_ref), ud))   # new position -> old index
    vo = np.lexsort((np.asarray(v_nodes_ref), vd))
    assert uo[0] == 0 and vo[0] == 0
    u_new_of_old = np.empty(n_u, np.int64); u_new_of_old[uo] = np.arange(n_u)
    v_new_of_old = np.empty(n_v, np.int64); v_new_of_old[vo] = np.arange(n_v)
    uu = u_new_of_old[u]
    vv = v_new_of_old[v - n_u]
    o = np.lexsort((vv, uu))
    return dict(u_nodes=np.asarray(u_nodes_ref, np.int64)[uo], v_nodes=np.asarray(v_nodes_ref, np.int64)[vo],
                u=uu[o], v=vv[o] + n_u, r=r[o],
                node_labels=np.concatenate([labels[:n_u][uo], labels[n_u:][vo]]))
