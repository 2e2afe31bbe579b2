"""Synthetic MovieLens-shaped rating matrices and dataset presets.

The real ml_100k / ml_1m files are downloaded at run time by the reference
(data_utils.py:65-85) and are not available offline, so every MovieLens config
runs on a seeded synthetic matrix with matching shape, nnz and rating histogram.
The *output contract* is the reference's: ``adj_train`` is a scipy CSR float32
whose stored value is ``rating_label + 1`` (preprocessing.py:190-197) plus
``(u_indices, v_indices, labels)`` triples and the sorted ``class_values``.
"""
import os

import numpy as np
import scipy.sparse as ssp

ML_RATING_PROBS = (0.056, 0.107, 0.261, 0.349, 0.227)


def synth_ratings(num_users, num_items, nnz, num_classes=5, seed=0, probs=ML_RATING_PROBS):
    """Draw ``nnz`` unique (u, v) pairs with log-normal user activity / item popularity.

    Returns (u[int64], v[int64], label[int64]) in a seeded random order.
    """
    rng = np.random.default_rng(seed)
    a = rng.lognormal(0.0, 1.0, num_users)
    p = rng.lognormal(0.0, 1.4, num_items)
    a /= a.sum()
    p /= p.sum()
    have = np.zeros(0, np.int64)
    while len(have) < nnz:
        need = int((nnz - len(have)) * 1.3) + 1024
        u = rng.choice(num_users, need, p=a)
        v = rng.choice(num_items, need, p=p)
        have = np.unique(np.concatenate([have, u.astype(np.int64) * num_items + v]))
    have = rng.permutation(have)[:nnz]
    u, v = have // num_items, have % num_items
    if probs is None or len(probs) != num_classes:
        probs = np.full(num_classes, 1.0 / num_classes)
    labels = rng.choice(num_classes, nnz, p=np.asarray(probs) / np.sum(probs)).astype(np.int64)
    return u, v, labels


def build_adj(u, v, labels, num_users, num_items):
    """CSR float32 with value = label + 1 (reference preprocessing.py:194-197)."""
    return ssp.csr_matrix((labels.astype(np.float32) + 1.0, (u, v)), shape=(num_users, num_items),
                          dtype=np.float32)


PRESETS = {
    # name: (num_users, num_items, nnz_train, num_classes, max_nodes_per_hop, adj_dropout)
    "ml_100k": (943, 1682, 80000, 5, 200, 0.2),      # u1.base size, testing mode
    "ml_1m": (6040, 3706, 900188, 5, 100, 0.0),      # 1,000,209 - ceil(10%) (C4)
    "ml_1m_r02": (6040, 3706, 216045, 5, 100, 0.0),  # ratio 0.2 (C5)
    "flixster": (3000, 3000, 23556, 10, 10000, 0.2),  # Monti split sizes, 10 rating levels, no cap
    "tiny": (60, 40, 600, 5, 10, 0.2),
}


FLIXSTER_FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                "flixster_ratings.npz")


def load_flixster():
    """The REAL Monti et al. flixster split (3000 x 3000, 23 556 train / 2 617 test ratings, 10 rating levels
    0.5..5), extracted once from the reference's raw_data/flixster/training_test_dataset.mat by
    tests/golden/make_flixster_fixture.py (split logic of preprocessing.py:203-330, testing mode).  Same dict layout
    as ``make_synthetic_dataset``; max_nodes_per_hop 10000 never samples (max degree 153)."""
    z = np.load(FLIXSTER_FIXTURE)
    nu, nv = int(z["num_users"]), int(z["num_items"])
    tu, tv, tl = z["train_u"], z["train_v"], z["train_l"]
    return dict(name="flixster", adj_train=build_adj(tu, tv, tl, nu, nv), train=(tu, tv, tl),
                test=(z["test_u"], z["test_v"], z["test_l"]), class_values=z["class_values"].astype(np.float32),
                num_users=nu, num_items=nv, max_nodes_per_hop=10000, adj_dropout=0.2,
                num_relations=len(z["class_values"]), real=True)


def make_synthetic_dataset(name="ml_1m", seed=0, num_test=2000):
    """Returns dict(adj_train, train=(u,v,labels), test=(u,v,labels), class_values, ...).

    Train pairs are the nonzeros of adj_train (the reference trains on pairs that ARE
    in the matrix and removes the target edge per subgraph, util_functions.py:238);
    test pairs are additional pairs NOT in adj_train.
    """
    if name == "flixster" and os.path.exists(FLIXSTER_FIXTURE):
        return load_flixster()          # the one BASELINE config whose real data is small enough to ship as a fixture
    nu, nv, nnz, R, mnph, adj_dropout = PRESETS[name]
    num_test = max(0, min(num_test, (nu * nv - nnz) // 2))   # unique pairs must exist (tiny presets)
    u, v, lab = synth_ratings(nu, nv, nnz + num_test, R, seed)
    tu, tv, tl = u[:nnz], v[:nnz], lab[:nnz]
    adj = build_adj(tu, tv, tl, nu, nv)
    perm = np.random.default_rng(seed + 1).permutation(nnz)
    return dict(name=name, adj_train=adj, train=(tu[perm], tv[perm], tl[perm]),
                test=(u[nnz:], v[nnz:], lab[nnz:]),
                class_values=np.arange(1, R + 1, dtype=np.float32),
                num_users=nu, num_items=nv, max_nodes_per_hop=mnph, adj_dropout=adj_dropout,
                num_relations=R)
