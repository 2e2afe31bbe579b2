#!/usr/bin/env python
"""Builds tests/golden/flixster_ratings.npz from the Monti et al. flixster split that ships with the reference
(raw_data/flixster/training_test_dataset.mat, MATLAB v7.3 = HDF5) and, with --vectors, the golden extraction vectors
tests/golden/flixster_cases.npz produced by the reference's OWN subgraph_extraction_labeling on a sample of real pairs.

    python tests/golden/make_flixster_fixture.py [--reference /root/reference] [--vectors]

h5py is not available offline, so this file carries a minimal reader for exactly what the .mat needs: superblock v0,
old-style groups (symbol-table B-tree + local heap), version-1 object headers, contiguous or chunked (B-tree v1)
dense datasets of IEEE doubles with the deflate / shuffle filters.  The split logic restates the reference's
load_data_monti (preprocessing.py:203-330) in testing mode (train = train + val, Main.py --testing).
"""
import argparse
import os
import struct
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5(object):
    def __init__(self, path):
        self.b = open(path, "rb").read()
        off = 0
        while self.b[off:off + 8] != b"\x89HDF\r\n\x1a\n":       # the superblock sits behind MATLAB's 512-byte user block
            off = 512 if off == 0 else off * 2
        sb = off + 8
        ver = self.b[sb]
        assert ver == 0, "superblock version %d not handled" % ver
        assert self.b[sb + 5] == 8 and self.b[sb + 6] == 8     # sizes of offsets / lengths
        self.base = struct.unpack_from("<Q", self.b, sb + 16)[0]
        # root group symbol-table entry: link name offset, object header address, cache type, reserved, scratch pad
        ent = sb + 16 + 32
        _, ohdr, cache = struct.unpack_from("<QQI", self.b, ent)
        assert cache == 1
        self.root_btree, self.root_heap = struct.unpack_from("<QQ", self.b, ent + 24)
        self.root = self._group_entries(self.root_btree, self.root_heap)

    def a(self, addr):
        return self.base + addr

    # ---- old-style group: B-tree of symbol nodes + local heap of names ----
    def _heap_data(self, heap_addr):
        p = self.a(heap_addr)
        assert self.b[p:p + 4] == b"HEAP"
        size, _, data_addr = struct.unpack_from("<QQQ", self.b, p + 8)
        return self.a(data_addr)

    def _group_entries(self, btree_addr, heap_addr):
        names = {}
        heap = self._heap_data(heap_addr)

        def walk(addr):
            p = self.a(addr)
            if self.b[p:p + 4] == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", self.b, p + 4)
                assert ntype == 0
                q = p + 8 + 16                      # signature/type/level/used + two sibling pointers
                for k in range(used):
                    q += 8                          # key k (heap offset)
                    child = struct.unpack_from("<Q", self.b, q)[0]
                    q += 8
                    walk(child)
            else:
                assert self.b[p:p + 4] == b"SNOD", self.b[p:p + 4]
                nsym = struct.unpack_from("<H", self.b, p + 6)[0]
                q = p + 8
                for k in range(nsym):
                    name_off, ohdr = struct.unpack_from("<QQ", self.b, q)
                    end = self.b.index(b"\x00", heap + name_off)
                    names[self.b[heap + name_off:end].decode()] = ohdr
                    q += 40
        walk(btree_addr)
        return names

    # ---- version-1 object header ----
    def _messages(self, ohdr_addr):
        p = self.a(ohdr_addr)
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", self.b, p)
        assert ver == 1
        blocks = [(p + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            q, size = blocks.pop(0)
            end = q + size
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", self.b, q)
                body = q + 8
                if mtype == 0x10:                   # continuation
                    coff, clen = struct.unpack_from("<QQ", self.b, body)
                    blocks.append((self.a(coff), clen))
                out.append((mtype, body, msize))
                q = body + msize
        return out

    def dataset(self, name):
        """dense dataset as a numpy array in HDF5 (row-major) dimension order"""
        shape = dtype = layout = None
        filters = []
        for mtype, p, size in self._messages(self.root[name]):
            if mtype == 0x1:                         # dataspace
                ver, rank, flags = struct.unpack_from("<BBB", self.b, p)
                q = p + (8 if ver == 1 else 4)
                shape = struct.unpack_from("<%dQ" % rank, self.b, q)
            elif mtype == 0x3:                       # datatype
                cls_ver, b0, b1, b2, tsize = struct.unpack_from("<BBBBI", self.b, p)
                assert (cls_ver & 0x0F) == 1 and tsize == 8 and (b0 & 1) == 0, "expected little-endian float64"
                dtype = np.dtype("<f8")
            elif mtype == 0xB:                       # filter pipeline
                ver, nf = struct.unpack_from("<BB", self.b, p)
                assert ver == 1
                q = p + 8
                for _ in range(nf):
                    fid, nlen, fflags, ncd = struct.unpack_from("<HHHH", self.b, q)
                    q += 8 + ((nlen + 7) & ~7) + 4 * ncd
                    if ncd % 2:
                        q += 4
                    filters.append(fid)
            elif mtype == 0x8:                       # data layout
                ver, cls = struct.unpack_from("<BB", self.b, p)
                assert ver == 3
                if cls == 1:
                    addr, sz = struct.unpack_from("<QQ", self.b, p + 2)
                    layout = ("contiguous", addr, sz)
                elif cls == 2:
                    nd = self.b[p + 2]
                    bt = struct.unpack_from("<Q", self.b, p + 3)[0]
                    cdims = struct.unpack_from("<%dI" % nd, self.b, p + 11)
                    layout = ("chunked", bt, cdims)
                else:
                    raise NotImplementedError("compact layout")
        assert shape is not None and dtype is not None and layout is not None, name
        if layout[0] == "contiguous":
            p = self.a(layout[1])
            return np.frombuffer(self.b, dtype, int(np.prod(shape)), p).reshape(shape).copy()
        out = np.zeros(shape, dtype)
        cdims = layout[2][:-1]
        rank = len(shape)

        def walk(addr):
            p = self.a(addr)
            assert self.b[p:p + 4] == b"TREE"
            ntype, level, used = struct.unpack_from("<BBH", self.b, p + 4)
            assert ntype == 1
            q = p + 24
            for k in range(used):
                csize, fmask = struct.unpack_from("<II", self.b, q)
                offs = struct.unpack_from("<%dQ" % (rank + 1), self.b, q + 8)
                q += 8 + 8 * (rank + 1)
                child = struct.unpack_from("<Q", self.b, q)[0]
                q += 8
                if level > 0:
                    walk(child)
                    continue
                raw = self.b[self.a(child):self.a(child) + csize]
                for i, fid in reversed(list(enumerate(filters))):
                    if fmask & (1 << i):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:                    # shuffle: byte planes -> elements
                        n = len(raw) // dtype.itemsize
                        raw = np.frombuffer(raw, np.uint8).reshape(dtype.itemsize, n).T.tobytes()
                    else:
                        raise NotImplementedError("filter %d" % fid)
                chunk = np.frombuffer(raw, dtype).reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        walk(layout[1])
        return out


def load_matlab_dense(path, name):
    """reference preprocessing.py:32-52 for a dense field: float32, transposed (MATLAB is column-major)"""
    return H5(path).dataset(name).astype(np.float32).T


def flixster_split(path, testing=True):
    """load_data_monti (preprocessing.py:203-330) for flixster: returns dict of train / test (u, v, label) arrays,
    class_values, shape.  testing=True merges the 20 % validation part back into train (Main.py --testing)."""
    M = load_matlab_dense(path, "M")
    Otr = load_matlab_dense(path, "Otraining")
    Ote = load_matlab_dense(path, "Otest")
    num_users, num_items = M.shape
    ratings = np.sort(np.unique(M[np.where(M)])).tolist()
    rating_dict = {r: i for i, r in enumerate(ratings)}
    labels = np.full((num_users, num_items), -1, np.int32)
    uu, vv = np.where(M)
    labels[uu, vv] = np.array([rating_dict[r] for r in M[uu, vv]])
    tr = np.stack(np.where(Otr), 1)
    te = np.stack(np.where(Ote), 1)
    num_train = len(tr)
    num_val = int(np.ceil(num_train * 0.2))
    rand_idx = list(range(len(tr)))
    np.random.seed(42)
    np.random.shuffle(rand_idx)                       # preprocessing.py:287-289
    tr = tr[rand_idx]
    val, train = tr[:num_val], tr[num_val:]
    if testing:                                       # preprocessing.py:320-324: train = val + train
        train = np.concatenate([val, train], 0)
    out = dict(num_users=num_users, num_items=num_items,
               class_values=np.asarray(ratings, np.float32),
               train_u=train[:, 0].astype(np.int32), train_v=train[:, 1].astype(np.int32),
               test_u=te[:, 0].astype(np.int32), test_v=te[:, 1].astype(np.int32))
    out["train_l"] = labels[out["train_u"], out["train_v"]].astype(np.int32)
    out["test_l"] = labels[out["test_u"], out["test_v"]].astype(np.int32)
    assert (out["train_l"] >= 0).all() and (out["test_l"] >= 0).all()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--vectors", action="store_true", help="also write flixster_cases.npz with the reference's own outputs")
    args = ap.parse_args()
    path = os.path.join(args.reference, "raw_data", "flixster", "training_test_dataset.mat")
    d = flixster_split(path)
    print("flixster: %d x %d, train %d, test %d, classes %s" % (d["num_users"], d["num_items"], len(d["train_u"]),
                                                              len(d["test_u"]), d["class_values"].tolist()))
    np.savez_compressed(os.path.join(HERE, "flixster_ratings.npz"), **d)
    if not args.vectors:
        return
    sys.path.insert(0, ROOT)
    from igmc_b200.data import build_adj
    from oracle import ref_shim
    A = build_adj(d["train_u"], d["train_v"], d["train_l"], d["num_users"], d["num_items"])
    rng = np.random.default_rng(7)
    pick_tr = rng.choice(len(d["train_u"]), 40, replace=False)
    pick_te = rng.choice(len(d["test_u"]), 24, replace=False)
    pu = np.concatenate([d["train_u"][pick_tr], d["test_u"][pick_te]])
    pv = np.concatenate([d["train_v"][pick_tr], d["test_v"][pick_te]])
    pl = np.concatenate([d["train_l"][pick_tr], d["test_l"][pick_te]])
    keys = ("u_nodes", "v_nodes", "u", "v", "r", "node_labels")
    acc = {k: [] for k in keys}
    ys = []
    indexers = ref_shim.make_indexers(A)
    for i, j, l in zip(pu, pv, pl):
        c, _ = ref_shim.extract_ref_canonical(A, int(i), int(j), int(l), d["class_values"], 1, 1.0, 10000, indexers)
        for k in keys:
            acc[k].append(np.asarray(c[k], np.int64))
        ys.append(float(c["y"]))
    out = dict(pairs=np.stack([pu, pv, pl]).astype(np.int64), y=np.asarray(ys, np.float64))
    for k in keys:
        out[k] = np.concatenate(acc[k])
        out[k + "_off"] = np.concatenate([[0], np.cumsum([len(x) for x in acc[k]])]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "flixster_cases.npz"), **out)
    print("wrote %d reference-produced cases" % len(pu))


if __name__ == "__main__":
    main()
