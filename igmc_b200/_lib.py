"""ctypes binding of libigmc_b200.so (the C-ABI declared in include/igmc_b200.h).

There is NO fallback: if the shared library is missing or an entry point is absent the import of
the hot path fails loudly.  Build it with ``python -m igmc_b200.build`` (or
``__graft_entry__.build()``).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libigmc_b200.so")

MAX_LAYERS = 8
REDUCE_WS_FLOATS = MAX_LAYERS * 32 * 64 + MAX_LAYERS * 32 + 64   # IGMC_REDUCE_WS_FLOATS
MAX_HOP = 3
HIDDEN = 32
LIN1_OUT = 128

ERR_NAMES = {
    0: "ok",
    1: "a subgraph side exceeded the node-list capacity (cap)",
    2: "batch edge capacity exceeded",
    3: "batch node capacity exceeded",
    4: "subgraph larger than the shared-memory plan of the model kernels (n_cap)",
    5: "malformed batch (edges not grouped by graph / node outside its graph / no target node)",
}

p_i32 = C.c_void_p  # all device pointers travel as raw addresses
vp = C.c_void_p


class CSR(C.Structure):
    _fields_ = [("row_ptr", vp), ("col_idx", vp), ("rating", vp), ("col_ptr", vp), ("row_idx", vp),
                ("num_users", C.c_int32), ("num_items", C.c_int32)]


class Pairs(C.Structure):
    _fields_ = [("idx", vp), ("links_u", vp), ("links_v", vp), ("links_label", vp), ("pair_id", vp)]


class ExtractWS(C.Structure):
    _fields_ = [("nodes_u", vp), ("nodes_v", vp), ("n_u", vp), ("n_v", vp), ("row_cnt", vp), ("m_cnt", vp),
                ("col_cnt", vp), ("hop_off", vp), ("sync", vp)]


class BatchOut(C.Structure):
    _fields_ = [("node_cap", C.c_int32), ("edge_cap", C.c_int32), ("feat_dim", C.c_int32),
                ("x", vp), ("node_label", vp), ("batch", vp), ("node_gid", vp), ("edge_index", vp),
                ("edge_type", vp), ("y", vp), ("node_ptr", vp), ("edge_ptr", vp), ("graph_nu", vp),
                ("counts", vp), ("adj_in_ptr", vp), ("adj_in", vp), ("adj_eid", vp), ("adj_tmp", vp)]


class Store(C.Structure):
    _fields_ = [(k, vp) for k in ("node_off", "edge_off", "node_label", "node_gid", "edge_src", "edge_dst",
                                  "edge_type", "y", "graph_nu", "adj_ptr", "adj_in", "adj_eid")]


class Adj(C.Structure):
    _fields_ = [("in_ptr", vp), ("in_adj", vp), ("in_eid", vp), ("out_ptr", vp), ("out_adj", vp),
                ("out_eid", vp), ("tmp", vp), ("symmetric", C.c_int32)]


class Model(C.Structure):
    _fields_ = [("num_layers", C.c_int32), ("num_relations", C.c_int32), ("num_bases", C.c_int32),
                ("in_dim0", C.c_int32),
                ("off_att", C.c_int32 * MAX_LAYERS), ("off_basis", C.c_int32 * MAX_LAYERS),
                ("off_root", C.c_int32 * MAX_LAYERS), ("off_bias", C.c_int32 * MAX_LAYERS),
                ("off_lin1_w", C.c_int32), ("off_lin1_b", C.c_int32), ("off_lin2_w", C.c_int32),
                ("off_lin2_b", C.c_int32), ("conv_param_count", C.c_int32), ("param_count", C.c_int32),
                ("multiply_by", C.c_float), ("readout", C.c_int32), ("list_hint", C.c_int32)]


class Dropout(C.Structure):
    _fields_ = [("adj_dropout", C.c_float), ("hidden_dropout", C.c_float), ("seed", C.c_uint64),
                ("seed_dev", vp), ("edge_keep", vp), ("hidden_keep", vp)]


class Saved(C.Structure):
    _fields_ = [("states", vp), ("zsave", vp), ("inv_deg", vp), ("feat", vp), ("hid", vp),
                ("hid_gscale", vp), ("pred", vp), ("target", vp), ("node_cap", C.c_int32), ("dstate", vp), ("wprep", vp), ("prof", vp),
                ("gate", vp)]


class Stage(C.Structure):
    _fields_ = [("tab", vp), ("ent", vp), ("inv_deg", vp), ("tab_ints", C.c_int32), ("lcap", C.c_int32),
                ("chunk", C.c_int32), ("cluster", C.c_int32)]


MAX_RANKS = 8


class Comm(C.Structure):
    _fields_ = [("grad", vp * MAX_RANKS), ("flag", vp * MAX_RANKS), ("state", vp), ("world", C.c_int32),
                ("rank", C.c_int32), ("stride", C.c_int32), ("pad_", C.c_int32)]


class SortPool(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("k", "width", "state_stride", "c1", "c2", "kw2", "t1", "t2", "dense_dim",
                                         "off_conv1_w", "off_conv1_b", "off_conv2_w", "off_conv2_b", "off_lin1_w",
                                         "off_lin1_b", "off_lin2_w", "off_lin2_b", "param_begin", "param_end")]


class SortPoolSaved(C.Structure):
    _fields_ = [(k, vp) for k in ("rank", "perm", "act1", "pool", "flat", "hid", "hid_gscale", "pred", "dhid",
                                  "gpart")]


_SIGS = {
    "igmc_extract_batch": [C.POINTER(CSR), C.POINTER(Pairs), C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64, vp, C.c_int,
                           vp, vp, vp, vp, C.POINTER(ExtractWS), vp, C.c_int, C.c_int, C.POINTER(BatchOut), vp, vp],
    "igmc_assemble_batch": [C.POINTER(Store), vp, C.c_int, C.POINTER(BatchOut), vp, vp],
    "igmc_batch_ptrs": [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp],
    "igmc_batch_prepare": [vp, C.c_int64, vp, vp, vp, C.c_int, C.c_int, C.POINTER(Adj), vp, vp],
    "igmc_forward": [C.POINTER(Model), vp, vp, vp, vp, C.POINTER(Adj), C.c_int, C.c_int, C.POINTER(Dropout),
                     C.c_int, C.POINTER(Saved), vp, C.c_float, vp, vp, C.c_int, C.POINTER(Stage), vp, vp],
    "igmc_backward": [C.POINTER(Model), vp, vp, vp, vp, C.POINTER(Adj), C.c_int, C.c_int, C.POINTER(Dropout),
                      C.POINTER(Saved), vp, vp, vp, C.c_int, C.POINTER(Stage), vp, vp],
    "igmc_forward_backward": [C.POINTER(Model), vp, vp, vp, vp, C.POINTER(Adj), C.c_int, C.c_int, C.POINTER(Dropout),
                              C.POINTER(Saved), vp, C.c_float, vp, vp, vp, vp, C.c_int, C.POINTER(Stage),
                              C.POINTER(Stage), vp, vp],
    "igmc_raw_grad_count": [C.POINTER(Model)],
    "igmc_stage_plan": [C.POINTER(Model), C.c_int, C.c_int, C.c_int, C.POINTER(Stage)],
    "igmc_stage_lists": [C.POINTER(Model), vp, vp, C.POINTER(Adj), C.c_int, C.c_int, C.POINTER(Dropout), C.c_int,
                         C.POINTER(Stage), C.POINTER(Stage), vp, vp],
    "igmc_grad_reduce": [C.POINTER(Model), vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, C.c_float, C.c_float,
                         C.c_float, vp, vp, vp, vp],
    "igmc_adam_step": [vp, vp, vp, vp, vp, C.c_int, C.c_float, vp, C.c_float, C.c_float, C.c_float, C.c_float,
                       C.c_float, vp, vp, C.c_float, vp],
    "igmc_reduce_update": [C.POINTER(Model), vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_float, C.c_float,
                           C.POINTER(Comm), vp, vp, vp, C.c_float, vp, C.c_float, C.c_float, C.c_float, C.c_float,
                           C.c_float, vp, vp, C.c_float, vp, vp, vp, C.c_int, vp, vp],
    "igmc_comm_alloc": [C.c_int64, C.POINTER(C.c_void_p), C.c_char_p],
    "igmc_comm_open": [C.c_char_p, C.POINTER(C.c_void_p)],
    "igmc_comm_close": [vp],
    "igmc_comm_free": [vp],
    "igmc_prep_weights": [C.POINTER(Model), vp, vp, vp],
    "igmc_gate_wait": [vp, C.c_int, C.c_int, vp],
    "igmc_build_info": [],
    "igmc_model_plan": [C.POINTER(Model), C.c_int, C.c_int, C.c_int],
    "igmc_sortpool_plan": [C.POINTER(SortPool), C.c_int, C.c_int],
    "igmc_sortpool_forward": [C.POINTER(SortPool), vp, vp, vp, C.c_int, C.c_int, C.POINTER(Dropout), C.c_int,
                              C.POINTER(SortPoolSaved), vp, C.c_float, vp, vp, vp, vp],
    "igmc_sortpool_backward": [C.POINTER(SortPool), vp, vp, vp, C.c_int, C.c_int, C.POINTER(SortPoolSaved), vp, vp,
                               C.c_float, vp, vp, vp],
}

EXPORTS = tuple(_SIGS)
_lib = None


def load():
    """dlopen the library once; raises if it (or a declared symbol) is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "igmc_b200: %s not found - build the CUDA extension first (python -m igmc_b200.build). "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = C.c_int
    if lib.igmc_build_info() != 100:
        raise RuntimeError("igmc_b200: library was not built for sm_100a")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("igmc_b200: %s failed with code %d" % (what, rc))


def ptr(t):
    """device address of a torch tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()
