"""Graph build: the host-side mirror of the reference's builders over the device kernels.

Reference surface kept (same names, argument meaning, return shapes):
    edges_to_user_csr(edges_2xE, num_users)                         lightgcn_cu.py:259-276
    build_cred_weighted_mats(train_edges, U, I, cred_u, device)     lightgcn_cu.py:368-399
    build_message_passing_mats(train_edges, U, I, cred_u, device)   Version-2/lighgcn_cu_pop.py:429-452
    build_message_passing_mats_degree_aware(...)                    version_1/lightgcn_cu_pop_Degree-Aware Message.py:349-403

The reference returns two torch.sparse_coo tensors.  Here the two "matrices" are views
(`Operator`) of one `CredGraph` living in HBM: one coalesced sparsity pattern held in both row
orders with two value arrays each -- what the forward and the adjoint propagation need without
ever transposing at run time.  `Operator` still answers `.indices()/.values()/.coalesce()/.shape`
so that parity checks read like the reference's own code.

HBM layout of a CredGraph (E = train edges incl. duplicates, nnz = distinct (u, i) pairs):
    deg_u  int32[U], deg_i int32[I]
    samp_indptr int64[U+1], samp_idx int32[E]        duplicate-keeping user rows (sampler, eval mask)
    by_user: indptr int64[U+1], idx int32[nnz], val_fwd = A, val_bwd = C^T     float32[nnz]
    by_item: indptr int64[I+1], idx int32[nnz], val_fwd = C, val_bwd = A^T     float32[nnz]
    work schedule per order: perm int32[n_rows] (degree-descending), chunk tables of rows > 256 nnz
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import CsrStruct, check, lib, ptr, stream_ptr, workspace


def _as_device_edges(edges_2xE, device):
    if isinstance(edges_2xE, torch.Tensor):
        e = edges_2xE.to(device=device, dtype=torch.int32)
    else:
        a = np.asarray(edges_2xE)
        if a.ndim != 2 or a.shape[0] != 2:
            raise ValueError(f"edges must have shape [2, E], got {a.shape}")
        e = torch.from_numpy(np.ascontiguousarray(a.astype(np.int32, copy=False))).to(device)
    return e[0].contiguous(), e[1].contiguous()


def _cuda_device(device) -> torch.device:
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise _lib.CgxError(f"credgcn runs on CUDA devices only (got device={device!r}); there is no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


@dataclass
class Csr:
    """One row order of the coalesced operator pattern (struct cgx_csr + the tensors it points at)."""
    n_rows: int
    n_cols: int
    nnz: int
    indptr: torch.Tensor
    idx: torch.Tensor
    val_fwd: torch.Tensor
    val_bwd: torch.Tensor
    perm: torch.Tensor | None = None
    chunk_ptr: torch.Tensor | None = None
    chunk_row: torch.Tensor | None = None
    n_long: int = 0
    n_chunks: int = 0
    n_huge: int = 0
    arrive: torch.Tensor | None = None
    work: torch.Tensor | None = None
    idx_hint: torch.Tensor | None = None       # idx | hot << 31 (cgx_hot_hints); None = no hints
    n_hot: int = 0
    _struct: CsrStruct | None = field(default=None, repr=False)

    def struct(self) -> CsrStruct:
        if self._struct is None:
            s = CsrStruct()
            s.n_rows, s.n_cols, s.nnz = self.n_rows, self.n_cols, self.nnz
            s.indptr, s.idx = self.indptr.data_ptr(), self.idx.data_ptr()
            s.val_fwd, s.val_bwd = self.val_fwd.data_ptr(), self.val_bwd.data_ptr()
            s.n_long, s.n_chunks = self.n_long, self.n_chunks
            s.perm = self.perm.data_ptr()
            s.chunk_ptr = self.chunk_ptr.data_ptr() if self.n_long else None
            s.chunk_row = self.chunk_row.data_ptr() if self.n_long else None
            s.n_huge = self.n_huge
            s.arrive = self.arrive.data_ptr() if self.n_long else None
            s.work = self.work.data_ptr() if self.work is not None else None
            s.idx_hint = self.idx_hint.data_ptr() if self.idx_hint is not None else None
            self._struct = s
        return self._struct

    def ref(self):
        return C.byref(self.struct())

    def schedule(self):
        """Degree-descending row order + chunk tables of the long rows (cgx_row_schedule)."""
        dev = self.indptr.device
        ws = workspace(lib().cgx_row_schedule_workspace_bytes(self.n_rows), dev)
        n_long, n_chunks, n_huge = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self.perm = torch.empty(self.n_rows, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib().cgx_row_schedule(ptr(self.indptr), self.n_rows, ptr(self.perm), C.byref(n_long),
                                         C.byref(n_chunks), C.byref(n_huge), ptr(ws), ws.numel(), stream_ptr(dev)))
            self.n_long, self.n_chunks, self.n_huge = int(n_long.value), int(n_chunks.value), int(n_huge.value)
            if self.n_long:
                self.arrive = torch.zeros(self.n_long, dtype=torch.int32, device=dev)
                self.chunk_ptr = torch.empty(self.n_long + 1, dtype=torch.int32, device=dev)
                self.chunk_row = torch.empty(self.n_chunks, dtype=torch.int32, device=dev)
                check(lib().cgx_row_schedule_chunks(self.n_rows, self.n_long, self.n_chunks, ptr(self.chunk_ptr),
                                                    ptr(self.chunk_row), ptr(ws), ws.numel(), stream_ptr(dev)))
            n_items = self.n_chunks + self.n_rows - self.n_long
            self.work = torch.empty(max(n_items, 1), 4, dtype=torch.int32, device=dev)
            check(lib().cgx_row_schedule_work(ptr(self.indptr), ptr(self.perm), self.n_rows, self.n_long,
                                              self.n_chunks, ptr(self.chunk_ptr), ptr(self.chunk_row),
                                              ptr(self.work), stream_ptr(dev)))
        self._struct = None

    def set_hot_columns(self, col_by_degree: torch.Tensor, n_hot: int):
        """Mark the n_hot highest-degree columns as hot rows of the gathered table (cgx_hot_hints).  The hint array
        keeps its address when it is rebuilt (a captured CUDA graph stays valid); n_hot = 0 removes the hints."""
        n_hot = max(0, min(int(n_hot), self.n_cols))
        if n_hot == self.n_hot:
            return
        self.n_hot = n_hot
        if n_hot == 0 or self.nnz == 0:
            self.idx_hint, self.n_hot = None, 0
        else:
            dev = self.indptr.device
            if self.idx_hint is None:
                self.idx_hint = torch.empty_like(self.idx)
            ws = workspace(lib().cgx_hot_hints_workspace_bytes(self.n_cols), dev)
            with torch.cuda.device(dev):
                check(lib().cgx_hot_hints(ptr(self.idx), self.nnz, self.n_cols, ptr(col_by_degree), n_hot,
                                          ptr(self.idx_hint), ptr(ws), ws.numel(), stream_ptr(dev)))
        self._struct = None

    def row_ids(self) -> torch.Tensor:
        counts = (self.indptr[1:] - self.indptr[:-1])
        return torch.repeat_interleave(torch.arange(self.n_rows, device=self.indptr.device), counts)


class CredGraph:
    """Device-resident graph: degrees, sampling CSR and both operator orders."""

    def __init__(self, num_users, num_items, variant, device):
        self.num_users, self.num_items, self.variant, self.device = int(num_users), int(num_items), variant, device
        self.num_edges = 0
        self.nnz = 0
        self.deg_u = self.deg_i = None
        self.samp_indptr = self.samp_idx = None
        self.by_user: Csr | None = None
        self.by_item: Csr | None = None
        self._ws_cache: dict = {}

    # operator views in the reference's vocabulary
    def operator(self, which: str) -> "Operator":
        return Operator(self, which)

    def deg_i_float(self) -> np.ndarray:
        """deg_i as the reference returns it: NumPy float32 (lightgcn_cu.py:384, 399)."""
        return self.deg_i.cpu().numpy().astype(np.float32)

    def user_csr_numpy(self):
        return self.samp_indptr.cpu().numpy(), self.samp_idx.cpu().numpy().astype(np.int64)

    # Hot rows (north_star 2: "staging of hot (high-degree) rows"): when a gathered table is larger than L2, the rows of
    # its highest-degree columns -- HOT_BYTES worth of them -- are loaded with the L2 evict_last priority by the SpMM
    # and everything else streams through evict_first.  Half of the 126 MB L2 by default: the two L2 partitions each
    # keep their own copy of a line that SMs of both dies read.
    HOT_BYTES = 48 << 20

    def set_emb_dim(self, d: int, hot_bytes: int | None = None):
        """(Re)build the hot-row hints for embedding width d.  No hints when the table fits L2 anyway."""
        hot_bytes = self.HOT_BYTES if hot_bytes is None else int(hot_bytes)
        l2 = _lib.get_option("L2_TABLE_BYTES")
        for csr, other in ((self.by_user, self.by_item), (self.by_item, self.by_user)):
            beyond = csr.n_cols * d * 4 > l2
            csr.set_hot_columns(other.perm, hot_bytes // (4 * d) if beyond else 0)

    def propagate_workspace(self, d: int, order=None) -> torch.Tensor:
        """Scratch of cgx_propagate_fwd / _bwd.  order ("gs" / "jacobi" or the C enum): the Gauss-Seidel order needs
        half the layer buffers; None sizes it for either order."""
        o = _lib.ORDERS.get(order, order)
        key = ("prop", d, o)
        if key not in self._ws_cache:
            self.set_emb_dim(d)
            if o is None:
                n = lib().cgx_propagate_workspace_bytes(self.by_user.ref(), self.by_item.ref(), d)
            else:
                n = lib().cgx_propagate_workspace_bytes_for(self.by_user.ref(), self.by_item.ref(), d, int(o))
            self._ws_cache[key] = workspace(n, self.device)
        return self._ws_cache[key]


class Operator:
    """Quacks like the coalesced torch.sparse_coo tensor the reference builders return."""

    def __init__(self, graph: CredGraph, which: str):
        assert which in ("A", "C")
        self.graph, self.which = graph, which

    @property
    def shape(self):
        g = self.graph
        return (g.num_users, g.num_items) if self.which == "A" else (g.num_items, g.num_users)

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def coalesce(self):
        return self

    def is_coalesced(self):
        return True

    def _csr(self) -> Csr:
        return self.graph.by_user if self.which == "A" else self.graph.by_item

    def indices(self) -> torch.Tensor:
        c = self._csr()
        return torch.stack([c.row_ids(), c.idx[: c.nnz].to(torch.int64)])

    def values(self) -> torch.Tensor:
        c = self._csr()
        return c.val_fwd[: c.nnz]

    def _nnz(self) -> int:
        return self._csr().nnz

    def to_sparse_coo(self) -> torch.Tensor:
        return torch.sparse_coo_tensor(self.indices(), self.values(), size=self.shape).coalesce()

    def __repr__(self):
        return f"credgcn.Operator({self.which}, shape={self.shape}, nnz={self._nnz()}, variant={self.graph.variant})"


def numpy_damping_alpha(deg_i: np.ndarray) -> np.ndarray:
    """alpha_i = 1/log1p(max(deg_i, 1)) exactly as the reference evaluates it -- with NumPy on the
    host (Degree-Aware Message.py:379-380); libdevice's log1pf differs in the last bit."""
    return (np.float32(1.0) / np.log1p(np.maximum(deg_i.astype(np.float32), np.float32(1.0)))).astype(np.float32)


def build_graph(train_edges_2xE, num_users: int, num_items: int, cred_u, variant: str = "v2",
                device="cuda", reduce_item_degrees=None) -> CredGraph:
    """Degrees + sampling CSR + both coalesced operators, all on device, bit-exact vs the reference.

    reduce_item_degrees: for a user-sharded build -- a callable that turns this shard's int32 item
    degree histogram into the degrees over all shards (an all-reduce); `deg_i` of the result then
    holds the GLOBAL degrees (what the weights, the popularity law and alpha_i use)."""
    if variant not in _lib.VARIANTS:
        raise ValueError(f"variant must be one of {sorted(_lib.VARIANTS)}")
    dev = _cuda_device(device)
    U, I = int(num_users), int(num_items)
    eu, ei = _as_device_edges(train_edges_2xE, dev)
    E = eu.numel()
    if isinstance(cred_u, torch.Tensor):
        cred = cred_u.detach().reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
    else:
        cred = torch.from_numpy(np.ascontiguousarray(np.asarray(cred_u, dtype=np.float32))).to(dev)
    if cred.numel() != U:
        raise ValueError(f"cred_u has {cred.numel()} entries, expected num_users={U}")

    g = CredGraph(U, I, variant, dev)
    g.num_edges = E
    i32 = dict(dtype=torch.int32, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    i64 = dict(dtype=torch.int64, device=dev)
    g.deg_u, g.deg_i = torch.empty(U, **i32), torch.empty(I, **i32)
    g.samp_indptr, g.samp_idx = torch.empty(U + 1, **i64), torch.empty(max(E, 1), **i32)
    u_indptr, u_idx = torch.empty(U + 1, **i64), torch.empty(max(E, 1), **i32)
    u_vf, u_vb = torch.empty(max(E, 1), **f32), torch.empty(max(E, 1), **f32)
    i_indptr, i_idx = torch.empty(I + 1, **i64), torch.empty(max(E, 1), **i32)
    i_vf, i_vb = torch.empty(max(E, 1), **f32), torch.empty(max(E, 1), **f32)
    counts = torch.zeros(2, **i64)
    ws = workspace(lib().cgx_graph_build_workspace_bytes(E, U, I), dev)

    deg_i_global = None

    def run(alpha, deg_only):
        check(lib().cgx_graph_build(
            ptr(eu), ptr(ei), E, U, I, ptr(cred), _lib.VARIANTS[variant], ptr(alpha), ptr(deg_i_global),
            ptr(g.deg_u), ptr(g.deg_i),
            ptr(g.samp_indptr), ptr(g.samp_idx), ptr(u_indptr), ptr(u_idx), ptr(u_vf), ptr(u_vb),
            ptr(i_indptr), ptr(i_idx), ptr(i_vf), ptr(i_vb), ptr(counts), int(deg_only), ptr(ws), ws.numel(),
            stream_ptr(dev)))

    with torch.cuda.device(dev):
        alpha = None
        if variant == "da" or reduce_item_degrees is not None:
            run(None, True)                       # degrees first
            if reduce_item_degrees is not None:
                deg_i_global = reduce_item_degrees(g.deg_i.clone()).to(torch.int32).contiguous()
            if variant == "da":                   # alpha needs NumPy's log1p
                dsrc = deg_i_global if deg_i_global is not None else g.deg_i
                alpha = torch.from_numpy(numpy_damping_alpha(dsrc.cpu().numpy())).to(dev)
        run(alpha, False)
        nnz, bad = (int(x) for x in counts.cpu().tolist())
        g.deg_i_local = g.deg_i
        if deg_i_global is not None:
            g.deg_i = deg_i_global
    if bad:
        raise ValueError(f"{bad} edges have a user id outside [0, {U}) or an item id outside [0, {I})")
    g.nnz = nnz
    g.samp_idx = g.samp_idx[:E]
    g.by_user = Csr(U, I, nnz, u_indptr, u_idx[:max(nnz, 1)], u_vf[:max(nnz, 1)], u_vb[:max(nnz, 1)])
    g.by_item = Csr(I, U, nnz, i_indptr, i_idx[:max(nnz, 1)], i_vf[:max(nnz, 1)], i_vb[:max(nnz, 1)])
    if nnz < E:   # duplicates were merged: release the over-allocation
        for c in (g.by_user, g.by_item):
            c.idx, c.val_fwd, c.val_bwd = c.idx.clone(), c.val_fwd.clone(), c.val_bwd.clone()
    g.by_user.schedule()
    g.by_item.schedule()
    g.alpha = alpha
    return g


def user_csr_device(edges_2xE, num_users: int, num_items: int | None = None, device="cuda"):
    """edges_to_user_csr on device; returns (indptr int64[U+1], idx int32[E]) CUDA tensors."""
    dev = _cuda_device(device)
    eu, ei = _as_device_edges(edges_2xE, dev)
    E = eu.numel()
    if num_items is None:
        num_items = int(ei.max().item()) + 1 if E else 1
    indptr = torch.empty(num_users + 1, dtype=torch.int64, device=dev)
    idx = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    ws = workspace(lib().cgx_graph_build_workspace_bytes(E, num_users, num_items), dev)
    with torch.cuda.device(dev):
        check(lib().cgx_user_csr(ptr(eu), ptr(ei), E, num_users, num_items, ptr(indptr), ptr(idx), ptr(ws),
                                 ws.numel(), stream_ptr(dev)))
    return indptr, idx[:E]


# ------------------------------------------------------------------------------------------
# reference-named entry points
# ------------------------------------------------------------------------------------------
def edges_to_user_csr(edges_2xE, num_users: int, device="cuda"):
    """Drop-in for lightgcn_cu.py:259: returns NumPy (indptr int64[U+1], indices int64[E])."""
    indptr, idx = user_csr_device(edges_2xE, num_users, None, device)
    return indptr.cpu().numpy(), idx.cpu().numpy().astype(np.int64)


def build_cred_weighted_mats(train_edges, num_users, num_items, cred_u, device="cuda"):
    """Drop-in for lightgcn_cu.py:368: returns (M_ui [I x U] credibility operator,
    M_iu [U x I] base operator, deg_i float32 ndarray)."""
    g = build_graph(train_edges, num_users, num_items, cred_u, "cu", device)
    return g.operator("C"), g.operator("A"), g.deg_i_float()


def build_message_passing_mats(train_edges_2xE, num_users, num_items, cred_u, device="cuda", degree_aware=False):
    """Drop-in for Version-2/lighgcn_cu_pop.py:429 (and, with degree_aware=True, for the
    Method-A builder at Degree-Aware Message.py:349): returns (M_ui [U x I] base operator,
    M_iu [I x U] credibility operator) -- note the name swap against lightgcn_cu.py."""
    g = build_graph(train_edges_2xE, num_users, num_items, cred_u, "da" if degree_aware else "v2", device)
    if degree_aware:
        deg = g.deg_i_float()
        a = g.alpha.cpu().numpy()
        print(f"[POP] item_deg: min={deg.min():.0f} max={deg.max():.0f} mean={deg.mean():.2f} "
              f"nonzero={int(np.count_nonzero(deg))}/{num_items}")
        print(f"[POP] alpha_i:  min={a.min():.6f} median={np.median(a):.6f} max={a.max():.6f}")
    return g.operator("A"), g.operator("C")


def build_message_passing_mats_degree_aware(train_edges_2xE, num_users, num_items, cred_u, device="cuda"):
    return build_message_passing_mats(train_edges_2xE, num_users, num_items, cred_u, device, degree_aware=True)


class NormAdj:
    """Handle returned by build_norm_adj: the (U+I) x (U+I) symmetric operator of plain LightGCN
    (lightgcn.py:352-372) held as the two bipartite blocks of a CredGraph with credibility == 1."""

    def __init__(self, graph: CredGraph):
        self.graph = graph
        n = graph.num_users + graph.num_items
        self.shape = (n, n)

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def coalesce(self):
        return self

    def to_sparse_coo(self) -> torch.Tensor:
        g = self.graph
        a = g.operator("A")
        r, c = a.indices()
        v = a.values()
        idx = torch.cat([torch.stack([r, c + g.num_users]), torch.stack([c + g.num_users, r])], dim=1)
        return torch.sparse_coo_tensor(idx, torch.cat([v, v]), size=self.shape).coalesce()


def build_norm_adj(train_edges, num_users, num_items, device="cuda") -> NormAdj:
    """Drop-in for lightgcn.py:352: D^-1/2 A D^-1/2 of the bipartite graph (duplicate edges add up).
    The reference evaluates deg^-0.5 with torch.pow, which is not reproducible bit for bit across
    devices; values here are m / sqrt(deg_u * deg_i) in correctly rounded fp32 (within a few ulp)."""
    g = build_graph(train_edges, num_users, num_items, np.ones(num_users, dtype=np.float32), "cu", device)
    return NormAdj(g)


def graph_of(M_a, M_b) -> CredGraph:
    """The CredGraph behind a pair of operator views (order-insensitive; both must share it)."""
    if not (isinstance(M_a, Operator) and isinstance(M_b, Operator)):
        raise TypeError("expected the two operators returned by a credgcn build_*_mats call")
    if M_a.graph is not M_b.graph or {M_a.which, M_b.which} != {"A", "C"}:
        raise ValueError("the two operators must be the A and C views of the same graph")
    return M_a.graph
