"""autograd.Function shells around the libspkbgat kernels.

Math, per attention-layer group with heads h, a_h = [A1|A2|A3], a_2,h:
    P1 = X A1^T, P2 = X A2^T, P3 = Rel A3^T,  q* = P* a_2^T      (projection GEMMs, K1)
    m_e = P1[i] + P2[j] + P3[k](+P3[k2]),  s_e = q1[i] + q2[j] + q3[k](+q3[k2])
    ee_e = exp(-LeakyReLU(s_e)),  out_i = ELU( sum_e msk_e ee_e m_e / sum_e ee_e )   (K2)
which is GAT/layers.py:124-175 with the a.mm(edge_h) product re-associated so that no E-sized
feature tensor exists. The weights are handed to the kernels as "extended" matrices
    Wn [F, 2*Wd] = [A1^T | A1^T a_2^T | 0 || A2^T | A2^T a_2^T | 0],   Wr [Rd, Wd] = [A3^T | A3^T a_2^T | 0]
built from the reference-shaped parameters `a`, `a_2` with differentiable torch ops on these tiny
tensors, so autograd chains dWn, dWr back to a and a_2.
"""
import ctypes as C

import torch

from . import _lib
from .graph import KGraph

MAX_HEADS = 4
USE_TC = True          # route eligible products through the tcgen05 3xTF32 GEMM (else the exact-fp32 SIMT GEMM)
TC_MIN_ROWS = 1        # (tests lower/raise this to exercise both paths)


class Geometry:
    """Row format of the projected tables for H heads of width D (see include/spkbgat.h)."""

    def __init__(self, n_heads, d_head):
        assert 1 <= n_heads <= MAX_HEADS
        self.H, self.D = n_heads, d_head
        self.Dp = (d_head + 3) // 4 * 4
        self.Dt = self.H * self.Dp
        self.Wd = (self.Dt + self.H + 7) // 8 * 8
        if self.Wd > 512:
            raise ValueError(f"heads*out_features = {n_heads}*{d_head} exceeds the 512-float fused row")

    def struct(self):
        return _lib.Geom(self.H, self.D, self.Dp, self.Wd)


def extended_weights(a_list, a2_list, in_features, geom):
    """a_list[h]: [D, 2F+Rd], a2_list[h]: [1, D] (GAT/layers.py:100-105) -> Wn [F, 2Wd], Wr [Rd, Wd]."""
    F = in_features
    a0 = a_list[0]
    rd = a0.shape[1] - 2 * F
    Wn = a0.new_zeros(F, 2 * geom.Wd)
    Wr = a0.new_zeros(rd, geom.Wd)
    for h, (a, a2) in enumerate(zip(a_list, a2_list)):
        lo = h * geom.Dp
        at = a.t()                                       # [2F+Rd, D]
        qa = at.mm(a2.t()).squeeze(1)                    # [2F+Rd]  = a^T a_2^T
        Wn[:, lo:lo + geom.D] = at[:F]
        Wn[:, geom.Dt + h] = qa[:F]
        Wn[:, geom.Wd + lo:geom.Wd + lo + geom.D] = at[F:2 * F]
        Wn[:, geom.Wd + geom.Dt + h] = qa[F:2 * F]
        Wr[:, lo:lo + geom.D] = at[2 * F:]
        Wr[:, geom.Dt + h] = qa[2 * F:]
    return Wn, Wr


# ---- dense products --------------------------------------------------------------------------

def gemm_nn(A, B, out=None, accumulate=False):
    """out[M,N] (+)= A[M,K] @ B[K,N] on the library's GEMM (row-major, last-dim contiguous views allowed)."""
    assert A.dim() == 2 and B.dim() == 2 and A.shape[1] == B.shape[0]
    assert A.stride(1) == 1 and B.stride(1) == 1
    M, K = A.shape
    N = B.shape[1]
    if out is None:
        assert not accumulate
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    assert out.stride(1) == 1 and out.shape[0] == M and out.shape[1] == N
    if M and N:
        if K == 0:
            if not accumulate:
                out.zero_()
            return out
        lib = _lib.load()
        if USE_TC and M >= TC_MIN_ROWS and N <= 512 and lib.spk_gemm_nn_tc_supported(_lib.ptr(A), A.stride(0), M, N, K):
            # tcgen05 tensor cores, 3xTF32 (fp32-accurate)
            ws = torch.empty(lib.spk_gemm_tc_workspace_floats(N, K), dtype=torch.float32, device=A.device)
            _lib.check(lib.spk_gemm_nn_tc(_lib.ptr(A), A.stride(0), _lib.ptr(B), B.stride(0), _lib.ptr(out),
                                          out.stride(0), M, N, K, int(accumulate), _lib.ptr(ws), _lib.stream_ptr()),
                       "gemm_nn_tc")
        else:
            _lib.check(lib.spk_gemm_nn(_lib.ptr(A), A.stride(0), _lib.ptr(B), B.stride(0), _lib.ptr(out),
                                       out.stride(0), M, N, K, int(accumulate), _lib.stream_ptr()), "gemm_nn")
    return out


def gemm_tn(A, B, out=None, accumulate=False):
    """out[Ka,Nb] (+)= A[M,Ka]^T @ B[M,Nb], deterministic fixed-order split over M."""
    assert A.dim() == 2 and B.dim() == 2 and A.shape[0] == B.shape[0]
    assert A.stride(1) == 1 and B.stride(1) == 1
    M, Ka = A.shape
    Nb = B.shape[1]
    lib = _lib.load()
    if out is None:
        assert not accumulate
        out = torch.empty(Ka, Nb, dtype=torch.float32, device=A.device)
    if Ka and Nb:
        if USE_TC and M >= TC_MIN_ROWS and lib.spk_gemm_tn_tc_supported(_lib.ptr(A), A.stride(0), _lib.ptr(B), B.stride(0),
                                                                        M, Ka, Nb):
            ws = torch.empty(max(1, lib.spk_gemm_tn_tc_workspace_floats(M, Ka, Nb)), dtype=torch.float32, device=A.device)
            _lib.check(lib.spk_gemm_tn_tc(_lib.ptr(A), A.stride(0), _lib.ptr(B), B.stride(0), _lib.ptr(out),
                                          out.stride(0), M, Ka, Nb, int(accumulate), _lib.ptr(ws), _lib.stream_ptr()),
                       "gemm_tn_tc")
        else:
            ws = torch.empty(max(1, lib.spk_gemm_tn_workspace_floats(M, Ka, Nb)), dtype=torch.float32, device=A.device)
            _lib.check(lib.spk_gemm_tn(_lib.ptr(A), A.stride(0), _lib.ptr(B), B.stride(0), _lib.ptr(out), out.stride(0),
                                       M, Ka, Nb, int(accumulate), _lib.ptr(ws), _lib.stream_ptr()), "gemm_tn")
    return out


def tc_friendly(X):
    """Row stride multiple of 4 floats (16 B) so TMA can fetch the operand; pads a copy when needed."""
    if X.dim() != 2 or X.stride(1) != 1 or X.stride(0) % 4 == 0 or not USE_TC:
        return X
    f = X.shape[1]
    Xp = X.new_zeros(X.shape[0], (f + 3) // 4 * 4)
    Xp[:, :f] = X
    return Xp[:, :f]


class MatMulFn(torch.autograd.Function):
    """X @ W with the library GEMMs (relation_embed.mm(W) models.py:77, entity_embeddings.mm(W_entities) 175)."""

    @staticmethod
    def forward(ctx, X, W, dist=None):
        X = tc_friendly(X.contiguous()); W = W.contiguous()
        ctx.save_for_backward(X, W)
        ctx.dist = dist
        return gemm_nn(X, W)

    @staticmethod
    def backward(ctx, g):
        X, W = ctx.saved_tensors
        g = g.contiguous()
        dX = gemm_nn(g, W.t().contiguous()) if ctx.needs_input_grad[0] else None
        dW = gemm_tn(X, g) if ctx.needs_input_grad[1] else None
        if dW is not None and ctx.dist is not None:
            ctx.dist.all_reduce(dW)          # X rows are partitioned across ranks
        return dX, dW, None


def matmul(X, W, dist=None):
    """dist: DistContext when X's rows are partitioned across ranks (dW is then all-reduced)."""
    return MatMulFn.apply(X, W, dist)


# ---- fused attention-layer group -------------------------------------------------------------

def _hub_partial(hubs, ldpart, device):
    if hubs is None or hubs.n_tasks == 0:
        return None
    return torch.empty(hubs.n_tasks, ldpart, dtype=torch.float32, device=device)


def edge_attn_forward(graph, P1, P2, P3, geom, alpha, apply_elu, mask_csr, nanflag):
    """K2 launch. P1/P2: [n, >=Wd] views with unit inner stride; returns (out [N,H*D], den [N,H], sw [N,H])."""
    lib = _lib.load()
    n = graph.n_nodes
    dev = P1.device
    out = torch.empty(n, geom.H * geom.D, dtype=torch.float32, device=dev)
    den = torch.empty(n, geom.H, dtype=torch.float32, device=dev)
    sw = torch.empty(n, geom.H, dtype=torch.float32, device=dev)
    a = _lib.EdgeFwdArgs()
    a.segptr = graph.rowptr.data_ptr(); a.col = graph.col.data_ptr(); a.t1 = graph.t1.data_ptr()
    a.t2 = graph.t2.data_ptr() if graph.t2 is not None else None
    a.P1 = P1.data_ptr(); a.ld1 = P1.stride(0)
    a.P2 = P2.data_ptr(); a.ld2 = P2.stride(0)
    a.P3 = P3.data_ptr(); a.ld3 = P3.stride(0)
    if mask_csr is not None:
        a.mask = mask_csr.data_ptr(); a.mask_stride = mask_csr.stride(0)
    a.out = out.data_ptr(); a.ldo = out.stride(0); a.den = den.data_ptr(); a.sw = sw.data_ptr()
    a.nanflag = nanflag.data_ptr()
    a.n_rows = n; a.apply_elu = int(apply_elu); a.alpha = float(alpha)
    a.geom = geom.struct()
    ldpart = geom.Wd + 2 * MAX_HEADS
    partial = _hub_partial(graph.row_hubs, ldpart, dev)
    graph.row_hubs.fill(a.hub, partial, ldpart)
    _lib.check(lib.spk_edge_attn_fwd(C.byref(a), _lib.stream_ptr()), "edge_attn_fwd")
    return out, den, sw


def edge_attn_backward(graph, P1, P2, P3, geom, alpha, apply_elu, mask_csr, out, dout, den, dP1, dP2, dP3):
    """K3 + K4 launches. Fills dP1 [n_rows, Wd], dP2 [n_cols, Wd] (indexed by gathered node) and dP3 [R, Wd]."""
    lib = _lib.load()
    graph.build_backward()
    n, dev = graph.n_nodes, P1.device
    ldg = (geom.Dt + 7) // 8 * 8
    G = torch.empty(n, ldg, dtype=torch.float32, device=dev)
    rec = torch.empty(max(1, graph.n_edges), 2 * geom.H, dtype=torch.float32, device=dev)
    a = _lib.EdgeBwdRowsArgs()
    a.segptr = graph.rowptr.data_ptr(); a.col = graph.col.data_ptr(); a.t1 = graph.t1.data_ptr()
    a.t2 = graph.t2.data_ptr() if graph.t2 is not None else None
    a.P1 = P1.data_ptr(); a.ld1 = P1.stride(0)
    a.P2 = P2.data_ptr(); a.ld2 = P2.stride(0)
    a.P3 = P3.data_ptr(); a.ld3 = P3.stride(0)
    if mask_csr is not None:
        a.mask = mask_csr.data_ptr(); a.mask_stride = mask_csr.stride(0)
    a.out = out.data_ptr(); a.dout = dout.data_ptr(); a.ldo = out.stride(0)
    assert dout.stride(0) == out.stride(0)
    a.den = den.data_ptr()
    a.G = G.data_ptr(); a.ldg = ldg; a.dP1 = dP1.data_ptr(); a.ldd1 = dP1.stride(0); a.rec = rec.data_ptr()
    a.n_rows = n; a.apply_elu = int(apply_elu); a.alpha = float(alpha)
    a.geom = geom.struct()
    part_a = _hub_partial(graph.row_hubs, 2 * MAX_HEADS, dev)
    graph.row_hubs.fill(a.hub, part_a, 2 * MAX_HEADS)
    _lib.check(lib.spk_edge_attn_bwd_rows(C.byref(a), _lib.stream_ptr()), "edge_attn_bwd_rows")

    def seg(ptr, src, pos, hubs, dst, n_seg, tag):
        s = _lib.SegGatherArgs()
        s.segptr = ptr.data_ptr(); s.src = src.data_ptr(); s.pos = pos.data_ptr()
        s.G = G.data_ptr(); s.ldg = ldg; s.rec = rec.data_ptr()
        s.out = dst.data_ptr(); s.ldout = dst.stride(0); s.n_seg = n_seg
        s.flags = 1 if (n_seg > 0 and src.numel() < 5 * n_seg) else 0        # many short segments -> streaming kernel
        s.geom = geom.struct()
        part = _hub_partial(hubs, geom.Wd, dev)
        hubs.fill(s.hub, part, geom.Wd)
        _lib.current_tag = tag
        try:
            _lib.check(lib.spk_edge_attn_bwd_segments(C.byref(s), _lib.stream_ptr()), "edge_attn_bwd_segments")
        finally:
            _lib.current_tag = ""

    seg(graph.colptr, graph.csc_row, graph.csc_pos, graph.col_hubs, dP2, graph.n_cols, "cols")
    seg(graph.relptr, graph.rel_row, graph.rel_pos, graph.rel_hubs, dP3, graph.n_rel, "rels")


class AttentionGroupFn(torch.autograd.Function):
    """One fused group of <=4 heads: (X, Wn, Rel, Wr) -> ELU?(attention output) [N, H*D].

    Multi-GPU (graph.dist set): rows are partitioned; the gathered table P2~ must cover all nodes.
      * "proj"  exchange: all-gather the projected rows P2~ [n, Wd]; backward reduce-scatters the partial dP2~.
      * "input" exchange (chosen when the input is narrower than the projection, i.e. layer 1: F=50 vs Wd=208):
        all-gather X [n, F] and project all nodes locally (cheap on the tensor cores); backward projects the partial
        dP2~ back to input space and reduce-scatters [n, F]. 4x fewer bytes over NVLink for layer 1.
    """

    @staticmethod
    def forward(ctx, X, Wn, Rel, Wr, graph, geom, alpha, apply_elu, mask_csr, nanflag):
        X = tc_friendly(X.contiguous()); Wn = Wn.contiguous(); Rel = Rel.contiguous(); Wr = Wr.contiguous()
        dist = getattr(graph, "dist", None)
        Wd = geom.Wd
        P3 = gemm_nn(Rel, Wr)                   # [R, Wd]
        mode = "local"
        X_all = None
        if dist is None:
            P = gemm_nn(X, Wn)                  # [N, 2Wd] = [P1~ | P2~]
            P1, P2 = P[:, :Wd], P[:, Wd:]
        elif 2 * X.shape[1] <= Wd:
            mode = "input"
            X_all = tc_friendly(dist.all_gather_rows(X))
            P1 = gemm_nn(X, Wn[:, :Wd])
            P2 = gemm_nn(X_all, Wn[:, Wd:])     # every rank projects all nodes
        else:
            mode = "proj"
            P = gemm_nn(X, Wn)
            P1 = P[:, :Wd]
            P2 = dist.all_gather_rows(P[:, Wd:])
        out, den, sw = edge_attn_forward(graph, P1, P2, P3, geom, alpha, apply_elu, mask_csr, nanflag)
        ctx.save_for_backward(X, Wn, Rel, Wr, P1, P2, P3, out, den, X_all if X_all is not None else X.new_empty(0))
        ctx.graph, ctx.geom, ctx.alpha, ctx.apply_elu, ctx.mask_csr, ctx.mode = graph, geom, alpha, apply_elu, mask_csr, mode
        ctx.mark_non_differentiable(den, sw)
        return out, den, sw

    @staticmethod
    def backward(ctx, dout, _dden, _dsw):
        geom, graph, mode = ctx.geom, ctx.graph, ctx.mode
        dist = getattr(graph, "dist", None)
        X, Wn, Rel, Wr, P1, P2, P3, out, den, X_all = ctx.saved_tensors
        Wd = geom.Wd
        dout = dout.contiguous()
        n = X.shape[0]
        dP3 = torch.empty_like(P3)
        WnT = Wn.t().contiguous()               # [2Wd, F]
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dX = dWn = None
        if mode == "local":
            dP = torch.empty(n, 2 * Wd, dtype=torch.float32, device=X.device)
            edge_attn_backward(graph, P1, P2, P3, geom, ctx.alpha, ctx.apply_elu, ctx.mask_csr, out, dout, den,
                               dP[:, :Wd], dP[:, Wd:], dP3)
            if need_x:
                dX = gemm_nn(dP, WnT)
            if need_w:
                dWn = gemm_tn(X, dP)
        elif mode == "proj":
            dP = torch.empty(n, 2 * Wd, dtype=torch.float32, device=X.device)
            dP2_all = torch.empty_like(P2)      # partial over this rank's edges, all gathered nodes
            edge_attn_backward(graph, P1, P2, P3, geom, ctx.alpha, ctx.apply_elu, ctx.mask_csr, out, dout, den,
                               dP[:, :Wd], dP2_all, dP3)
            dist.reduce_scatter_rows(dP2_all, dP[:, Wd:])
            del dP2_all
            dist.all_reduce(dP3)
            if need_x:
                dX = gemm_nn(dP, WnT)
            if need_w:
                dWn = dist.all_reduce(gemm_tn(X, dP))
        else:                                   # "input"
            dP1 = torch.empty(n, Wd, dtype=torch.float32, device=X.device)
            dP2_all = torch.empty_like(P2)
            edge_attn_backward(graph, P1, P2, P3, geom, ctx.alpha, ctx.apply_elu, ctx.mask_csr, out, dout, den,
                               dP1, dP2_all, dP3)
            dist.all_reduce(dP3)
            if need_x:
                dX2_all = gemm_nn(dP2_all, WnT[Wd:])            # partial dX of every node through the gathered side
                dX = torch.empty(n, X.shape[1], dtype=torch.float32, device=X.device)
                dist.reduce_scatter_rows(dX2_all, dX)
                del dX2_all
                gemm_nn(dP1, WnT[:Wd], out=dX, accumulate=True)
            if need_w:
                dWn = torch.empty_like(Wn)
                gemm_tn(X, dP1, out=dWn[:, :Wd])
                gemm_tn(X_all, dP2_all, out=dWn[:, Wd:])
                dist.all_reduce(dWn)
        dRel = gemm_nn(dP3, Wr.t().contiguous()) if ctx.needs_input_grad[2] else None
        dWr = gemm_tn(Rel, dP3) if ctx.needs_input_grad[3] else None
        return dX, dWn, dRel, dWr, None, None, None, None, None, None


def attention_group(X, Rel, a_list, a2_list, graph, alpha, apply_elu, mask_csr, nanflag):
    """All heads of one layer (looping over groups of <=4): returns [N, sum_h D]."""
    outs = []
    F = X.shape[1]
    D = a_list[0].shape[0]
    for g0 in range(0, len(a_list), MAX_HEADS):
        al, a2l = a_list[g0:g0 + MAX_HEADS], a2_list[g0:g0 + MAX_HEADS]
        geom = Geometry(len(al), D)
        Wn, Wr = extended_weights(al, a2l, F, geom)
        m = None if mask_csr is None else mask_csr[g0:g0 + MAX_HEADS].contiguous()
        out, _, _ = AttentionGroupFn.apply(X, Wn, Rel, Wr, graph, geom, alpha, apply_elu, m, nanflag)
        outs.append(out)
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)


# ---- row-wise wrappers -----------------------------------------------------------------------

def rownorm_(x):
    """In-place L2 row normalisation of a contiguous [N, F] tensor (models.py:160-161)."""
    assert x.dim() == 2 and x.stride(1) == 1
    _lib.check(_lib.load().spk_rownorm(_lib.ptr(x), x.stride(0), _lib.ptr(x), x.stride(0), x.shape[0], x.shape[1],
                                       _lib.stream_ptr()), "rownorm")
    return x


def rownorm(x):
    x = x.contiguous()
    y = torch.empty_like(x)
    _lib.check(_lib.load().spk_rownorm(_lib.ptr(x), x.stride(0), _lib.ptr(y), y.stride(0), x.shape[0], x.shape[1],
                                       _lib.stream_ptr()), "rownorm")
    return y


def mask_from_index(idx, n_rows, device):
    """mask[unique(idx)] = 1 (models.py:167-173); duplicates are harmless so no unique() is needed."""
    mask = torch.zeros(n_rows, dtype=torch.float32, device=device)
    idx = idx.to(device=device, dtype=torch.int64).contiguous()
    if idx.numel():
        if int(idx.min()) < -n_rows or int(idx.max()) >= n_rows:
            raise IndexError("batch_entities index out of range")
        idx = torch.where(idx < 0, idx + n_rows, idx)
        _lib.check(_lib.load().spk_mask_from_index(_lib.ptr(idx), idx.numel(), _lib.ptr(mask), n_rows,
                                                   _lib.stream_ptr()), "mask_from_index")
    return mask


class ResidualNormFn(torch.autograd.Function):
    """normalize(EW + mask[:,None]*x2) (models.py:175-179)."""

    @staticmethod
    def forward(ctx, EW, x2, mask):
        EW = EW.contiguous(); x2 = x2.contiguous()
        n, w = EW.shape
        out = torch.empty_like(EW)
        inv = torch.empty(n, dtype=torch.float32, device=EW.device)
        _lib.check(_lib.load().spk_residual_norm_fwd(_lib.ptr(EW), EW.stride(0), _lib.ptr(x2), x2.stride(0),
                                                     _lib.ptr(mask), _lib.ptr(out), out.stride(0), _lib.ptr(inv),
                                                     n, w, _lib.stream_ptr()), "residual_norm_fwd")
        ctx.save_for_backward(out, mask, inv)
        return out

    @staticmethod
    def backward(ctx, g):
        out, mask, inv = ctx.saved_tensors
        g = g.contiguous()
        n, w = out.shape
        dew = torch.empty_like(out)
        dx2 = torch.empty_like(out)
        _lib.check(_lib.load().spk_residual_norm_bwd(_lib.ptr(g), g.stride(0), _lib.ptr(out), out.stride(0),
                                                     _lib.ptr(mask), _lib.ptr(inv), _lib.ptr(dew), dew.stride(0),
                                                     _lib.ptr(dx2), dx2.stride(0), n, w, _lib.stream_ptr()),
                   "residual_norm_bwd")
        return dew, dx2, None
