"""Drop-in for the reference's GAT/layers.py (same class names, constructor and forward signatures,
parameter names and shapes), running on libspkbgat's sm_100a kernels. No CPU path.

  SpecialSpmmFunctionFinal / SpecialSpmmFinal  <- GAT/layers.py:51-84
  SpGraphAttentionLayer                        <- GAT/layers.py:87-181
  ConvKB                                       <- GAT/layers.py:12-48 (lives in recon_b200/convkb.py; re-exported here)
"""
import torch
import torch.nn as nn

from . import _lib
from . import functional as SF
from .graph import KGraph, sort_pairs, _segment_ptr, _iota, _key_bits


class SpecialSpmmFunctionFinal(torch.autograd.Function):
    """out[i, :] = sum over edges e with edge[0, e] == i of edge_w[e, :]; gradient only w.r.t. edge_w
    (a gather of grad rows by edge[0]), exactly the contract of GAT/layers.py:51-79."""

    @staticmethod
    def forward(ctx, edge, edge_w, N, E, out_features):
        lib = _lib.load()
        if not edge_w.is_cuda:
            raise RuntimeError("recon_b200.SpecialSpmmFunctionFinal needs CUDA tensors (no CPU fallback)")
        dev = edge_w.device
        with torch.cuda.device(dev):
            return SpecialSpmmFunctionFinal._forward(ctx, edge, edge_w, N, dev)

    @staticmethod
    def _forward(ctx, edge, edge_w, N, dev):
        lib = _lib.load()
        rows64 = edge[0].to(device=dev, dtype=torch.int64).contiguous()
        if rows64.numel() and (int(rows64.min()) < 0 or int(rows64.max()) >= N):
            raise IndexError("edge row index out of range")
        keys, perm = sort_pairs(rows64.int(), _iota(rows64.numel(), dev), _key_bits(N))
        segptr = _segment_ptr(keys, N)
        w = edge_w.contiguous().float()
        width = w.shape[1]
        out = torch.empty(N, width, dtype=torch.float32, device=dev)
        _lib.check(lib.spk_spmm_rowsum_fwd(segptr.data_ptr(), _lib.ptr(perm) if perm.numel() else None,
                                           _lib.ptr(w) if w.numel() else None, w.stride(0), width,
                                           out.data_ptr(), out.stride(0), N, _lib.stream_ptr()), "spmm_rowsum_fwd")
        ctx.save_for_backward(rows64)
        ctx.width = width
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (rows64,) = ctx.saved_tensors
        grad_values = None
        if ctx.needs_input_grad[1]:
            g = grad_output.contiguous()
            e = rows64.numel()
            grad_values = torch.empty(e, ctx.width, dtype=torch.float32, device=g.device)
            _lib.check(_lib.load().spk_spmm_rowsum_bwd(_lib.ptr(rows64) if e else None, _lib.ptr(g), g.stride(0), ctx.width,
                                                       _lib.ptr(grad_values) if e else None, ctx.width, e,
                                                       _lib.stream_ptr()), "spmm_rowsum_bwd")
        return None, grad_values, None, None, None


class SpecialSpmmFinal(nn.Module):
    def forward(self, edge, edge_w, N, E, out_features):
        return SpecialSpmmFunctionFinal.apply(edge, edge_w, N, E, out_features)


def check_nanflag(nanflag):
    """The reference asserts `not isnan` at layers.py:147,167,172; same AssertionError here, one sync per forward.
    Bit 1 of the word is the batch-index range check of models.py:167-173 (IndexError in the reference)."""
    v = int(nanflag.item())
    if v != 0:
        nanflag.zero_()
        if v & 1:
            raise AssertionError("NaN in attention coefficients / aggregated features (reference: GAT/layers.py:147,167,172)")
        raise IndexError("batch_entities index out of range")


def edge_dropout_mask(p, n_heads, n_edges, device):
    """Inverted-dropout multipliers for the attention coefficients (nn.Dropout at layers.py:158),
    drawn directly in CSR order (i.i.d., so the order does not matter)."""
    keep = torch.rand(n_heads, n_edges, device=device) >= p
    return keep.to(torch.float32).mul_(1.0 / (1.0 - p))


class SpGraphAttentionLayer(nn.Module):
    """Sparse KBGAT attention layer; parameters `a` [out, 2*in + nrela], `a_2` [1, out] as in the reference."""

    def __init__(self, num_nodes, in_features, out_features, nrela_dim, dropout, alpha, concat=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.num_nodes = num_nodes
        self.alpha = alpha
        self.concat = concat
        self.nrela_dim = nrela_dim
        self.a = nn.Parameter(torch.zeros(size=(out_features, 2 * in_features + nrela_dim)))
        nn.init.xavier_normal_(self.a.data, gain=1.414)
        self.a_2 = nn.Parameter(torch.zeros(size=(1, out_features)))
        nn.init.xavier_normal_(self.a_2.data, gain=1.414)
        self.dropout = nn.Dropout(dropout)          # only .p is used; masks go to the fused kernel
        self.leakyrelu = nn.LeakyReLU(self.alpha)
        self.special_spmm_final = SpecialSpmmFinal()

    def forward(self, input, edge, edge_embed, edge_list_nhop, edge_embed_nhop, dropout_mask=None):
        """Stand-alone call with per-edge embeddings (GAT/layers.py:111): every edge is treated as its
        own relation row, so the same fused kernels (and their backward into edge_embed) apply.
        `dropout_mask`: optional [E] multipliers in the caller's edge order (extension for parity tests)."""
        if not input.is_cuda:
            raise RuntimeError("recon_b200.SpGraphAttentionLayer needs CUDA tensors (no CPU fallback)")
        with torch.cuda.device(input.device):
            return self._forward(input, edge, edge_embed, edge_list_nhop, edge_embed_nhop, dropout_mask)

    def _forward(self, input, edge, edge_embed, edge_list_nhop, edge_embed_nhop, dropout_mask):
        dev = input.device
        n = input.shape[0]
        has2 = edge_list_nhop is not None and edge_list_nhop.numel() > 0
        if has2:
            edge = torch.cat((edge.to(dev), edge_list_nhop.to(dev)), dim=1)
            edge_embed = torch.cat((edge_embed, edge_embed_nhop.to(dev)), dim=0)
        e = edge.shape[1]
        if edge_embed.shape[0] != e:                     # torch.cat(..., dim=1) at layers.py:129 raises the same way
            raise RuntimeError(f"edge_embed has {edge_embed.shape[0]} rows for {e} edges")
        etype = torch.arange(e, device=dev, dtype=torch.int64)
        graph = KGraph(edge.to(dev), etype, None, n, max(e, 1), device=dev)
        p = self.dropout.p
        mask_csr = None
        if dropout_mask is not None:
            mask_csr = graph.to_csr_order(dropout_mask.to(dev, torch.float32).reshape(1, e))
        elif self.training and p > 0:
            mask_csr = edge_dropout_mask(p, 1, e, dev)
        nanflag = torch.zeros(1, dtype=torch.int32, device=dev)
        out = SF.attention_group(input, edge_embed, [self.a], [self.a_2], graph, self.alpha, self.concat,
                                 mask_csr, nanflag)
        check_nanflag(nanflag)
        return out

    def __repr__(self):
        return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'


def __getattr__(name):
    """`from layers import SpGraphAttentionLayer, ConvKB` (GAT/models.py:6) keeps working: ConvKB lives in convkb.py
    (imported lazily: that module imports this one)."""
    if name == "ConvKB":
        from .convkb import ConvKB
        return ConvKB
    raise AttributeError(name)
