"""CPU-side checks of the drop-in boundary: the C-ABI library loads here (no GPU needed) and exports
every symbol include/spkbgat.h declares; the Python binding covers all of them; the product never
imports the oracle."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "spkbgat.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(spk_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    from recon_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        ge.build()
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.spk_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match_header_layout():
    import ctypes as C
    from recon_b200 import _lib
    assert C.sizeof(_lib.Geom) == 16
    assert C.sizeof(_lib.HubTasks) == 5 * 8 + 8 + 8 + 16 + 8        # + task_order (ABI 6)
    assert C.sizeof(_lib.EdgeFwdArgs) == 17 * 8 + 16 + C.sizeof(_lib.Geom) + C.sizeof(_lib.HubTasks)
    assert C.sizeof(_lib.EdgeBwdRowsArgs) == 21 * 8 + 16 + C.sizeof(_lib.Geom) + C.sizeof(_lib.HubTasks)
    assert C.sizeof(_lib.SegGatherArgs) == 8 * 8 + 8 + C.sizeof(_lib.Geom) + C.sizeof(_lib.HubTasks)
    assert C.sizeof(_lib.AttnWeightsArgs) == 16 * 8 + 10 * 4 + 5 * 8       # spk_attn_weights_args: 16 pointers, 10 ints, W0/ld0/W1/ld1/W2


def test_attn_weights_argument_checks_without_gpu():
    """spk_attn_weights_* validate their argument block before any device work (no GPU needed for the error paths)."""
    import ctypes as C
    from recon_b200 import _lib
    lib = _lib.load()
    w = _lib.AttnWeightsArgs()
    w.n_heads, w.F, w.Rd, w.D, w.mode = 5, 4, 4, 4, 0                      # more than 4 heads
    assert lib.spk_attn_weights_fwd(C.byref(w), None) != 0
    assert b"attn_weights" in lib.spk_last_error()
    w.n_heads, w.mode = 1, 7                                               # unknown mode
    assert lib.spk_attn_weights_bwd(C.byref(w), None) != 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "recon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"


def test_no_cpu_fallback():
    import torch
    from recon_b200 import SpKBGATModified, SpecialSpmmFunctionFinal
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU behaviour")
    m = SpKBGATModified(torch.randn(8, 4), torch.randn(3, 4), [4, 8], [4, 8], 0.0, 0.2, [2, 2], None)
    edge = torch.randint(0, 8, (2, 10)); et = torch.randint(0, 3, (10,))
    with pytest.raises(RuntimeError):
        m(None, torch.arange(8), (edge, et), torch.zeros(0, 4, dtype=torch.long))
    with pytest.raises(RuntimeError):
        SpecialSpmmFunctionFinal.apply(edge, torch.randn(10, 2), 8, 10, 2)


def test_extended_weights_reassociation_is_exact_in_fp64():
    """Host logic: Wn/Wr reproduce a.[h_i|h_j|r] and a_2.(a.[...]) (GAT/layers.py:129-143) -- pure torch, CPU."""
    import torch
    from recon_b200.functional import Geometry, extended_weights
    torch.manual_seed(0)
    F, Rd, D, H, n, r = 6, 5, 7, 3, 11, 4
    a = [torch.randn(D, 2 * F + Rd, dtype=torch.float64) for _ in range(H)]
    a2 = [torch.randn(1, D, dtype=torch.float64) for _ in range(H)]
    geom = Geometry(H, D)
    Wn, Wr = extended_weights(a, a2, F, geom)
    X = torch.randn(n, F, dtype=torch.float64); R = torch.randn(r, Rd, dtype=torch.float64)
    P = X @ Wn; P3 = R @ Wr
    i, j, k = 3, 7, 2
    for h in range(H):
        eh = torch.cat((X[i], X[j], R[k]))
        m = a[h] @ eh
        lo = h * geom.Dp
        got = P[i, lo:lo + D] + P[j, geom.Wd + lo:geom.Wd + lo + D] + P3[k, lo:lo + D]
        assert torch.allclose(got, m, atol=1e-12)
        s = (a2[h] @ m).item()
        got_s = (P[i, geom.Dt + h] + P[j, geom.Wd + geom.Dt + h] + P3[k, geom.Dt + h]).item()
        assert abs(s - got_s) < 1e-12
