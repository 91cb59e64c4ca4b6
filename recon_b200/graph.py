"""On-device edge construction: triples / adjacency tensors -> CSR, CSC and per-relation segments.

Host-side mirror of what the reference does in Python before and inside the layers:
  * GAT/models.py:141-148   nhop rows [s, r1, r2, t] -> edge_list_nhop=[t; s], edge_type_nhop=[r1, r2]
  * GAT/layers.py:124-127   concatenate 1-hop then 2-hop edges
  * GAT/layers.py:56-58     the coalesce-by-row hidden in sparse_coo_tensor + sparse.sum
The aggregation key is edge[0] (= triple tail, GAT/preprocess.py:78-80); edge[1] is the gather index.
All integer work runs in libspkbgat (spk_edges_concat, spk_sort_pairs, spk_segment_ptr, ...); torch is
used for allocation and for the O(#hubs) task tables only.
"""
import ctypes as C
import os

import torch

from . import _lib

HUB_THRESH = 512      # segments longer than this are split ...
HUB_CHUNK = 256       # ... into tasks of this many edges
HUB_MAX_TASKS = 4096  # ... but never more tasks than this per segment (longer tasks instead; long tasks are a tail, keep it high)
# relation segments (E / R edges each, sorted by aggregation row inside a relation): short tasks launched in ascending order
# of their first row, so the G rows gathered by the tasks in flight form one sliding window that fits in L2
REL_CHUNK = int(os.environ.get("SPK_REL_CHUNK", "64"))
REL_WINDOW_ORDER = os.environ.get("SPK_REL_ORDER", "1") != "0"


class HubSet:
    """Task table for the segments of one ordering that exceed HUB_THRESH (see spk_hub_tasks)."""

    def __init__(self, ptr, thresh=HUB_THRESH, chunk=HUB_CHUNK, order_by=None, order_bits=31):
        """order_by: optional per-entry int32 key array (e.g. the aggregation row of every entry, ascending inside a
        segment); tasks are then LAUNCHED in ascending order of the key of their first entry (task_order), while partials
        and their summation order stay in task order."""
        self.thresh = thresh
        self.n_tasks = 0
        self.n_hubs = 0
        self.task_seg = self.task_beg = self.task_end = self.hub_seg = self.hub_task_ptr = self.task_order = None
        if ptr.numel() <= 1:
            return
        ptr64 = ptr.long()
        deg = ptr64[1:] - ptr64[:-1]
        hub_seg = (deg > thresh).nonzero().flatten()
        self.n_hubs = int(hub_seg.numel())
        if self.n_hubs == 0:
            return
        # at most HUB_MAX_TASKS partials per hub (the finalize kernel adds them with one CTA): giant hubs get longer tasks
        hdeg = deg[hub_seg]
        hchunk = torch.clamp(((hdeg + HUB_MAX_TASKS - 1) // HUB_MAX_TASKS + 31) // 32 * 32, min=chunk)
        ntask = (hdeg + hchunk - 1) // hchunk
        tptr = torch.zeros(self.n_hubs + 1, dtype=torch.int64, device=ptr.device)
        tptr[1:] = torch.cumsum(ntask, 0)
        self.n_tasks = int(tptr[-1].item())
        task_hub = torch.repeat_interleave(torch.arange(self.n_hubs, device=ptr.device), ntask)
        local = torch.arange(self.n_tasks, device=ptr.device) - tptr[task_hub]
        seg = hub_seg[task_hub]
        tchunk = hchunk[task_hub]
        beg = ptr64[seg] + local * tchunk
        end = torch.minimum(beg + tchunk, ptr64[seg + 1])
        self.task_seg = seg.int().contiguous()
        self.task_beg = beg.int().contiguous()
        self.task_end = end.int().contiguous()
        self.hub_seg = hub_seg.int().contiguous()
        self.hub_task_ptr = tptr.int().contiguous()
        if order_by is not None and self.n_tasks > 1:
            key = _gather(order_by, self.task_beg)
            _, self.task_order = sort_pairs(key, _iota(self.n_tasks, ptr.device), order_bits)

    def fill(self, hub, partial, ldpart):
        """Populate a _lib.HubTasks struct; `partial` is the per-call scratch tensor (or None)."""
        hub.n_tasks = self.n_tasks
        hub.n_hubs = self.n_hubs
        hub.hub_thresh = self.thresh
        if self.n_tasks:
            hub.task_seg = self.task_seg.data_ptr(); hub.task_beg = self.task_beg.data_ptr()
            hub.task_end = self.task_end.data_ptr(); hub.hub_seg = self.hub_seg.data_ptr()
            hub.hub_task_ptr = self.hub_task_ptr.data_ptr()
            hub.partial = partial.data_ptr() if partial is not None else None; hub.ldpart = ldpart
            hub.task_order = self.task_order.data_ptr() if self.task_order is not None else None


def triples_to_adj(triples, is_unweigted=False, directed=True):
    """(head, rel, tail) id rows in file order -> the adjacency tensors the reference feeds the model:
    `preprocess.load_data` (GAT/preprocess.py:48-87) + `Corpus.__init__` (GAT/create_batch.py:28-31), i.e.
    edge_list = [rows = tail; cols = head] (the tail aggregates from the head), edge_type = relation id (1 when
    is_unweigted); directed=False puts the reversed edge (rows = head, cols = tail) in front of every edge, as the
    reference's append order does. Pure index glue on whatever device `triples` lives on."""
    tr = torch.as_tensor(triples).to(torch.int64).reshape(-1, 3)
    h, r, t = tr[:, 0], tr[:, 1], tr[:, 2]
    val = torch.ones_like(r) if is_unweigted else r
    if directed:
        return torch.stack((t, h), dim=0).contiguous(), val.contiguous()
    rows = torch.stack((h, t), dim=1).reshape(-1)            # reversed edge first, then the edge itself (preprocess.py:66-82)
    cols = torch.stack((t, h), dim=1).reshape(-1)
    return torch.stack((rows, cols), dim=0).contiguous(), torch.stack((val, val), dim=1).reshape(-1).contiguous()


def _key_bits(n):
    return max(1, int(n - 1).bit_length()) if n > 1 else 1


def sort_pairs(keys, vals, key_bits):
    """Stable radix sort by key of int32 (key, value) pairs on the device. Returns (keys, vals) sorted."""
    lib = _lib.load()
    n = keys.numel()
    if n == 0:
        return keys, vals
    ktmp, vtmp = torch.empty_like(keys), torch.empty_like(vals)
    ws = torch.empty(lib.spk_sort_workspace_bytes(n), dtype=torch.uint8, device=keys.device)
    in_tmp = C.c_int32(0)
    _lib.check(lib.spk_sort_pairs(keys.data_ptr(), vals.data_ptr(), ktmp.data_ptr(), vtmp.data_ptr(), n, key_bits,
                                  ws.data_ptr(), C.byref(in_tmp), _lib.stream_ptr()), "sort_pairs")
    return (ktmp, vtmp) if in_tmp.value else (keys, vals)


def _iota(n, device):
    v = torch.empty(n, dtype=torch.int32, device=device)
    _lib.check(_lib.load().spk_iota_i32(v.data_ptr(), n, _lib.stream_ptr()), "iota")
    return v


def _gather(src, idx):
    out = torch.empty(idx.numel(), dtype=torch.int32, device=src.device)
    _lib.check(_lib.load().spk_gather_i32(src.data_ptr(), idx.data_ptr(), idx.numel(), out.data_ptr(),
                                          _lib.stream_ptr()), "gather_i32")
    return out


def _segment_ptr(sorted_keys, n_seg):
    ptr = torch.empty(n_seg + 1, dtype=torch.int32, device=sorted_keys.device)
    _lib.check(_lib.load().spk_segment_ptr(sorted_keys.data_ptr() if sorted_keys.numel() else None,
                                           sorted_keys.numel(), n_seg, ptr.data_ptr(), _lib.stream_ptr()), "segment_ptr")
    return ptr


PACK_CHUNK = int(os.environ.get("SPK_PACK_CHUNK", str(8 << 20)))      # elements per pack + copy step of the host stager
_STAGING = {}          # device index -> {"buf": pinned int32 tensor, "event": last H2D that read it, "stream": copy stream}


class _HostStager:
    """HOST int64 edge tensors -> device int32 (row, col, t1, t2) arrays, produced on demand.

    The reference hands the adjacency over as int64 LongTensors (GAT/create_batch.py:433); they carry 32 bits of information
    per element, so each array is packed into a cached pinned int32 staging buffer by all host cores (spk_pack_index_host,
    range-checked: the reference's IndexError comes from here, before any device work) and copied on a copy stream.
    `get(name)` packs + enqueues one array and makes the current stream wait for it, so the caller can interleave: the rows
    travel first and their CSR sort runs on the device while the host is still packing the gather indices and the relation
    ids. Half the PCIe bytes of copying the int64 tensors, and the layout build overlaps both the packing and the transfer."""

    ORDER = ("row", "col", "t1", "t2")

    def __init__(self, edge, edge_type, nhop, e1, e2, max_idx, n_rel, device):
        self.e1, self.e2, self.e, self.device = e1, e2, e1 + e2, device
        self.has2 = e2 > 0
        n_arr = 4 if self.has2 else 3
        st = _STAGING.setdefault(device.index, {})
        if st.get("buf") is None or st["buf"].numel() < n_arr * self.e:
            st["buf"] = torch.empty(n_arr * self.e, dtype=torch.int32).pin_memory()
            st["event"] = None
        if st.get("stream") is None:
            st["stream"] = torch.cuda.Stream(device=device)
        if st.get("event") is not None:
            st["event"].synchronize()                      # the previous build's copies have left the staging buffer
        self.st, self.buf, self.cs = st, st["buf"], st["stream"]
        self.main = torch.cuda.current_stream(device)
        # All destination arrays are allocated NOW, before any of this build's kernels is queued, and the copy stream waits for
        # the work queued so far: memory the caching allocator hands out later could be a block that a just-queued main-stream
        # kernel (e.g. the sort's workspace) still uses, and the copy stream would overwrite it.
        self.dev = {name: torch.empty(self.e, dtype=torch.int32, device=device) for name in self.ORDER[:n_arr]}
        # (Measured alternative, rejected: sending the rows as raw int64 by DMA while the cores pack the next array. The
        # extra 80 MB on the link and the contention for host memory cost more than the packing it hides: 68.8 vs 64.5 ms.)
        self.cs.wait_stream(self.main)
        for t in self.dev.values():
            t.record_stream(self.cs)
        self.edge, self.edge_type = edge.contiguous(), edge_type.contiguous()
        self.nhop = nhop.contiguous() if self.has2 else None
        # name -> (1-hop source pointer, 2-hop column of [s, r1, r2, t], upper bound)
        self.plan = {"row": (self.edge.data_ptr(), 3, max_idx), "col": (self.edge.data_ptr() + 8 * e1, 0, max_idx),
                     "t1": (self.edge_type.data_ptr(), 1, n_rel), "t2": (None, 2, n_rel)}

    def get(self, name, wait=True):
        """Pack + queue the copy of one array; wait=False returns (array, event) and leaves the waiting to the caller."""
        lib = _lib.load()
        k = self.ORDER.index(name)
        src, hop_col, hi = self.plan[name]
        seg = self.buf[k * self.e:(k + 1) * self.e]
        dev_arr = self.dev.pop(name)
        # packed and copied in chunks: the link starts after the first 16 MB instead of after the whole array, so packing and
        # transfer of one array overlap as well (not only the transfer of one array with the packing of the next)
        spans = [(0, self.e1, src, 1)] if self.e1 else []
        if self.has2:
            spans.append((self.e1, self.e, self.nhop.data_ptr() + 8 * hop_col, 4))
        for s0, s1, sp, stride in spans:
            for c0 in range(s0, s1, PACK_CHUNK):
                c1 = min(s1, c0 + PACK_CHUNK)
                if sp is None:
                    seg[c0:c1].fill_(-1)                   # t2 of a 1-hop edge
                    rc = 0
                else:
                    rc = lib.spk_pack_index_host(sp + 8 * stride * (c0 - s0), c1 - c0, stride, 0, hi, seg.data_ptr() + 4 * c0, 0)
                if rc == 5:
                    raise IndexError("edge / relation index out of range for the given entity / relation tables")
                _lib.check(rc, "pack_index_host")
                with torch.cuda.stream(self.cs):
                    dev_arr[c0:c1].copy_(seg[c0:c1], non_blocking=True)
        with torch.cuda.stream(self.cs):
            ev = torch.cuda.Event()
            ev.record(self.cs)
        self.st["event"] = ev
        if not wait:
            return dev_arr, ev
        self.main.wait_event(ev)
        return dev_arr


class KGraph:
    """Device-resident segment layouts of one edge list.

    CSR  (key edge[0]) : rowptr[N+1], col[E], t1[E], t2[E]|None, perm[E] (CSR position -> original edge id),
                         row[E] (aggregation row of each CSR position)
    CSC  (key edge[1]) : colptr[Ncols+1], csc_row[E], csc_pos[E] (CSR position of the edge)
    REL  (key relation): relptr[R+1], rel_row[M], rel_pos[M]   (2-hop edges appear under both relations)
    """

    def __init__(self, edge, edge_type, nhop, n_nodes, n_rel, device=None, n_cols=None, build_backward=True):
        lib = _lib.load()
        if device is None:
            device = edge.device if edge.is_cuda else torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.n_nodes = int(n_nodes)
        self.n_cols = int(n_cols if n_cols is not None else n_nodes)
        self.n_rel = int(n_rel)
        self.e1 = int(edge.shape[1]) if edge.dim() == 2 else 0
        has2 = nhop is not None and nhop.numel() > 0
        self.e2 = int(nhop.shape[0]) if has2 else 0
        e = self.e1 + self.e2
        self.n_edges = e
        i32 = dict(dtype=torch.int32, device=device)
        max_idx = max(self.n_nodes, self.n_cols)
        main = torch.cuda.current_stream(device)
        host = (e > 0 and not edge.is_cuda and not edge_type.is_cuda and (not has2 or not nhop.is_cuda)
                and edge.dtype == torch.int64 and edge_type.dtype == torch.int64 and (not has2 or nhop.dtype == torch.int64))
        stager = None
        if host:
            # host-resident int64 tensors (the reference's calling convention): packed to int32 in pinned memory and copied
            # array by array on a copy stream, each right before its first use, so the device sorts overlap the packing
            stager = _HostStager(edge, edge_type, nhop if has2 else None, self.e1, self.e2, max_idx, self.n_rel, device)
            row = stager.get("row")
            col = t1 = t2 = None
            err = None
        else:
            edge = edge.to(device=device, dtype=torch.int64).contiguous()
            edge_type = edge_type.to(device=device, dtype=torch.int64).contiguous()
            nhop = nhop.to(device=device, dtype=torch.int64).contiguous() if has2 else None
            row = torch.empty(e, **i32); col = torch.empty(e, **i32); t1 = torch.empty(e, **i32)
            t2 = torch.empty(e, **i32) if has2 else None
            err = torch.zeros(1, **i32)
            _lib.check(lib.spk_edges_concat(edge.data_ptr() if self.e1 else None, self.e1,
                                            edge_type.data_ptr() if self.e1 else None,
                                            nhop.data_ptr() if has2 else None, self.e2,
                                            row.data_ptr(), col.data_ptr(), t1.data_ptr(),
                                            t2.data_ptr() if has2 else None, max_idx, self.n_rel,
                                            err.data_ptr(), _lib.stream_ptr()), "edges_concat")
        # CSR: stable sort by aggregation row
        keys, perm = sort_pairs(row, _iota(e, device), _key_bits(max_idx))
        if err is not None and int(err.item()) != 0:
            raise IndexError("edge / relation index out of range for the given entity / relation tables")
        self.row = keys
        self.perm = perm
        self.rowptr = _segment_ptr(keys, self.n_nodes)
        self.col = _gather(stager.get("col") if stager else col, perm)
        self.t1 = _gather(stager.get("t1") if stager else t1, perm)
        self.t2 = (_gather(stager.get("t2") if stager else t2, perm)) if has2 else None
        self.row_hubs = HubSet(self.rowptr)
        self.colptr = self.csc_row = self.csc_pos = self.col_hubs = self.csc_t1 = self.csc_t2 = None
        self.relptr = self.rel_row = self.rel_pos = self.rel_hubs = None
        if build_backward:
            self.build_backward()

    def build_backward(self):
        if self.colptr is not None:
            return
        lib = _lib.load()
        e, device = self.n_edges, self.device
        ckeys, cpos = sort_pairs(self.col.clone(), _iota(e, device), _key_bits(self.n_cols))
        self.colptr = _segment_ptr(ckeys, self.n_cols)
        self.csc_pos = cpos
        self.csc_row = _gather(self.row, cpos)
        self.col_hubs = HubSet(self.colptr)
        self.csc_t1 = _gather(self.t1, cpos)                     # relation ids in CSC order (column-major fused backward)
        self.csc_t2 = _gather(self.t2, cpos) if self.t2 is not None else None
        m = 2 * e if self.t2 is not None else e
        rkeys = torch.empty(m, dtype=torch.int32, device=device)
        rvals = torch.empty(m, dtype=torch.int32, device=device)
        _lib.check(lib.spk_rel_incidence(self.t1.data_ptr(), self.t2.data_ptr() if self.t2 is not None else None,
                                         e, self.n_rel, rkeys.data_ptr(), rvals.data_ptr(), _lib.stream_ptr()),
                   "rel_incidence")
        rkeys, rpos = sort_pairs(rkeys, rvals, _key_bits(self.n_rel + 1))
        self.relptr = _segment_ptr(rkeys, self.n_rel)
        self.rel_pos = rpos
        self.rel_row = _gather(self.row, rpos)
        if REL_WINDOW_ORDER:
            self.rel_hubs = HubSet(self.relptr, chunk=REL_CHUNK, order_by=self.rel_row, order_bits=_key_bits(self.n_nodes))
        else:
            self.rel_hubs = HubSet(self.relptr)

    def to_csr_order(self, per_edge):
        """Reorder a per-edge tensor given in the caller's edge order ([..., E]) into CSR order."""
        return per_edge.index_select(-1, self.perm.long()).contiguous()
