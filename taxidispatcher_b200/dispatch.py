"""The reference's Python-level dispatch surface on top of libtaxidispatch.so (B200, sm_100a).

Host functions mirror the reference's names, argument order and return layouts
(SURVEY.md section 8(b)):

  calculate_cost(distances, demand, cabs)          split.py:123-136, simulate.py:17-33
  solve(n, cost) -> x                              solver.py:11-27   (x[n*cab+cust] in {0,1})
  solve_dispatch(distances, demand, cabs)          split.py:139-155  -> (n, x, cost)
  LCM(n, c, ...)                                   heuristic.py:24-33, split.py:161-175,
                                                   greedy_opt.py:61-82, simulate.py:76-97
  LCM_java(cost)                                   Simulator.java:523-549
  find_pool(...) / find_pool_all(...) / pool_merge(...)   pool_n.c / findpool.c

PyTorch is used only for device buffers, pinned staging and streams.  All arithmetic runs in the
hand-written CUDA kernels behind the C ABI (include/taxidispatch.h).  There is no CPU fallback:
without the built library or without a CUDA device these functions raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import BIG_COST, INT32_MAX, POOL_REC_W, AssignStats, LcmParams, PoolStats, TaxiDispatchError, check

__all__ = ["BIG_COST", "Engine", "calculate_cost", "solve", "solve_dispatch", "solve_full", "solve_assignment", "LCM", "LCM_heuristic",
           "LCM_split", "LCM_greedy_opt", "LCM_simulate", "LCM_java", "find_pool", "find_pool_all", "find_pool_block",
           "find_pool_pairs", "pool_merge",
           "TaxiDispatchError"]


def _require_cuda():
    if not torch.cuda.is_available():
        raise TaxiDispatchError(_lib.TD_ERR_NO_DEVICE, "taxidispatcher_b200",
                                "torch sees no CUDA device; this engine has no CPU fallback")


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_PINNED = {}   # numel -> reusable pinned staging buffer (cudaHostAlloc per call costs more than small copies)
COPIED = {"h2d": 0, "d2h": 0}   # bytes the host-facing functions moved over PCIe so far (bench.py reports the per-call difference)


def _d2h(t: torch.Tensor) -> np.ndarray:
    """Device tensor -> numpy, counted."""
    arr = t.cpu().numpy()
    COPIED["d2h"] += arr.nbytes
    return arr


def _d2h_int(t: torch.Tensor) -> int:
    COPIED["d2h"] += t.element_size()
    return int(t.item())


def _h2d_i32(a, pinned: bool = True) -> torch.Tensor:
    """Host array-like -> int32 CUDA tensor through pinned staging (asynchronous copy on the current stream).
    Staging buffers are reused; the stream is synchronised before a buffer is overwritten again."""
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    COPIED["h2d"] += arr.nbytes
    if not pinned or arr.size == 0:
        return torch.from_numpy(arr).to("cuda")
    if arr.size > (1 << 22):                       # big matrices: one-off pinned copy
        return torch.from_numpy(arr).pin_memory().to("cuda", non_blocking=True)
    bucket = 1 << max(8, int(arr.size - 1).bit_length())        # power-of-two buckets: few allocations
    slot = _PINNED.get(bucket)
    if slot is None:
        slot = {"buf": torch.empty(bucket, dtype=torch.int32).pin_memory(), "ev": None}
        _PINNED[bucket] = slot
    if slot["ev"] is not None:
        slot["ev"].synchronize()                   # the previous copy out of this buffer has finished
    stage = slot["buf"][: arr.size]
    stage.numpy()[:] = arr.reshape(-1)
    t = stage.to("cuda", non_blocking=True).reshape(arr.shape)
    ev = torch.cuda.Event()
    ev.record()
    slot["ev"] = ev
    return t


def _h2d_i32_many(arrays) -> list:
    """Several small host arrays -> int32 CUDA tensors through ONE pinned staging buffer and ONE asynchronous copy (each
    array starts on a 16-byte boundary: the pool kernels stage their tables with bulk copies).  The per-copy cost of a
    few-KB transfer is its fixed latency, not its bytes."""
    arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.int32)) for a in arrays]
    offs, total = [], 0
    for arr in arrs:
        offs.append(total)
        total += (arr.size + 3) & ~3
    if total == 0 or total > (1 << 22) or any(arr.size == 0 for arr in arrs):
        return [_h2d_i32(arr) for arr in arrs]
    COPIED["h2d"] += sum(arr.nbytes for arr in arrs)
    bucket = 1 << max(8, int(total - 1).bit_length())
    slot = _PINNED.get(("many", bucket))
    if slot is None:
        slot = {"buf": torch.empty(bucket, dtype=torch.int32).pin_memory(), "ev": None}
        _PINNED[("many", bucket)] = slot
    if slot["ev"] is not None:
        slot["ev"].synchronize()
    host = slot["buf"].numpy()
    for arr, off in zip(arrs, offs):
        host[off: off + arr.size] = arr.reshape(-1)
    dev = slot["buf"][:total].to("cuda", non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    slot["ev"] = ev
    return [dev[off: off + arr.size].reshape(arr.shape) for arr, off in zip(arrs, offs)]


class Engine:
    """Device-pointer front end: torch CUDA tensors in, torch CUDA tensors out, workspaces cached
    per (operation, size) so that steady-state calls allocate nothing."""

    def __init__(self, device: Optional[int] = None):
        _require_cuda()
        self.lib = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._ws = {}

    # -- scratch --------------------------------------------------------------------------------
    def _workspace(self, key, nbytes: int) -> torch.Tensor:
        """One grow-only scratch buffer per operation (calls are stream-ordered, so reuse across sizes is safe)."""
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = None
            self._ws.pop(key, None)                                   # release before growing
            ws = torch.empty(max(int(nbytes * 5 // 4), 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    # -- K1 ---------------------------------------------------------------------------------------
    def cost_matrix(self, dist: torch.Tensor, cab_to: torch.Tensor, cust_from: torch.Tensor, fill: int = BIG_COST,
                    cutoff: Optional[int] = None, out: Optional[torch.Tensor] = None,
                    rows: Optional[tuple] = None) -> torch.Tensor:
        """Padded n x n matrix, or -- rows=(lo, hi) -- its cab-row block [lo, hi) as a (hi - lo) x n tensor."""
        n_cabs, n_cust = int(cab_to.numel()), int(cust_from.numel())
        n = max(n_cabs, n_cust)
        n_stands = int(dist.shape[0]) if dist.numel() else 0
        lo, hi = (0, n) if rows is None else (int(rows[0]), int(rows[1]))
        if out is None:
            out = torch.empty((hi - lo, n), dtype=torch.int32, device=self.device)
        rc = self.lib.td_cost_matrix_rows(_ptr(dist), n_stands, _ptr(cab_to), n_cabs, _ptr(cust_from), n_cust, int(fill),
                                          -1 if cutoff is None else int(cutoff), lo, hi - lo, _ptr(out), _stream())
        check(rc, "td_cost_matrix_rows")
        return out

    # -- K3 ---------------------------------------------------------------------------------------
    def lcm(self, cost: torch.Tensor, mask_value: int, stop_above: int = INT32_MAX, stop_at_value: int = INT32_MAX,
            sum_below: int = INT32_MAX, residual_size: int = 0, max_iters: int = -1):
        """Returns device tensors (rows[n], cols[n], scal) where scal = int64[2] {total, n_pairs | last_min << 32}.
        Use lcm_host_view() to decode after a synchronisation."""
        n = int(cost.shape[0])
        rows = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
        cols = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
        scal = torch.zeros(4, dtype=torch.int32, device=self.device)  # [total lo, total hi, n_pairs, last_min]
        prm = LcmParams(int(mask_value), int(stop_above), int(stop_at_value), int(sum_below), int(residual_size),
                        int(max_iters))
        nbytes = self.lib.td_lcm_workspace_bytes(n)
        ws = self._workspace("lcm", nbytes)
        base = scal.data_ptr()
        rc = self.lib.td_lcm(_ptr(cost), n, ctypes.byref(prm), _ptr(rows), _ptr(cols), ctypes.c_void_p(base + 8),
                             ctypes.c_void_p(base), ctypes.c_void_p(base + 12), _ptr(ws), ws.numel(), _stream())
        check(rc, "td_lcm")
        return rows, cols, scal

    @staticmethod
    def lcm_host_view(rows, cols, scal):
        s = scal.cpu().numpy()
        total = int(np.frombuffer(s[:2].tobytes(), dtype=np.int64)[0])
        k = int(s[2])
        return {"total": total, "n_pairs": k, "last_min": int(s[3]), "rows": rows[:k].cpu().numpy(),
                "cols": cols[:k].cpu().numpy()}

    # -- K2 ---------------------------------------------------------------------------------------
    def assign(self, cost: torch.Tensor, want_x: bool = False, want_stats: bool = False,
               n_real_rows: Optional[int] = None, n_real_cols: Optional[int] = None):
        """Exact optimum of the padded n x n matrix.  n_real_rows / n_real_cols (one of them == n) name the real block of
        an unbalanced instance: the constant padding rows / columns are then not searched (td_assign_exact_rect); the
        objective and the padded output layout are the same either way."""
        n = int(cost.shape[0])
        nr = n if n_real_rows is None else int(n_real_rows)
        nc = n if n_real_cols is None else int(n_real_cols)
        col = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
        obj = torch.zeros(1, dtype=torch.int64, device=self.device)
        x = torch.empty(n * n, dtype=torch.uint8, device=self.device) if want_x else None
        nbytes = self.lib.td_assign_rect_workspace_bytes(n, nr, nc)
        ws = self._workspace("assign", nbytes)
        st = AssignStats() if want_stats else None
        rc = self.lib.td_assign_exact_rect(_ptr(cost), n, nr, nc, _ptr(col), _ptr(obj), _ptr(x),
                                           ctypes.byref(st) if st is not None else None, _ptr(ws), ws.numel(), _stream())
        check(rc, "td_assign_exact_rect")
        return col[:n], obj, x, st

    def assign_duals(self, n: int, n_real_rows: Optional[int] = None, n_real_cols: Optional[int] = None):
        """Dual potentials (u per cab row, v per customer column; int64 device tensors) of the LAST assign() call."""
        nr = n if n_real_rows is None else int(n_real_rows)
        nc = n if n_real_cols is None else int(n_real_cols)
        u = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        v = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        ws = self._ws["assign"]
        check(self.lib.td_assign_read_duals(_ptr(ws), n, nr, nc, _ptr(u), _ptr(v), _stream()), "td_assign_read_duals")
        return u[:n], v[:n]

    def assign_certify(self, cost: torch.Tensor, col: torch.Tensor, u: torch.Tensor, v: torch.Tensor,
                       n_real_rows: Optional[int] = None, n_real_cols: Optional[int] = None) -> dict:
        """One sweep over the cost matrix: the complementary-slackness certificate of (col, u, v).  Returns the
        td_assign_certificate fields plus `optimal` (all conditions hold)."""
        n = int(cost.shape[0])
        nr = n if n_real_rows is None else int(n_real_rows)
        nc = n if n_real_cols is None else int(n_real_cols)
        cert = torch.zeros(ctypes.sizeof(_lib.AssignCertificate), dtype=torch.uint8, device=self.device)
        ws = self._workspace("certify", self.lib.td_assign_certify_workspace_bytes(n))
        check(self.lib.td_assign_certify(_ptr(cost), n, nr, nc, _ptr(col), _ptr(u), _ptr(v), _ptr(cert), _ptr(ws), ws.numel(),
                                         _stream()), "td_assign_certify")
        c = _lib.AssignCertificate.from_buffer_copy(cert.cpu().numpy().tobytes())
        out = {f: getattr(c, f) for f, _ in _lib.AssignCertificate._fields_ if f != "reserved"}
        out["optimal"] = (c.min_reduced_cost >= 0 and c.max_matched_slack == 0 and c.sign_violations == 0 and
                          c.dual_objective == c.matched_real_cost)
        return out

    # -- K4 ---------------------------------------------------------------------------------------
    def pool_find(self, demand: torch.Tensor, dist: torch.Tensor, pool_size: int, shard: int = 0, n_shards: int = 8,
                  max_feasible: Optional[int] = None, out: Optional[torch.Tensor] = None, want_stats: bool = True,
                  cnt_out: Optional[torch.Tensor] = None):
        """out: [cap, 9] int32 device rows (cap >= n // pool_size); cnt_out: 1-element int32 device view that
        receives the number of surviving plans (-1: the record list overflowed, only when want_stats=False)."""
        n = int(demand.shape[0])
        n_stands = int(dist.shape[0])
        cap = n // 2 + 1
        if out is None:
            out = torch.empty((cap, POOL_REC_W), dtype=torch.int32, device=self.device)
        cnt = cnt_out if cnt_out is not None else torch.zeros(1, dtype=torch.int32, device=self.device)
        mf = int(max_feasible) if max_feasible is not None else self._ws.get(("pool_mf", n, pool_size), 1 << 21)
        for _ in range(8):
            nbytes = self.lib.td_pool_workspace_bytes(n, n_stands, pool_size, mf)
            ws = self._workspace("pool", nbytes)
            st = PoolStats()
            rc = self.lib.td_pool_find(_ptr(demand), n, _ptr(dist), n_stands, pool_size, shard, n_shards, _ptr(out),
                                       int(out.shape[0]), _ptr(cnt), ctypes.byref(st) if want_stats else None,
                                       _ptr(ws), ws.numel(), mf, _stream())
            if rc == _lib.TD_ERR_CAPACITY and want_stats and st.feasible > mf:
                mf = int(st.feasible) + 1024          # grow the materialised list and retry
                self._ws[("pool_mf", n, pool_size)] = mf
                continue
            check(rc, "td_pool_find")
            return out, cnt, (st if want_stats else None)
        raise TaxiDispatchError(_lib.TD_ERR_CAPACITY, "td_pool_find")

    def pool_find_shards(self, demand: torch.Tensor, dist: torch.Tensor, pool_size: int, shard_begin: int = 0,
                         shard_count: int = 8, n_shards: int = 8, max_feasible: Optional[int] = None,
                         out: Optional[torch.Tensor] = None, counts_out: Optional[torch.Tensor] = None,
                         want_stats: bool = True, defer_stats: bool = False):
        """Consecutive logical shards in one call (one enumeration + one selection launch for all of them).
        out: [shard_count, cap, 9]; counts_out: [shard_count] int32 (device).  Returns (out, counts, stats list).
        defer_stats=True launches asynchronously (single pass, no host round trip) and returns a token for
        pool_read_stats() instead of the stats."""
        n = int(demand.shape[0])
        n_stands = int(dist.shape[0])
        cap = n // 2 + 1
        if out is None:
            out = torch.empty((shard_count, cap, POOL_REC_W), dtype=torch.int32, device=self.device)
        cnt = counts_out if counts_out is not None else torch.zeros(shard_count, dtype=torch.int32, device=self.device)
        key = ("pool_mf", n, pool_size, shard_count)
        mf = int(max_feasible) if max_feasible is not None else self._ws.get(key, (1 << 21) * min(shard_count, 4))
        for _ in range(8):
            nbytes = self.lib.td_pool_shards_workspace_bytes(n, n_stands, pool_size, shard_count, mf)
            ws = self._workspace("pool", nbytes)
            st = (PoolStats * shard_count)()
            rc = self.lib.td_pool_find_shards(_ptr(demand), n, _ptr(dist), n_stands, pool_size, shard_begin, shard_count,
                                              n_shards, _ptr(out), int(out.shape[1]), _ptr(cnt),
                                              st if (want_stats and not defer_stats) else None, _ptr(ws), ws.numel(), mf,
                                              _stream())
            if rc == _lib.TD_ERR_CAPACITY and want_stats:
                need = sum(int(s.feasible) for s in st)
                if need > mf:
                    mf = need + 1024                      # grow the materialised list and retry
                    self._ws[key] = mf
                    continue
            check(rc, "td_pool_find_shards")
            if want_stats and not defer_stats and max_feasible is None and st[0].passes > 1:
                # the feasible set did not fit the record list in one pass: remember a capacity that does (bounded
                # by ~4 GiB per list) so that later calls -- in particular asynchronous ones, which cannot fall back
                # to cost windows -- run single-pass
                need = sum(int(s.feasible) for s in st) + 1024
                self._ws[key] = max(mf, min(need, 1 << 28))
            if defer_stats:
                return out, cnt, (ws, shard_count)
            return out, cnt, (list(st) if want_stats else None)
        raise TaxiDispatchError(_lib.TD_ERR_CAPACITY, "td_pool_find_shards")

    def pool_find_shards_headed(self, demand: torch.Tensor, dist: torch.Tensor, pool_size: int, shard_begin: int,
                                shard_count: int, n_shards: int, out: torch.Tensor, max_feasible: Optional[int] = None):
        """Asynchronous single pass into headed blocks: out is [shard_count, cap + 1, 9] int32 (device), row 0 of every
        block = {count, evaluated lo / hi, feasible lo / hi, ...} (td_pool_find_shards_headed).  Needs a record capacity
        that holds every feasible plan -- remembered by an earlier synchronous pool_find_shards call of the same shape."""
        n = int(demand.shape[0])
        n_stands = int(dist.shape[0])
        cap = int(out.shape[1]) - 1
        key = ("pool_mf", n, pool_size, shard_count)
        mf = int(max_feasible) if max_feasible is not None else self._ws.get(key, (1 << 21) * min(shard_count, 4))
        nbytes = self.lib.td_pool_shards_workspace_bytes(n, n_stands, pool_size, shard_count, mf)
        ws = self._workspace("pool", nbytes)
        rc = self.lib.td_pool_find_shards_headed(_ptr(demand), n, _ptr(dist), n_stands, pool_size, shard_begin, shard_count,
                                                 n_shards, _ptr(out), cap, _ptr(ws), ws.numel(), mf, _stream())
        check(rc, "td_pool_find_shards_headed")
        return out

    def pool_merge_headed(self, blocks: torch.Tensor, slot_shard: Optional[torch.Tensor], n: int, pool_size: int):
        """blocks [n_slots, cap + 1, 9] headed blocks (device).  Returns (plans, count) device tensors."""
        n_slots, cap = int(blocks.shape[0]), int(blocks.shape[1]) - 1
        total = n_slots * cap
        out = torch.empty((max(total, 1), POOL_REC_W), dtype=torch.int32, device=self.device)
        cnt = torch.zeros(1, dtype=torch.int32, device=self.device)
        ws = self._workspace("merge", self.lib.td_pool_merge_workspace_bytes(total, n))
        rc = self.lib.td_pool_merge_headed(_ptr(blocks), _ptr(slot_shard), n_slots, cap, n, pool_size, _ptr(out), _ptr(cnt),
                                           _ptr(ws), ws.numel(), _stream())
        check(rc, "td_pool_merge_headed")
        return out, cnt

    def pool_merge_headed_packed(self, blocks: torch.Tensor, slot_shard: Optional[torch.Tensor], n: int, pool_size: int):
        """Merge of headed blocks with everything the host wants in ONE buffer: row 0 = {merged count}, rows 1..n//pool_size
        = the merged plans, then one header row per block.  One device->host copy completes a whole `findpool` job."""
        n_slots, cap = int(blocks.shape[0]), int(blocks.shape[1]) - 1
        total = n_slots * cap
        keep = n // pool_size + 1                                     # merged plans are customer-disjoint
        pack = torch.empty((1 + max(total, keep) + n_slots, POOL_REC_W), dtype=torch.int32, device=self.device)
        ws = self._workspace("merge", self.lib.td_pool_merge_workspace_bytes(total, n))
        rc = self.lib.td_pool_merge_headed(_ptr(blocks), _ptr(slot_shard), n_slots, cap, n, pool_size,
                                           ctypes.c_void_p(pack.data_ptr() + 4 * POOL_REC_W), _ptr(pack), _ptr(ws), ws.numel(),
                                           _stream())
        check(rc, "td_pool_merge_headed")
        pack[1 + keep: 1 + keep + n_slots] = blocks[:, 0, :]          # the headers ride along (a few hundred bytes)
        host = _d2h(pack[: 1 + keep + n_slots])
        m = int(host[0, 0])
        counts, ev, fe = self.headers_host_view(host[1 + keep:])
        return host[1: 1 + m].copy(), counts, ev, fe

    @staticmethod
    def headers_host_view(headers: np.ndarray):
        """headers: [n_slots, 9] int32 header rows -> (counts, evaluated, feasible) int64 arrays"""
        h = np.asarray(headers).astype(np.int64) & 0xFFFFFFFF
        counts = np.asarray(headers)[:, 0].astype(np.int64)
        return counts, h[:, 1] | (h[:, 2] << 32), h[:, 3] | (h[:, 4] << 32)

    def pool_read_stats(self, token):
        """Completes a defer_stats=True call: waits for the stream, returns (stats list, overflowed)."""
        ws, shard_count = token
        st = (PoolStats * shard_count)()
        ov = ctypes.c_int(0)
        check(self.lib.td_pool_read_stats(_ptr(ws), shard_count, st, ctypes.byref(ov), _stream()), "td_pool_read_stats")
        COPIED["d2h"] += int(self.lib.td_pool_read_stats_bytes())
        return list(st), bool(ov.value)

    def pool_pairs(self, frm: torch.Tensor, to: torch.Tensor, dist: torch.Tensor, accept_all: bool = True,
                   max_loss: float = 1.01):
        """Simulator.findPool: returns (pairs [cap,4] = custA, custB, plan, cost; count) device tensors."""
        n = int(frm.numel())
        cap = n // 2 + 1
        out = torch.empty((cap, 4), dtype=torch.int32, device=self.device)
        cnt = torch.zeros(1, dtype=torch.int32, device=self.device)
        nbytes = self.lib.td_pool_pairs_workspace_bytes(n)
        ws = self._workspace("pairs", nbytes)
        rc = self.lib.td_pool_pairs(_ptr(frm), _ptr(to), n, _ptr(dist), int(dist.shape[0]), 1 if accept_all else 0,
                                    float(max_loss), _ptr(out), cap, _ptr(cnt), _ptr(ws), ws.numel(), _stream())
        check(rc, "td_pool_pairs")
        return out, cnt

    def pool_merge(self, shard_plans: torch.Tensor, total: int, n: int, pool_size: int):
        out = torch.empty((max(total, 1), POOL_REC_W), dtype=torch.int32, device=self.device)
        cnt = torch.zeros(1, dtype=torch.int32, device=self.device)
        nbytes = self.lib.td_pool_merge_workspace_bytes(total, n)
        ws = self._workspace("merge", nbytes)
        rc = self.lib.td_pool_merge(_ptr(shard_plans), total, n, pool_size, _ptr(out), _ptr(cnt), _ptr(ws), ws.numel(),
                                    _stream())
        check(rc, "td_pool_merge")
        return out, cnt


    def pool_merge_padded(self, slot_plans: torch.Tensor, slot_counts: torch.Tensor, slot_shard: Optional[torch.Tensor],
                          n: int, pool_size: int):
        """slot_plans [n_slots, cap, 9], slot_counts [n_slots] (device).  Returns (plans, count) device tensors."""
        n_slots, cap = int(slot_plans.shape[0]), int(slot_plans.shape[1])
        total = n_slots * cap
        out = torch.empty((max(total, 1), POOL_REC_W), dtype=torch.int32, device=self.device)
        cnt = torch.zeros(1, dtype=torch.int32, device=self.device)
        nbytes = self.lib.td_pool_merge_workspace_bytes(total, n)
        ws = self._workspace("merge", nbytes)
        rc = self.lib.td_pool_merge_padded(_ptr(slot_plans), _ptr(slot_counts), _ptr(slot_shard), n_slots, cap, n, pool_size,
                                           _ptr(out), _ptr(cnt), _ptr(ws), ws.numel(), _stream())
        check(rc, "td_pool_merge_padded")
        return out, cnt


_engine: Optional[Engine] = None


def engine() -> Engine:
    global _engine
    if _engine is None:
        _engine = Engine()
    return _engine


# ================================================================================================
# reference-shaped host API
# ================================================================================================
def _cost_device(distances, demand: Sequence, cabs: Sequence, fill: int, cutoff: Optional[int]):
    """Device-resident cost matrix of calculate_cost (None when n == 0)."""
    n_cabs, n_cust = len(cabs), len(demand)
    n = max(n_cabs, n_cust)
    if n == 0:
        return 0, None
    eng = engine()
    dist = _h2d_i32(distances)
    cab_to = _h2d_i32([c[2] for c in cabs]) if n_cabs else torch.empty(0, dtype=torch.int32, device=eng.device)
    cust_from = _h2d_i32([d[1] for d in demand]) if n_cust else torch.empty(0, dtype=torch.int32, device=eng.device)
    if n_cabs and n_cust:
        s = int(dist.shape[0])
        if int(cab_to.max()) >= s or int(cab_to.min()) < 0 or int(cust_from.max()) >= s or int(cust_from.min()) < 0:
            raise IndexError("stand index out of range of the distance table")  # the reference raises IndexError too
    return n, eng.cost_matrix(dist, cab_to, cust_from, fill, cutoff)


def calculate_cost(distances, demand: Sequence, cabs: Sequence, fill: int = BIG_COST, cutoff: Optional[int] = None,
                   as_list: bool = False):
    """split.py:123-136: (n, cost) with cost[c_idx][d_idx] = distances[cab.to][customer.from],
    square-padded with `fill` (big_cost).  cutoff=DROP_TIME gives simulate.py:27 / Simulator.java:509.
    demand / cabs are lists of (id, from, to); indexing is positional.  n == 0 -> (0, 0)
    (simulate.py:21).  cost is an int32 ndarray (list of rows with as_list=True)."""
    n, cost_d = _cost_device(distances, demand, cabs, fill, cutoff)
    if n == 0:
        return 0, 0
    cost = cost_d.cpu().numpy()
    return n, (cost.tolist() if as_list else cost)


def solve_full(n: int, cost):
    """Exact optimum; returns (x, col_of_row, objective, stats).  x is the reference vector."""
    if n == 0:
        return np.zeros(0, np.uint8), np.zeros(0, np.int32), 0, None
    eng = engine()
    c = _h2d_i32(np.asarray(cost).reshape(n, n))
    col, obj, x, st = eng.assign(c, want_x=True, want_stats=True)
    return x.cpu().numpy(), col.cpu().numpy(), int(obj.item()), st


def solve(n: int, cost):
    """solver.py:11-27: `x` with x[n*cab + cust] == 1 for the chosen cells.  n == 0 -> (0, [])
    exactly like solver.py:12."""
    if n == 0:
        return 0, []
    return solve_full(n, cost)[0]


def solve_dispatch(distances, demand, cabs, fill: int = BIG_COST, cutoff: Optional[int] = None):
    """split.py:139-155 / simulate.py:36-53: (n, x, cost).  n == 0 -> (0, [], 0) (simulate.py:38).
    The cost matrix goes from K1 to K2 on the device; the host copy is only the returned value."""
    n, cost_d = _cost_device(distances, demand, cabs, fill, cutoff)
    if n == 0:
        return 0, [], 0
    # every real call is unbalanced (dummy rows / columns of `fill`): only the real block is searched
    _, obj, x, _ = engine().assign(cost_d, want_x=True, n_real_rows=max(len(cabs), 1) if len(cabs) < n else n,
                                   n_real_cols=max(len(demand), 1) if len(demand) < n else n)
    if int(obj.item()) == -(1 << 63):              # the kernel withdraws its result when it gave up (td_assign.cu)
        raise TaxiDispatchError(_lib.TD_ERR_NOT_CONVERGED, "td_assign_exact_rect")
    return n, x.cpu().numpy(), cost_d.cpu().numpy()


def solve_assignment(dist_d: torch.Tensor, cab_to, cust_from, fill: int = BIG_COST, cutoff: Optional[int] = None):
    """K1 + K2 for callers that only need the matching (split.py's range solves, the Simulator): stand indices in,
    (n, col_of_row[n], cost_of_row[n]) out as small host arrays -- the n x n matrix and the n^2 solution vector never
    leave the device.  dist_d is the stand table already on the device (it is shared by every range of a split).
    cost_of_row[i] = cost[i][col_of_row[i]] (== fill for a dummy pairing, split.py:89)."""
    n_cabs, n_cust = len(cab_to), len(cust_from)
    n = max(n_cabs, n_cust)
    if n == 0:
        return 0, np.zeros(0, np.int32), np.zeros(0, np.int32)
    eng = engine()
    cab_d = _h2d_i32(cab_to) if n_cabs else torch.empty(0, dtype=torch.int32, device=eng.device)
    cust_d = _h2d_i32(cust_from) if n_cust else torch.empty(0, dtype=torch.int32, device=eng.device)
    cost_d = eng.cost_matrix(dist_d, cab_d, cust_d, fill, cutoff)
    col, _, _, _ = eng.assign(cost_d, n_real_rows=max(n_cabs, 1) if n_cabs < n else n,
                              n_real_cols=max(n_cust, 1) if n_cust < n else n)
    col_h = _d2h(col)
    if n and int(col_h.min()) < 0:                 # the kernel withdraws the matching when it gave up (td_assign.cu)
        raise TaxiDispatchError(_lib.TD_ERR_NOT_CONVERGED, "td_assign_exact_rect")
    picked = cost_d.gather(1, col.to(torch.int64).unsqueeze(1)).squeeze(1)      # n cells: plumbing, not compute
    return n, col_h, _d2h(picked)


def LCM(n: int, c, mask_value: int = BIG_COST, stop_above: int = INT32_MAX, stop_at_value: int = INT32_MAX,
        sum_below: int = INT32_MAX, residual_size: int = 0, max_iters: int = -1):
    """Generic LCM; c is an n x n matrix (or flat n*n) whose row-major flattening is cost[cab][cust].
    Returns dict(total, n_pairs, rows, cols, last_min)."""
    if n == 0:
        return {"total": 0, "n_pairs": 0, "rows": np.zeros(0, np.int32), "cols": np.zeros(0, np.int32),
                "last_min": INT32_MAX}
    eng = engine()
    cost = _h2d_i32(np.asarray(c).reshape(n, n))
    return Engine.lcm_host_view(*eng.lcm(cost, mask_value, stop_above, stop_at_value, sum_below, residual_size, max_iters))


def LCM_heuristic(n, c, mask: int = 100):
    """heuristic.py:24-33 -> total_cost (mask 100, every value summed)."""
    return LCM(n, c, mask_value=mask)["total"]


def LCM_split(n, c, big_cost: int = BIG_COST):
    """split.py:161-175 -> total_cost (mask big_cost, dummy costs not summed)."""
    return LCM(n, c, mask_value=big_cost, sum_below=big_cost)["total"]


def LCM_greedy_opt(n, c, threshold: int = 10, big_cost: int = BIG_COST):
    """greedy_opt.py:61-82 -> (total_cost, allocated_supply, allocated_demand)."""
    r = LCM(n, c, mask_value=big_cost, stop_above=threshold, sum_below=big_cost)
    return r["total"], r["rows"].tolist(), r["cols"].tolist()


def LCM_simulate(n, c, threshold: int = 20, big_cost: int = BIG_COST):
    """simulate.py:76-97 -> (total_cost, allocated_supply, allocated_demand, allocated)."""
    tot, sup, dem = LCM_greedy_opt(n, c, threshold, big_cost)
    return tot, sup, dem, list(zip(sup, dem))


def LCM_java(cost, big_cost: int = BIG_COST, max_non_lcm: int = 600):
    """Simulator.java:523-549 -> (pairs, LCM_min_val)."""
    cost = np.asarray(cost)
    n = cost.shape[0]
    if n == 0:
        return [], big_cost
    r = LCM(n, cost, mask_value=big_cost, stop_at_value=big_cost, residual_size=max_non_lcm)
    mn = r["last_min"]
    return list(zip(r["rows"].tolist(), r["cols"].tolist())), (big_cost if mn >= big_cost else mn)


def find_pool(demand, dist, pool_size: int, shard: int = 0, n_shards: int = 8):
    """One pool_n process (pool_n.c:209-238): returns (plans [m,9] int32 in the record layout of
    pool_n.c:123-134, stats dict evaluated/feasible/kept)."""
    eng = engine()
    dem = _h2d_i32(np.asarray(demand, dtype=np.int32).reshape(-1, 5))
    d = _h2d_i32(dist)
    out, cnt, st = eng.pool_find(dem, d, pool_size, shard, n_shards)
    m = int(cnt.item())
    return out[:m].cpu().numpy(), {"evaluated": st.evaluated, "feasible": st.feasible, "kept": st.kept,
                                   "rounds": st.rounds, "passes": st.passes}


def find_pool_block(demand, dist, pool_size: int, shard_begin: int, shard_count: int, n_shards: int = 8):
    """Consecutive logical shards in one device call.  Returns [(plans, stats dict)] per shard."""
    if shard_count <= 0:
        return []
    eng = engine()
    dem = _h2d_i32(np.asarray(demand, dtype=np.int32).reshape(-1, 5))
    d = _h2d_i32(dist)
    out, cnt, st = eng.pool_find_shards(dem, d, pool_size, shard_begin, shard_count, n_shards)
    counts = cnt.cpu().numpy()
    plans = out.cpu().numpy()
    return [(plans[s, : int(counts[s])].copy(),
             {"evaluated": st[s].evaluated, "feasible": st[s].feasible, "kept": st[s].kept, "rounds": st[s].rounds,
              "passes": st[s].passes}) for s in range(shard_count)]


def find_pool_pairs(frm, to, dist, accept_all: bool = True, max_loss: float = 1.01):
    """Simulator.findPool (Simulator.java:681-758): list of (custA, custB, plan, cost) in scan order;
    custA / custB are positions in the given arrays; from < 0 marks a removed row."""
    n = len(frm)
    if n < 2:
        return np.zeros((0, 4), np.int32)
    eng = engine()
    out, cnt = eng.pool_pairs(_h2d_i32(frm), _h2d_i32(to), _h2d_i32(dist), accept_all, max_loss)
    return out[: int(cnt.item())].cpu().numpy()


def pool_merge(shard_plans, n: int, pool_size: int):
    """findpool.c:83-108 on shard outputs given in shard order."""
    eng = engine()
    parts = [np.asarray(p, dtype=np.int32).reshape(-1, POOL_REC_W) for p in shard_plans]
    allp = np.concatenate(parts, axis=0) if parts else np.zeros((0, POOL_REC_W), np.int32)
    if len(allp) == 0:
        return allp
    out, cnt = eng.pool_merge(_h2d_i32(allp), len(allp), n, pool_size)
    return out[: int(cnt.item())].cpu().numpy()


class PoolJobGraph:
    """The steady-state `findpool` job of one shape as ONE CUDA graph: host->device copy of the demand and the stand table
    out of a pinned buffer, enumeration + selection of this rank's shards into headed blocks, the merge, and the
    device->host copy of the packed result into a pinned buffer.  A call is then: fill the pinned input, one graph launch,
    one stream synchronisation -- the ~10 launches and copies of the job cost one submission (at 8 GPUs the job takes
    0.45 ms on the devices: queuing it piece by piece from Python takes longer than running it).  When the job is spread
    over ranks the all_gather of the blocks is issued between two graphs (front: copy-in + shards, back: merge + copy-out).  Built by find_pool_all / parallel.find_pool_sharded once the asynchronous path has
    succeeded for the shape (the record capacity is known then); an overflow (count -1) makes the caller drop the graph
    and take the cost-window path."""

    def __init__(self, eng: "Engine", n: int, n_stands: int, pool_size: int, n_shards: int, shard_begin: int,
                 shard_count: int, gather=None, slots: Optional[int] = None, slot_shard: Optional[torch.Tensor] = None):
        self.eng, self.n, self.S, self.k = eng, n, n_stands, pool_size
        dev = eng.device
        cap = n // 2 + 1
        self.cap = cap
        self.dem_words = n * 5
        self.dist_off = (self.dem_words + 3) & ~3                      # 16-byte boundary
        total_in = self.dist_off + n_stands * n_stands
        self.h_in = torch.empty(total_in, dtype=torch.int32).pin_memory()
        self.d_in = torch.empty(total_in, dtype=torch.int32, device=dev)
        slots = shard_count if slots is None else slots
        world_slots = slots if gather is None else gather[1] * slots
        self.blocks = torch.zeros((slots, cap + 1, POOL_REC_W), dtype=torch.int32, device=dev)
        self.all_blocks = self.blocks if gather is None else torch.zeros((world_slots, cap + 1, POOL_REC_W), dtype=torch.int32, device=dev)
        self.keep = n // pool_size + 1
        self.n_slots = world_slots
        total = world_slots * cap
        self.rows = 1 + self.keep + world_slots
        self.pack = torch.zeros((1 + max(total, self.keep) + world_slots, POOL_REC_W), dtype=torch.int32, device=dev)
        self.h_out = torch.empty((self.rows, POOL_REC_W), dtype=torch.int32).pin_memory()
        self.in_bytes = (self.dem_words + n_stands * n_stands) * 4
        self.out_bytes = self.rows * POOL_REC_W * 4
        dem_d = self.d_in[: self.dem_words].reshape(n, 5)
        dist_d = self.d_in[self.dist_off:].reshape(n_stands, n_stands)
        ws_m = eng._workspace("merge", eng.lib.td_pool_merge_workspace_bytes(total, n))
        torch.cuda.synchronize()
        self.gather = gather
        self.graph = torch.cuda.CUDAGraph()
        self.graph_post = None

        def front():
            self.d_in.copy_(self.h_in, non_blocking=True)
            if shard_count > 0:
                eng.pool_find_shards_headed(dem_d, dist_d, pool_size, shard_begin, shard_count, n_shards,
                                            out=self.blocks[:shard_count])

        def back():
            rc = eng.lib.td_pool_merge_headed(_ptr(self.all_blocks), _ptr(slot_shard), world_slots, cap, n, pool_size,
                                              ctypes.c_void_p(self.pack.data_ptr() + 4 * POOL_REC_W), _ptr(self.pack),
                                              _ptr(ws_m), ws_m.numel(), _stream())
            check(rc, "td_pool_merge_headed")
            self.pack[1 + self.keep: 1 + self.keep + world_slots] = self.all_blocks[:, 0, :]
            self.h_out.copy_(self.pack[: self.rows], non_blocking=True)

        if gather is None:
            with torch.cuda.graph(self.graph):
                front()
                back()
        else:
            # the collective stays OUTSIDE the graphs (a captured NCCL kernel keeps the communicator busy at teardown):
            # graph - all_gather - graph, three submissions per job
            with torch.cuda.graph(self.graph):
                front()
            self.graph_post = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_post):
                back()
        self._slot_shard = slot_shard                                    # keep alive: the graph holds its address
        self._ws_ptrs = (eng._ws["pool"].data_ptr() if shard_count > 0 else 0, ws_m.data_ptr())

    def valid(self) -> bool:
        """The graph holds raw pointers into the engine's grow-only workspaces: a call of another shape may have
        reallocated them, after which the graph must not run again."""
        ws_p, ws_m = self.eng._ws.get("pool"), self.eng._ws.get("merge")
        return (self._ws_ptrs[0] == 0 or (ws_p is not None and ws_p.data_ptr() == self._ws_ptrs[0])) and \
               ws_m is not None and ws_m.data_ptr() == self._ws_ptrs[1]

    def run(self, dem_np: np.ndarray, dist_np: np.ndarray):
        """(merged plans, counts per slot, evaluated per slot, feasible per slot) -- like Engine.pool_merge_headed_packed"""
        host = self.h_in.numpy()
        host[: self.dem_words] = dem_np.reshape(-1)
        host[self.dist_off:] = dist_np.reshape(-1)
        COPIED["h2d"] += self.in_bytes
        self.graph.replay()
        if self.graph_post is not None:
            self.gather[0](self.all_blocks, self.blocks)
            self.graph_post.replay()
        torch.cuda.current_stream().synchronize()
        COPIED["d2h"] += self.out_bytes
        out = self.h_out.numpy()
        m = int(out[0, 0])
        counts, ev, fe = Engine.headers_host_view(out[1 + self.keep:])
        return out[1: 1 + m].copy(), counts, ev, fe


_POOL_GRAPHS = {}          # shape key -> PoolJobGraph, or a success counter before the graph is built


def release_pool_graphs():
    """Drops every captured job graph (they hold raw pointers into the engine's workspaces)."""
    if _POOL_GRAPHS:
        _POOL_GRAPHS.clear()
        if torch.cuda.is_available():
            torch.cuda.synchronize()


def _capture_pool_job(*args, **kwargs):
    """PoolJobGraph, or a counter value that never reaches the capture threshold again when the capture fails (the
    piecewise path keeps serving the shape; the result of the current call has been computed already)."""
    try:
        return PoolJobGraph(*args, **kwargs)
    except Exception:                                     # noqa: BLE001 -- any capture failure means "no graph", never "no result"
        return -(1 << 30)


def _pool_graphs_enabled() -> bool:
    return os.environ.get("TD_POOL_GRAPH", "1") != "0"


def find_pool_all(demand, dist, pool_size: int, n_shards: int = 8):
    """What `findpool` produces (findpool.c:122-176): all logical shards, merged in shard order."""
    eng = engine()
    dem_np = np.asarray(demand, dtype=np.int32).reshape(-1, 5)
    n = dem_np.shape[0]
    stats = {"evaluated": 0, "feasible": 0, "kept_per_shard": [], "rounds": 0}
    fast_key = ("pool_single_pass_ok", n, pool_size, n_shards)
    dist_np = np.ascontiguousarray(np.asarray(dist, dtype=np.int32))
    gkey = ("all", n, dist_np.shape[0], pool_size, n_shards, eng.device.index)
    job = _POOL_GRAPHS.get(gkey)
    if isinstance(job, PoolJobGraph) and not job.valid():
        _POOL_GRAPHS.pop(gkey, None)
        job = None
    if isinstance(job, PoolJobGraph):
        # steady state of a repeated shape: the whole job is one graph launch (PoolJobGraph)
        plans, counts, ev, fe = job.run(dem_np, dist_np)
        if int(counts.min()) >= 0:
            stats.update(evaluated=int(ev.sum()), feasible=int(fe.sum()), kept_per_shard=[int(c) for c in counts], kept=len(plans))
            return plans, stats
        _POOL_GRAPHS.pop(gkey, None)                      # the input outgrew the record list
        eng._ws[fast_key] = False
    dem, d = _h2d_i32_many([dem_np, dist_np])             # one pinned staging buffer, one copy
    if n_shards <= 64 and eng._ws.get(fast_key):
        # steady state: enumeration, selection and merge are queued back to back into headed blocks; ONE device->host copy
        # (merged plans + count + the per-shard counters) and one synchronisation complete the job
        cap = n // 2 + 1
        bkey = ("pool_blocks", n, n_shards)
        blocks = eng._ws.get(bkey)
        if blocks is None:
            blocks = eng._ws[bkey] = torch.zeros((n_shards, cap + 1, POOL_REC_W), dtype=torch.int32, device=eng.device)
        eng.pool_find_shards_headed(dem, d, pool_size, 0, n_shards, n_shards, out=blocks)
        plans, counts, ev, fe = eng.pool_merge_headed_packed(blocks, None, n, pool_size)
        if int(counts.min()) >= 0:
            stats.update(evaluated=int(ev.sum()), feasible=int(fe.sum()), kept_per_shard=[int(c) for c in counts],
                         kept=len(plans))
            if _pool_graphs_enabled() and n > 0 and dist_np.ndim == 2:
                seen = _POOL_GRAPHS.get(gkey, 0) + 1      # second success of the shape: capture the job
                _POOL_GRAPHS[gkey] = seen if seen < 2 else _capture_pool_job(eng, n, dist_np.shape[0], pool_size, n_shards, 0, n_shards)
            return plans, stats
        eng._ws[fast_key] = False                         # input outgrew the record list: cost windows below
        _POOL_GRAPHS.pop(gkey, None)
    parts, counts = [], []
    single = True
    for b in range(0, n_shards, 64):                       # one call serves up to 64 consecutive shards
        cnt_sh = min(64, n_shards - b)
        out, cnt, st = eng.pool_find_shards(dem, d, pool_size, b, cnt_sh, n_shards)
        parts.append(out)
        counts.append(cnt)
        single = single and st[0].passes == 1
        stats["evaluated"] += sum(int(s.evaluated) for s in st)
        stats["feasible"] += sum(int(s.feasible) for s in st)
        stats["rounds"] += int(st[0].rounds)
        stats["kept_per_shard"] += [int(s.kept) for s in st]
    # next time: asynchronous single pass, provided the remembered record capacity can hold every feasible plan
    eng._ws[fast_key] = n_shards <= 64 and (single or eng._ws.get(("pool_mf", n, pool_size, n_shards), 0) > stats["feasible"])
    slot_plans = parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)
    slot_counts = counts[0] if len(counts) == 1 else torch.cat(counts, dim=0)
    merged, cnt = eng.pool_merge_padded(slot_plans, slot_counts, None, n, pool_size)
    m = _d2h_int(cnt)
    stats["kept"] = m
    return _d2h(merged[:m]), stats
