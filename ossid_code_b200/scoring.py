"""Multi-object, multi-GPU frame scoring on top of the C ABI.

The reference scores one (object, frame) per ``networkInference`` call and takes
``scores.argmax()`` (python/ossid/scripts/online_learning.py:464-468).  ``FrameScorer`` batches
that over the objects of a frame, returns the per-object top-k by (score desc, hypothesis index
asc) -- top-1 is exactly the reference's argmax -- and shards hypotheses across the ranks of a
``torch.distributed`` group: each rank scores a contiguous slice of every object's hypothesis
list, and the only data-path collective is one all-gather of k (score, index) records per object.
"""
from __future__ import annotations

import contextlib
from typing import List, Optional, Sequence

import numpy as np
import torch

from .engine import ZsContext, get_context, poses_to_rt12


def shard_range(n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of ``n`` hypotheses owned by ``rank``; global index = lo + local index."""
    per = -(-n // world) if n > 0 else 0
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def spatial_order(points) -> np.ndarray:
    """Permutation that sorts model points along a 3-D Morton (Z-order) curve.

    The scorer is invariant to the order of a hypothesis' points (shared MLP + max-pool), so the hot
    path uploads the cloud in this order: the 32 points a warp projects together then land on a
    compact image patch and their frame gathers share 128-byte lines instead of touching 32.
    """
    p = np.asarray(points, dtype=np.float64)
    lo, hi = p.min(axis=0), p.max(axis=0)
    q = ((p - lo) / np.where(hi > lo, hi - lo, 1.0) * 1023.0).astype(np.uint64)

    def spread(v):                                   # 10 bits -> every third bit
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v

    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    return np.argsort(code, kind="stable")


def merge_topk(scores: torch.Tensor, idx: torch.Tensor, k: int):
    """Merge candidate lists (..., C) -> (..., k) by (score desc, index asc); empty slots are (-inf, -1).

    Deterministic and identical on every rank, so top-1 does not depend on the GPU count.
    """
    s = torch.where(idx < 0, torch.full_like(scores, float("-inf")), scores)
    s = torch.where(s != s, torch.full_like(s, float("-inf")), s)          # NaN never wins
    big = torch.iinfo(idx.dtype).max
    i_key = torch.where(idx < 0, torch.full_like(idx, big), idx)
    o1 = torch.argsort(i_key, dim=-1, stable=True)
    s1, i1 = torch.gather(s, -1, o1), torch.gather(idx, -1, o1)
    o2 = torch.argsort(-s1, dim=-1, stable=True)
    s2, i2 = torch.gather(s1, -1, o2), torch.gather(i1, -1, o2)
    if s2.shape[-1] < k:
        pad = k - s2.shape[-1]
        s2 = torch.cat([s2, s2.new_full(s2.shape[:-1] + (pad,), float("-inf"))], -1)
        i2 = torch.cat([i2, i2.new_full(i2.shape[:-1] + (pad,), -1)], -1)
    return s2[..., :k].contiguous(), i2[..., :k].contiguous()


def allgather_topk(s: torch.Tensor, i: torch.Tensor, k: int, group=None, ctx: Optional[ZsContext] = None):
    """(n_obj,k) local candidates on each rank -> merged (n_obj,k), same on every rank.

    One all-gather of (score bits, index) records, then the merge.  With a ``ctx`` (CUDA tensors) the merge is one
    launch of the segmented top-k kernel over the (n_obj, world*k) candidates: they are laid out rank-major, every
    rank's list is already ordered by (score desc, index asc) and ranks own ascending index ranges, so "position
    ascending" is "global index ascending" among equal scores - the same rule as ``merge_topk``.
    """
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return merge_topk(s, i, k)
    # one record per candidate: score bits and index side by side (indices ship as int32, not as floats)
    rec = torch.stack([s.to(torch.float32).view(torch.int32), i.to(torch.int32)], dim=-1).contiguous()
    if rec.is_cuda:
        out = torch.empty((world,) + tuple(rec.shape), dtype=rec.dtype, device=rec.device)
        dist.all_gather_into_tensor(out, rec, group=group)
    else:                                            # gloo (CPU tests) has no all_gather_into_tensor
        parts = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(parts, rec, group=group)
        out = torch.stack(parts)
    cand = out.permute(1, 0, 2, 3).contiguous()                      # (n_obj, world, k, 2)
    n_obj = s.shape[0]
    gs = cand[..., 0].reshape(n_obj, world * k).view(torch.float32)
    gi = cand[..., 1].reshape(n_obj, world * k)
    return merge_gathered(gs, gi.to(i.dtype), k, ctx if s.is_cuda else None)


def merge_gathered(gs: torch.Tensor, gi: torch.Tensor, k: int, ctx: Optional[ZsContext] = None):
    """Merge rank-major candidate lists (n_obj, world*k) -> (n_obj,k): one segmented top-k launch (``ctx``) or the
    torch reference ``merge_topk``; both order by (score desc, global index asc) with (-inf, -1) for empty slots."""
    if ctx is None:
        return merge_topk(gs, gi, k)
    n_obj, width = gs.shape
    gs = torch.where((gi < 0) | (gs != gs), torch.full_like(gs, float("-inf")), gs)       # empty slots and NaN never win
    S, I = ctx.topk_segments(gs.reshape(-1).contiguous(), _merge_segments(n_obj, width, gs.device), k,
                             index_map=gi.to(torch.int32).reshape(-1).contiguous())
    return S, torch.where(S == float("-inf"), torch.full_like(I, -1), I).to(gi.dtype)


_seg_cache = {}


def _merge_segments(n_obj: int, width: int, device):
    key = (n_obj, width, str(device))
    if key not in _seg_cache:
        _seg_cache[key] = torch.tensor([[o * width, width, 0, 0] for o in range(n_obj)], dtype=torch.int32, device=device)
    return _seg_cache[key]


class FrameScorer:
    """Scores all objects of one RGB-D frame and returns per-object top-k hypotheses.

    ``weights``: list of folded weight dicts (``weights.fold_state_dict``); ``weight_of(obj_index)``
    picks one per object (the reference keys two YCB-V scorers on object-id parity,
    online_learning.py:461-463).
    """

    def __init__(self, weights: Sequence[dict], device: Optional[int] = None, precision: str = "bf16",
                 inconst_ratio_th: float = 100.0, k: int = 8, chunk: int = 32768, group=None,
                 ctx: Optional[ZsContext] = None, reorder_points: bool = True):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        if device is None:
            device = torch.cuda.current_device()
        self.ctx = ctx or get_context(device)
        self.dtype = torch.float32 if precision == "fp32" else torch.bfloat16
        self.th, self.k, self.chunk, self.group = float(inconst_ratio_th), int(k), int(chunk), group
        for slot, w in enumerate(weights):
            self.ctx.set_weights(slot, w)
        self.n_weights = len(weights)
        self.reorder_points = reorder_points
        self._feat = None
        self._pooled = None
        self._scores = None
        self._resident = None
        self._segments = None
        self._copy_stream = None
        self._scored = (0, [], 0)     # see last_scored
        self.forced_rank_world = None
        self.stage_events = None     # set to [] to record (stage, units, start_event, end_event) per launch group

    @property
    def last_scored(self) -> int:
        """Hypotheses that passed the pre-filter in the last run_resident() on this rank (reads device counts back)."""
        total, devs, cap = self._scored
        return total - cap + (int(torch.cat(devs).sum().item()) if devs else 0)

    def _mark(self, stage, units, prev=None):
        """Stage timing hook: closes the previous stage's CUDA-event pair and opens the next one."""
        if self.stage_events is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.ctx.device))
        if prev is not None:
            self.stage_events.append((prev[0], prev[1], prev[2], ev))
        return (stage, units, ev) if stage is not None else None

    def stage_times_ms(self):
        """Sum of event-timed milliseconds, launch-group count and units per stage (after a synchronize)."""
        out = {}
        for stage, units, a, b in self.stage_events or []:
            t = out.setdefault(stage, dict(ms=0.0, calls=0, units=0))
            t["ms"] += a.elapsed_time(b)
            t["calls"] += 1
            t["units"] += units
        return out

    # -- distributed plumbing ---------------------------------------------------------
    def _rank_world(self):
        if self.forced_rank_world is not None:        # tests: emulate rank r of P on one device, no process group
            return self.forced_rank_world
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    # -- upload (host -> HBM) ---------------------------------------------------------
    def prefetch_poses(self, objects: List[dict]):
        """Start the host-to-device copy of this rank's pose slices on a side stream (the poses are >80 % of a frame's
        upload bytes) so that it overlaps the kernels of the frame before; ``upload(..., prefetched=...)`` consumes it."""
        ctx = self.ctx
        rank, world = self._rank_world()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=ctx.device)
        compute = torch.cuda.current_stream(ctx.device)
        out = []
        with torch.cuda.stream(self._copy_stream):
            for ob in objects:
                lo, hi = shard_range(len(ob["pose_hypos"]), rank, world)
                p12 = poses_to_rt12(torch.as_tensor(ob["pose_hypos"])[lo:hi], ctx.device)
                p12.record_stream(compute)              # allocated on the copy stream, read by kernels on the compute stream
                out.append(p12)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return out, ev

    def upload(self, img_u8, depth, cam_K, objects: List[dict], weight_of=lambda o: 0, prefetched=None):
        """Copy one frame's inputs to the device and keep them resident; this rank's pose slices only."""
        from .zephyr_utils import K2meta
        ctx = self.ctx
        rank, world = self._rank_world()
        meta = {k: float(v) for k, v in K2meta(np.asarray(cam_K)).items()}
        ctx.set_frame_u8(img_u8, depth, meta, blur=True)
        res = []
        for o, ob in enumerate(objects):
            slot = o % 64
            host = ob.get("_zs_host")
            if host is None:
                # clouds are static assets (the reference preloads them once, online_learning.py:303-311): keep a
                # float32, Morton-ordered, pinned host copy so that every frame's upload is one async copy each
                pts, cols, nrms = (torch.as_tensor(ob[key]).to(torch.float32) for key in
                                   ("model_points", "model_colors", "model_normals"))
                if self.reorder_points:
                    perm = torch.from_numpy(spatial_order(pts.numpy()))
                    pts, cols, nrms = pts[perm], cols[perm], nrms[perm]
                host = tuple(t.contiguous().pin_memory() for t in (pts, cols, nrms))
                ob["_zs_host"] = host
            pts, cols, nrms = host
            ctx.set_object(slot, pts, cols, nrms, token=host)      # skipped when this slot already holds this cloud
            M = len(ob["pose_hypos"])
            lo, hi = shard_range(M, rank, world)
            poses12 = prefetched[0][o] if prefetched is not None else \
                poses_to_rt12(torch.as_tensor(ob["pose_hypos"])[lo:hi], ctx.device)
            res.append(dict(slot=slot, poses12=poses12, lo=lo, M=M, wslot=weight_of(o) % max(self.n_weights, 1)))
        if prefetched is not None:
            torch.cuda.current_stream(ctx.device).wait_event(prefetched[1])
        self._resident = res
        self._segments = None
        return res

    # -- compute (everything resident) --------------------------------------------------
    def run_resident(self):
        """Featurise + score + top-k for the uploaded frame.  Returns (scores (n_obj,k), index (n_obj,k))
        tensors on the device; indices are global hypothesis indices per object, -1 = empty slot."""
        ctx, k = self.ctx, self.k
        res = self._resident
        tensor_cores = self.dtype == torch.bfloat16
        # 1. free-space pre-filter per object.  Tensor-core path: the kept count stays on the device (the feature and
        #    MLP kernels read it, zs_set_dynamic_count) and buffers are laid out by capacity, so a filtered frame is as
        #    asynchronous as an unfiltered one.  fp32 parity path: reads back one count per object.
        keeps, n_keeps, n_devs = [], [], []
        for r in res:
            keep, n_dev, M = None, None, r["poses12"].shape[0]
            if self.th < 100 and M > 0:
                viol = ctx.violations(r["slot"], r["poses12"])
                if tensor_cores:
                    keep, n_dev = ctx.filter_async(viol, ctx.obj_npts[r["slot"]], self.th)
                else:
                    keep = ctx.filter(viol, ctx.obj_npts[r["slot"]], self.th)
            keeps.append(keep)
            n_devs.append(n_dev)
            n_keeps.append(M if (keep is None or n_dev is not None) else keep.shape[0])     # capacity when the count is on the device
        # 2. objects laid out by weight slot so that the head runs once per scorer over a contiguous range
        order = sorted(range(len(res)), key=lambda o: res[o]["wslot"])
        offs, total = {}, 0
        for o in order:
            offs[o] = total
            total += n_keeps[o]
        self._scored = (total, [d for d in n_devs if d is not None],
                        sum(n for n, d in zip(n_keeps, n_devs) if d is not None))
        if self._pooled is None or self._pooled.shape[0] < total:
            self._pooled = torch.zeros((max(total, 1), 1024), dtype=torch.float32, device=ctx.device)
            self._scores = torch.zeros((max(total, 1),), dtype=torch.float32, device=ctx.device)
        # 3. features -> shared MLP + max-pool, chunked so that the feature buffer stays bounded
        same_n = len({ctx.obj_npts[r["slot"]] for r in res}) == 1
        if same_n and all(kp is None for kp in keeps) and total > 0:
            # no pre-filter, one cloud size: the MLP kernel runs once per chunk of the scorer's concatenated hypothesis
            # list (several objects per launch) instead of once per object - fewer launch prologues and a last wave of
            # CTA pairs that is 1/443 instead of 1/136 of the launch
            N = ctx.obj_npts[res[0]["slot"]]
            for ws in sorted({res[o]["wslot"] for o in order}):
                members = [o for o in order if res[o]["wslot"] == ws]
                lo, hi = offs[members[0]], offs[members[-1]] + n_keeps[members[-1]]
                for cs in range(lo, hi, self.chunk):
                    ce = min(cs + self.chunk, hi)
                    need = (ce - cs) * N * 8
                    if self._feat is None or self._feat.numel() < need or self._feat.dtype != self.dtype:
                        self._feat = torch.empty((max(need, min(self.chunk, hi - lo) * N * 8),), dtype=self.dtype,
                                                 device=ctx.device)
                    feat = self._feat[:need].view(ce - cs, N, 8)
                    t = self._mark("features", (ce - cs) * N)
                    for o in members:
                        a, b = max(cs, offs[o]), min(ce, offs[o] + n_keeps[o])
                        if a < b:
                            ctx.features(res[o]["slot"], res[o]["poses12"][a - offs[o]: b - offs[o]], n_keep=b - a,
                                         out=feat[a - cs: b - cs])
                    t = self._mark("pool", (ce - cs) * N, t)
                    ctx.pool(ws, feat, out=self._pooled[cs:ce])
                    self._mark(None, 0, t)
            order_for_loop = []
        else:
            order_for_loop = order
        for o in order_for_loop:
            r, keep, n_keep, n_dev = res[o], keeps[o], n_keeps[o], n_devs[o]
            poses12, N = r["poses12"], ctx.obj_npts[r["slot"]]
            for s in range(0, n_keep, self.chunk):
                e = min(s + self.chunk, n_keep)
                need = (e - s) * N * 8
                if self._feat is None or self._feat.numel() < need or self._feat.dtype != self.dtype:
                    self._feat = torch.empty((max(need, min(self.chunk, n_keep) * N * 8),), dtype=self.dtype,
                                             device=ctx.device)
                feat = self._feat[:need].view(e - s, N, 8)
                with (ctx.dynamic_count(n_dev, s) if n_dev is not None else contextlib.nullcontext()):
                    t = self._mark("features", (e - s) * N)
                    if keep is None:
                        ctx.features(r["slot"], poses12[s:e], n_keep=e - s, out=feat)
                    else:
                        ctx.features(r["slot"], poses12, keep_idx=keep[s:e], out=feat)
                    t = self._mark("pool", (e - s) * N, t)
                    ctx.pool(r["wslot"], feat, out=self._pooled[offs[o] + s: offs[o] + e])
                    self._mark(None, 0, t)
        # 4. head, one pass per scorer
        for ws in sorted({res[o]["wslot"] for o in order}):
            members = [o for o in order if res[o]["wslot"] == ws]
            lo = offs[members[0]]
            hi = offs[members[-1]] + n_keeps[members[-1]]
            if hi > lo:
                t = self._mark("head", hi - lo)
                ctx.head(ws, self._pooled[lo:hi], tensor_cores, out=self._scores[lo:hi])
                self._mark(None, 0, t)
        # 5. per-object top-k in one launch (one CTA per object); indices mapped back to global hypothesis indices
        #    inside the kernel.  Without a pre-filter the segment table only depends on the uploaded frame and is built
        #    once per upload; with device-side counts it is assembled on the device (no read-back).
        if all(kp is None for kp in keeps):
            if self._segments is None:
                self._segments = torch.tensor([[offs[o], n_keeps[o], r["lo"], 0] for o, r in enumerate(res)],
                                              dtype=torch.int32).pin_memory().to(ctx.device, non_blocking=True)
            S, I = ctx.topk_segments(self._scores, self._segments, k)
        elif all(d is not None or n_keeps[o] == 0 for o, d in enumerate(n_devs)):
            seg = torch.tensor([[offs[o], 0, r["lo"], 0] for o, r in enumerate(res)], dtype=torch.int32).pin_memory() \
                .to(ctx.device, non_blocking=True)
            zero = torch.zeros((1,), dtype=torch.int32, device=ctx.device)
            seg[:, 1] = torch.cat([d if d is not None else zero for d in n_devs])
            index_map = torch.zeros((max(total, 1),), dtype=torch.int32, device=ctx.device)
            for o, d in enumerate(n_devs):
                if d is not None:
                    index_map[offs[o]: offs[o] + n_keeps[o]] = keeps[o][: n_keeps[o]]
            S, I = ctx.topk_segments(self._scores, seg, k, index_map=index_map)
        else:
            top_s, top_i = [], []
            for o, r in enumerate(res):
                ts, ti = ctx.topk(self._scores[offs[o]: offs[o] + n_keeps[o]], k, r["lo"], index_map=keeps[o])
                top_s.append(ts)
                top_i.append(ti)
            S, I = torch.stack(top_s), torch.stack(top_i)
        rank, world = self._rank_world()
        if world > 1 and self.forced_rank_world is None:
            S, I = allgather_topk(S, I, k, self.group, ctx=ctx)
        return S, I

    # -- post-scoring refinement (online_learning.py:471-479) ----------------------------------
    def refine_winners(self, objects: List[dict], I, n_refine: int = 1, icp_max_dist: float = 0.01, max_iter: int = 30):
        """ICP-refine the ``n_refine`` best hypotheses of every object of the uploaded frame in one launch per object.

        ``objects``: the list given to ``upload`` (host ``pose_hypos``); ``I``: (n_obj,k) global winner indices from
        ``run_resident`` (entries < 0 are skipped).  Target cloud = the resident depth at ``uv_original`` of each winner,
        as the reference passes ``uv_original[pred_idx]`` (online_learning.py:476-479).  Returns
        ``(poses (n_obj,n_refine,4,4) float64, stats (n_obj,n_refine,4) float32)`` numpy; skipped slots keep identity / zeros.
        """
        ctx = self.ctx
        I = np.asarray(I.cpu() if torch.is_tensor(I) else I)
        n_obj = len(objects)
        poses = np.tile(np.eye(4), (n_obj, n_refine, 1, 1))
        stats = np.zeros((n_obj, n_refine, 4), np.float32)
        pending = []
        for o, (ob, r) in enumerate(zip(objects, self._resident)):
            idx = [int(i) for i in I[o, :n_refine] if i >= 0]
            if not idx:
                continue
            p12 = poses_to_rt12(torch.as_tensor(np.asarray(ob["pose_hypos"])[idx]), ctx.device)
            _, uv, _, _ = ctx.features(r["slot"], p12, dtype=self.dtype, want_uv=True)
            out, st = ctx.icp_refine(p12, ob["_zs_host"][0], uv, max_dist=icp_max_dist, max_iter=max_iter)
            pending.append((o, len(idx), out, st))
        for o, m, out, st in pending:                       # read back after every object's launches are queued
            poses[o, :m, :3, :] = out.cpu().numpy().astype(np.float64).reshape(m, 3, 4)
            stats[o, :m] = st.cpu().numpy()
        return poses, stats

    # -- public end-to-end call -------------------------------------------------------------
    def score_frame(self, img_u8, depth, cam_K, objects: List[dict], weight_of=lambda o: 0):
        """Host buffers in, host results out: ``(scores (n_obj,k), indices (n_obj,k))`` numpy arrays."""
        self.upload(img_u8, depth, cam_K, objects, weight_of)
        S, I = self.run_resident()
        return S.cpu().numpy(), I.cpu().numpy()

    def score_frames(self, frames: List[dict], weight_of=lambda o: 0, depth: int = 2):
        """Stream of frames (BASELINE.json config 5: multi-frame scoring between finetune steps).

        ``frames``: dicts with ``img`` (uint8), ``depth``, ``cam_K``, ``objects``.  Uploads, kernels and the
        read-back of each frame's top-k are all asynchronous; the host only blocks when ``depth`` frames are in
        flight, and the pose hypotheses of frame f+1 (most of a frame's bytes) travel on a side stream while the kernels
        of frame f run.  Returns a list of
        ``(scores (n_obj,k), indices (n_obj,k))`` numpy pairs, one per frame.  The free-space pre-filter
        (inconst_ratio_th < 100) keeps its counts on the device on the tensor-core path; only the fp32 parity path
        synchronises once per object to read the kept count.
        """
        stream = torch.cuda.current_stream(self.ctx.device)
        inflight, out = [], []

        def drain():
            ev, s_h, i_h = inflight.pop(0)
            ev.synchronize()
            out.append((s_h.numpy().copy(), i_h.numpy().copy()))

        nxt = self.prefetch_poses(frames[0]["objects"]) if frames else None
        for f, fr in enumerate(frames):
            self.upload(fr["img"], fr["depth"], fr["cam_K"], fr["objects"], weight_of, prefetched=nxt)
            nxt = self.prefetch_poses(frames[f + 1]["objects"]) if f + 1 < len(frames) else None    # overlaps this frame's kernels
            S, I = self.run_resident()
            s_h = torch.empty(S.shape, dtype=S.dtype, pin_memory=True)
            i_h = torch.empty(I.shape, dtype=I.dtype, pin_memory=True)
            s_h.copy_(S, non_blocking=True)
            i_h.copy_(I, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            inflight.append((ev, s_h, i_h))
            if len(inflight) >= depth:
                drain()
        while inflight:
            drain()
        return out
