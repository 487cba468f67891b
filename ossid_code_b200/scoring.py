"""Multi-object, multi-GPU frame scoring on top of the C ABI.

The reference scores one (object, frame) per ``networkInference`` call and takes
``scores.argmax()`` (python/ossid/scripts/online_learning.py:464-468).  ``FrameScorer`` batches
that over the objects of a frame, returns the per-object top-k by (score desc, hypothesis index
asc) -- top-1 is exactly the reference's argmax -- and shards hypotheses across the ranks of a
``torch.distributed`` group: each rank scores a contiguous slice of every object's hypothesis
list, and the only data-path collective is one all-gather of a per-rank candidate record
(k scores, indices and poses per object), merged identically on every rank by ``zs_merge_topk``.

With the bf16 tensor-core scorer the k candidates of every object are then re-scored by the
fp32-accurate scorer (``rerank``) and ordered by that, so the winning hypothesis is the fp32
argmax of the candidates whatever the GPU count.

PyTorch is plumbing here (device buffers, streams, the process group); every computation inside a
step is a kernel of libzs.so.  All per-frame device and pinned-host buffers are owned by the
``FrameScorer`` and reused, so a warmed-up step allocates nothing.
"""
from __future__ import annotations

import contextlib
from typing import List, Optional, Sequence

import numpy as np
import torch

from .engine import ZsContext, get_context, poses_to_rt12

MAX_OBJECTS = 64          # ZS_MAX_OBJECTS (include/zs.h): model-cloud slots per context
MAX_WEIGHT_SLOTS = 4      # ZS_MAX_WEIGHT_SLOTS


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


# ---- candidate records: what one rank contributes to the all-gather ---------------------------------------------
# int32 [rec_ints]:  n_obj*k score bits | n_obj*k global indices (-1 = empty) | n_obj x {kept by the pre-filter,
# fallback violation count} | (optional) n_obj*k*12 pose floats of the candidates (for the fp32 re-rank).
def record_pose_offset(n_obj: int, k: int) -> int:
    """Offset (in ints) of the pose section: rounded up to 4 ints so that every pose row is 16-byte aligned."""
    return (2 * n_obj * k + 2 * n_obj + 3) & ~3


def record_ints(n_obj: int, k: int, with_poses: bool) -> int:
    return record_pose_offset(n_obj, k) + (12 * n_obj * k if with_poses else 0)


def merge_topk(scores: torch.Tensor, idx: torch.Tensor, k: int):
    """Torch reference of the candidate merge: (..., C) -> (..., k) by (score desc, index asc); empty slots
    (index < 0) are (-inf, -1); a NaN never beats a real score and is reported as -inf.

    Deterministic and identical on every rank, so top-1 does not depend on the GPU count.  The device path is
    ``zs_merge_topk`` (one launch); this function is its checker and the CPU (gloo) path.
    """
    s = torch.where(idx < 0, torch.full_like(scores, float("-inf")), scores)
    nan = s != s
    s = torch.where(nan, torch.full_like(s, float("-inf")), s)
    big = torch.iinfo(idx.dtype).max
    # order: real scores, then NaNs, then empty slots; inside each class (score desc, index asc)
    cls = torch.where(idx < 0, torch.full_like(idx, 2), torch.where(nan, torch.ones_like(idx), torch.zeros_like(idx)))
    i_key = torch.where(idx < 0, torch.full_like(idx, big), idx)
    o1 = torch.argsort(i_key, dim=-1, stable=True)
    s1, i1, c1 = torch.gather(s, -1, o1), torch.gather(idx, -1, o1), torch.gather(cls, -1, o1)
    o2 = torch.argsort(-s1, dim=-1, stable=True)
    s2, i2, c2 = torch.gather(s1, -1, o2), torch.gather(i1, -1, o2), torch.gather(c1, -1, o2)
    o3 = torch.argsort(c2, dim=-1, stable=True)
    s3, i3 = torch.gather(s2, -1, o3), torch.gather(i2, -1, o3)
    if s3.shape[-1] < k:
        pad = k - s3.shape[-1]
        s3 = torch.cat([s3, s3.new_full(s3.shape[:-1] + (pad,), float("-inf"))], -1)
        i3 = torch.cat([i3, i3.new_full(i3.shape[:-1] + (pad,), -1)], -1)
    return s3[..., :k].contiguous(), i3[..., :k].contiguous()


def merge_records_reference(gathered: torch.Tensor, n_obj: int, k: int):
    """Torch/CPU restatement of ``zs_merge_topk`` (include/zs.h) on (world, rec_ints) int32 records, including the
    never-empty rule applied to the object's whole hypothesis list."""
    g = gathered.cpu()
    world = g.shape[0]
    nk = n_obj * k
    S = g[:, :nk].contiguous().view(torch.float32).reshape(world, n_obj, k).clone()
    I = g[:, nk:2 * nk].reshape(world, n_obj, k).clone()
    info = g[:, 2 * nk:2 * nk + 2 * n_obj].reshape(world, n_obj, 2)
    for o in range(n_obj):
        kept = info[:, o, 0] > 0
        if bool(kept.any()):
            I[~kept, o, :] = -1                                     # fallback candidates of ranks that kept nothing
        else:
            have = [w for w in range(world) if int(I[w, o, 0]) >= 0]
            best = min(have, key=lambda w: (int(info[w, o, 1]), w)) if have else -1
            for w in range(world):
                if w != best:
                    I[w, o, :] = -1
    return merge_topk(S.permute(1, 0, 2).reshape(n_obj, world * k), I.permute(1, 0, 2).reshape(n_obj, world * k), k)


def allgather_records(rec: torch.Tensor, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One all-gather of this rank's candidate record -> (world, rec_ints), same on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world, rec.numel()), dtype=rec.dtype, device=rec.device)
    if rec.is_cuda:
        dist.all_gather_into_tensor(out, rec, group=group)
    else:                                            # gloo (CPU tests) has no all_gather_into_tensor
        parts = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(parts, rec, group=group)
        out.copy_(torch.stack(parts))
    return out


def allgather_topk(s: torch.Tensor, i: torch.Tensor, k: int, group=None, ctx: Optional[ZsContext] = None, info=None):
    """(n_obj,k) local candidates on each rank -> merged (n_obj,k), same on every rank: one all-gather of the packed
    record, then the merge (``zs_merge_topk`` with a ``ctx``, else the torch reference)."""
    n_obj = s.shape[0]
    if info is None:
        info = torch.tensor([[1, 0]] * n_obj, dtype=torch.int32, device=s.device)
    rec = torch.cat([s.to(torch.float32).contiguous().view(torch.int32).reshape(-1), i.to(torch.int32).reshape(-1),
                     info.to(torch.int32).reshape(-1)])
    g = allgather_records(rec, group)
    if ctx is not None and s.is_cuda:
        S, I = ctx.merge_topk(g, n_obj, k)
    else:
        S, I = merge_records_reference(g, n_obj, k)
    return S, I.to(i.dtype)


class _Plan:
    """Everything about a frame that depends only on (hypothesis counts, cloud slots, scorer of each object, rank,
    world): the row layout of the concatenated hypothesis list and the small device tables.  Cached per shape."""
    pass


class FrameScorer:
    """Scores all objects of one RGB-D frame and returns per-object top-k hypotheses.

    ``weights``: list of folded weight dicts (``weights.fold_state_dict``); ``weight_of(obj_index)``
    picks one per object (the reference keys two YCB-V scorers on object-id parity,
    online_learning.py:461-463).  ``precision``: "bf16" = bf16 features + bf16 tcgen05 scorer (1e-2),
    "fp32" = fp32-accurate scorer (1e-4).  ``rerank`` (bf16 only, default on): the k candidates of every object are
    re-scored by the fp32-accurate scorer and ordered by it, which makes the reported top-1 the fp32 argmax of the
    candidates (python/ossid/scripts/online_learning.py:466-467) at any GPU count.
    """

    def __init__(self, weights: Sequence[dict], device: Optional[int] = None, precision: str = "bf16",
                 inconst_ratio_th: float = 100.0, k: int = 8, chunk: int = 32768, group=None,
                 ctx: Optional[ZsContext] = None, reorder_points: bool = True, rerank: Optional[bool] = None,
                 fused: bool = True, mask_th: float = 0.5, graph: bool = True):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        if len(weights) > MAX_WEIGHT_SLOTS:
            raise ValueError(f"at most {MAX_WEIGHT_SLOTS} scorers per context, got {len(weights)}")
        if device is None:
            device = torch.cuda.current_device()
        self.ctx = ctx or get_context(device)
        self.precision = precision
        self.split = precision == "fp32"      # fp32-accurate: split-bf16 features + the 3-term tcgen05 scorer + the fp32 head
        self.th, self.k, self.chunk, self.group = float(inconst_ratio_th), int(k), int(chunk), group
        self.mask_th = float(mask_th)     # filterHypoByMask's th for objects that come with a detection mask / boxes
        self.rerank = (precision == "bf16") if rerank is None else bool(rerank and precision == "bf16")
        # bf16 path without pre-filter: features are computed inside the MLP kernel (zs_pool_fused) instead of being
        # written to HBM by zs_features and read back; fused=False keeps the two-kernel sequence (bit-identical results)
        self.fused = bool(fused) and precision == "bf16"
        # single-GPU steps are replayed from a CUDA graph after their first two runs (one eager warm-up, one capture): a
        # frame is 15-40 launches, and for a frame of a few thousand hypotheses (the reference's own call size) issuing
        # them from Python costs more than running them.  Same kernels, same buffers, same results.
        self.use_graph = bool(graph)
        self._graphs, self._graph_warm_gen = {}, {}
        self._buf_gen = 0                 # bumped whenever a FrameScorer-owned device buffer is (re)allocated
        self._weights = list(weights)
        self._wtoken = [object() for _ in weights]       # ownership tokens of this scorer's weight slots
        self.n_weights = len(weights)
        self._sync_weights()
        self.reorder_points = reorder_points
        self._feat = None
        self._pooled = None
        self._scores = None
        self._resident = None
        self._plan = None
        self._plans = {}
        self._copy_stream = None
        self._img_dev = self._depth_dev = None
        self._mask_dev, self._masks = None, []
        self._raw = [None, None]          # staging of the raw (m,4,4) pose blocks, double-buffered
        self._raw_free = [None, None]     # event: the pack kernel that last read _raw[b] has run
        self._p12 = [None, None]
        self._buf = 0
        self._pin = {}                    # pinned result ring
        self._gathered = None
        self._rr = None                   # re-rank scratch
        self._result_flat = None          # int32 view over the last result: n_obj*k score bits, then n_obj*k indices
        self._scored = (0, [], 0)     # see last_scored
        self.forced_rank_world = None
        self.stage_events = None     # set to [] to record (stage, units, start_event, end_event) per launch group

    # -- weights ------------------------------------------------------------------------------------
    def _sync_weights(self):
        """(Re-)upload this scorer's weights into slots 0..n-1 unless the context still credits them to it: another
        user of the shared per-device context (a ``PointNet2SSG``, a second ``FrameScorer``) may have taken a slot."""
        for slot, (w, tok) in enumerate(zip(self._weights, self._wtoken)):
            if self.ctx.weight_owner.get(slot) is not tok:
                self.ctx.set_weights(slot, w, owner=tok)

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

    # -- per-shape plan -----------------------------------------------------------------
    def _make_plan(self, counts, wslots, npts, masked=False):
        rank, world = self._rank_world()
        key = (tuple(counts), tuple(wslots), tuple(npts), rank, world, self.k, bool(masked))
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        dev, k, n_obj = self.ctx.device, self.k, len(counts)
        plan = _Plan()
        plan.n_obj, plan.rank, plan.world = n_obj, rank, world
        plan.lo = [shard_range(M, rank, world)[0] for M in counts]
        plan.m_loc = [shard_range(M, rank, world)[1] - shard_range(M, rank, world)[0] for M in counts]
        plan.M = list(counts)
        plan.wslot = list(wslots)
        # objects laid out by weight slot so that MLP and head run once per scorer over a contiguous row range
        plan.order = sorted(range(n_obj), key=lambda o: wslots[o])
        plan.off, total = {}, 0
        for o in plan.order:
            plan.off[o] = total
            total += plan.m_loc[o]
        plan.total = total
        to_dev = lambda rows: torch.tensor(rows, dtype=torch.int32).reshape(-1, 4).to(dev)
        # top-k segments {first score row, count, index base, index-map offset}; candidate poses {first row, lo, 0, 0}
        plan.seg = to_dev([[plan.off[o], plan.m_loc[o], plan.lo[o], 0] for o in range(n_obj)])
        plan.pose_seg = to_dev([[plan.off[o], plan.lo[o], 0, 0] for o in range(n_obj)])
        # re-rank rows: object o's k candidates sit at rows [rr_row[o]*k, +k), objects again grouped by scorer
        plan.rr_row = {o: r for r, o in enumerate(plan.order)}
        plan.rr_seg = to_dev([[plan.rr_row[o] * k, k, 0, (o - plan.rr_row[o]) * k] for o in range(n_obj)])
        plan.rec_ints = record_ints(n_obj, k, self.rerank)
        plan.pose_at = record_pose_offset(n_obj, k)
        plan.rec = torch.zeros((plan.rec_ints,), dtype=torch.int32, device=dev)
        plan.filtered = self.th < 100 or bool(masked)
        if not plan.filtered:      # no pre-filter: every rank "kept" its hypotheses; with one, zs_filter writes this section
            plan.rec[2 * n_obj * k: 2 * n_obj * k + 2 * n_obj] = torch.tensor([[1, 0]] * n_obj, dtype=torch.int32).reshape(-1).to(dev)
        else:                      # device-side pre-filter state: violation counts, kept lists (= the top-k index map) and
            plan.viol = torch.zeros((max(total, 1),), dtype=torch.int32, device=dev)     # counts (= the segment table)
            plan.index_map = torch.zeros((max(total, 1),), dtype=torch.int32, device=dev)
            plan.seg_dyn = plan.seg.clone()
        plan.out = torch.zeros((2 * n_obj * k,), dtype=torch.int32, device=dev)
        plan.P = torch.zeros((n_obj, k, 12), dtype=torch.float32, device=dev) if self.rerank else None
        self._plans[key] = plan
        return plan

    # -- upload (host -> HBM) ---------------------------------------------------------
    def _stage_poses(self, objects: List[dict], plan, buf: int, stream):
        """Copy this rank's slice of every object's (M,4,4) block into the raw staging buffer ``buf`` on ``stream``
        (row order = the plan's scorer-grouped layout).  The cast to float32 rows happens on the device."""
        ctx = self.ctx
        hosts = [torch.as_tensor(ob["pose_hypos"]) for ob in objects]
        for t in hosts:
            if t.ndim != 3 or t.shape[1:] != (4, 4):
                raise ValueError(f"pose hypotheses must be (M,4,4), got {tuple(t.shape)}")
        dtype = torch.float32 if all(t.dtype == torch.float32 for t in hosts) else torch.float64
        raw = self._raw[buf]
        if raw is None or raw.dtype != dtype or raw.shape[0] < plan.total:
            cap = max(plan.total, 1) if raw is None or raw.dtype != dtype else max(plan.total, int(raw.shape[0] * 1.25))
            torch.cuda.current_stream(ctx.device).synchronize()      # growth only: nothing may still read the old block
            raw = self._raw[buf] = torch.empty((cap, 4, 4), dtype=dtype, device=ctx.device)
            self._raw_free[buf] = None
        if self._raw_free[buf] is not None:
            stream.wait_event(self._raw_free[buf])
        with torch.cuda.stream(stream):
            for o, t in enumerate(hosts):
                if plan.m_loc[o]:
                    src = t[plan.lo[o]: plan.lo[o] + plan.m_loc[o]]
                    if src.dtype != dtype:
                        src = src.to(dtype)
                    raw[plan.off[o]: plan.off[o] + plan.m_loc[o]].copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        return ev

    def prefetch_poses(self, objects: List[dict], weight_of=lambda o: 0):
        """Start the host-to-device copy of this rank's pose slices on a side stream (the poses are >80 % of a frame's
        upload bytes) so that it overlaps the kernels of the frame before; ``upload(..., prefetched=...)`` consumes it."""
        plan = self._plan_for(objects, weight_of)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.ctx.device)
        buf = self._buf ^ 1
        ev = self._stage_poses(objects, plan, buf, self._copy_stream)
        return plan, buf, ev

    def _host_cloud(self, ob):
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
        return host

    def _plan_for(self, objects, weight_of):
        if len(objects) > MAX_OBJECTS:
            raise ValueError(f"a frame may carry at most {MAX_OBJECTS} objects per FrameScorer call (ZS_MAX_OBJECTS), got "
                             f"{len(objects)}: score it in groups")
        counts = [len(ob["pose_hypos"]) for ob in objects]
        wslots = [weight_of(o) % max(self.n_weights, 1) for o in range(len(objects))]
        npts = [len(ob["model_points"]) for ob in objects]
        masked = any(("mask" in ob) or ("boxes" in ob) for ob in objects)
        return self._make_plan(counts, wslots, npts, masked)

    def upload(self, img_u8, depth, cam_K, objects: List[dict], weight_of=lambda o: 0, prefetched=None):
        """Copy one frame's inputs to the device and keep them resident; this rank's pose slices only."""
        from .zephyr_utils import K2meta
        ctx = self.ctx
        stream = torch.cuda.current_stream(ctx.device)
        meta = {k: float(v) for k, v in K2meta(np.asarray(cam_K)).items()}
        # frame: persistent device buffers, two async copies, blur + /255 + HSV on the device
        img = torch.as_tensor(img_u8)
        dep = torch.as_tensor(depth)
        if img.dtype != torch.uint8:
            raise ValueError("img_u8 must be uint8")
        if dep.dtype != torch.float32:
            dep = dep.to(torch.float32)
        if self._img_dev is None or self._img_dev.shape != img.shape:
            self._img_dev = torch.empty(img.shape, dtype=torch.uint8, device=ctx.device)
            self._depth_dev = torch.empty(dep.shape, dtype=torch.float32, device=ctx.device)
        self._img_dev.copy_(img, non_blocking=True)
        self._depth_dev.copy_(dep, non_blocking=True)
        ctx.set_frame_u8(self._img_dev, self._depth_dev, meta, blur=True)
        if prefetched is not None:
            plan, buf, ev = prefetched
        else:
            plan = self._plan_for(objects, weight_of)
            buf = self._buf ^ 1
            ev = self._stage_poses(objects, plan, buf, stream)
        self._masks = [None] * len(objects)
        for o, ob in enumerate(objects):
            pts, cols, nrms = host = self._host_cloud(ob)
            ctx.set_object(o, pts, cols, nrms, token=host)      # skipped when this slot already holds this cloud
            # detection prior of the object (the reference's DTOID stage, online_learning.py:383-405): a mask, or boxes
            # that are rasterised on the device against this frame's depth; applied as filterHypoByMask in the pre-filter
            if "mask" in ob or "boxes" in ob:
                if self._mask_dev is None or self._mask_dev.shape[0] < len(objects) or self._mask_dev.shape[1:] != tuple(ctx.frame_hw):
                    self._mask_dev = torch.zeros((len(objects),) + tuple(ctx.frame_hw), dtype=torch.uint8, device=ctx.device)
                    self._buf_gen += 1
                m = self._mask_dev[o]
                if "mask" in ob:
                    m.copy_((torch.as_tensor(ob["mask"]) != 0).to(torch.uint8), non_blocking=True)
                else:
                    ctx.boxes_to_mask(ob["boxes"], ob["box_scores"], ob.get("box_expand", 1.2), out=m)
                self._masks[o] = m
        # one cast kernel over the whole concatenated block: (total,4,4) f32/f64 -> (total,12) f32
        stream.wait_event(ev)
        p12 = self._p12[buf]
        if p12 is None or p12.shape[0] < plan.total:
            p12 = self._p12[buf] = torch.empty((max(int(plan.total * 1.25), 1), 12), dtype=torch.float32, device=ctx.device)
            self._buf_gen += 1
        ctx.pack_poses(self._raw[buf][: plan.total], out=p12)
        done = torch.cuda.Event()
        done.record(stream)
        self._raw_free[buf] = done
        self._buf = buf
        self._plan = plan
        self._resident = [dict(slot=o, poses12=p12[plan.off[o]: plan.off[o] + plan.m_loc[o]], lo=plan.lo[o], M=plan.M[o],
                               wslot=plan.wslot[o]) for o in range(plan.n_obj)]
        ctx.reserve(min(max(plan.total, plan.n_obj * self.k), 1 << 30))
        return self._resident

    # -- compute (everything resident) --------------------------------------------------
    def run_resident(self, local_record: bool = False):
        """Featurise + score + top-k for the uploaded frame.  Returns (scores (n_obj,k), index (n_obj,k))
        tensors on the device; indices are global hypothesis indices per object, -1 = empty slot.
        ``local_record=True`` (tests that emulate ranks on one device) returns this rank's candidate record instead."""
        rank, world = self._rank_world()
        if not self.use_graph or local_record or self.stage_events is not None or world > 1 or self.forced_rank_world:
            return self._run_resident(local_record)
        # the launch sequence only depends on the plan (shapes) and on which pose buffer the upload used
        key = (id(self._plan), self._buf)
        entry = self._graphs.get(key)
        if entry is None:
            self._sync_weights()
            gens = (self._buf_gen, self.ctx.generation)
            if self._graph_warm_gen.get(key) != gens:     # first run: eager (sizes every buffer, loads every kernel)
                out = self._run_resident(False)
                self._graph_warm_gen[key] = (self._buf_gen, self.ctx.generation)
                return out
            try:
                l0 = self.ctx.launches
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    S, I = self._run_resident(False)
                entry = self._graphs[key] = (g, S, I, self._result_flat, self._scored, self.ctx.launches - l0,
                                             (self._buf_gen, self.ctx.generation))
            except Exception as exc:                      # capture refused: keep launching eagerly
                import warnings
                warnings.warn(f"CUDA-graph capture of the scoring step failed ({exc}); launching kernel by kernel")
                self.use_graph = False
                torch.cuda.synchronize(self.ctx.device)
                return self._run_resident(False)
        g, S, I, flat, scored, n_launch, gens = entry
        if gens != (self._buf_gen, self.ctx.generation):  # a buffer the recorded launches point at has moved: re-capture
            del self._graphs[key]
            self._graph_warm_gen.pop(key, None)
            return self.run_resident()
        self._sync_weights()                              # (re-)uploads go into the same buffers the graph reads
        g.replay()
        self.ctx.graph_launches += n_launch
        self._result_flat, self._scored = flat, scored
        return S, I

    def _run_resident(self, local_record: bool = False):
        ctx, k = self.ctx, self.k
        res, plan = self._resident, self._plan
        n_obj, nk = plan.n_obj, plan.n_obj * k
        head_tc = "fp32_tc" if self.split else "tf32"   # tf32 head on the bf16 path, 3-term tf32 (fp32-accurate) on the 1e-4 path
        self._sync_weights()
        rec = plan.rec
        S_loc, I_loc = rec[:nk].view(torch.float32).view(n_obj, k), rec[nk:2 * nk].view(n_obj, k)
        info = rec[2 * nk: 2 * nk + 2 * n_obj]
        # 1. free-space pre-filter per object: the kept count stays on the device (the feature and MLP kernels read it,
        #    zs_set_dynamic_count) and buffers are laid out by capacity, so a filtered frame is as asynchronous as an
        #    unfiltered one.
        keeps, n_keeps, n_devs = [], [], []
        filtered = plan.filtered
        pre = []
        for o, r in enumerate(res):
            keep, n_dev, M = None, None, r["poses12"].shape[0]
            if filtered and M > 0:
                # the kept list lands in the top-k index map and the kept count in the segment table: no glue kernels
                a = plan.off[o]
                keep, n_dev = plan.index_map[a: a + M], plan.seg_dyn.view(-1)[4 * o + 1: 4 * o + 2]
                pre.append((r["slot"], r["poses12"], self._masks[o], plan.viol[a: a + M], keep, n_dev, info[2 * o: 2 * o + 2]))
            keeps.append(keep)
            n_devs.append(n_dev)
            n_keeps.append(M)                    # capacity: the live count stays on the device
        if pre:
            t = self._mark("prefilter", sum(sg[1].shape[0] for sg in pre))
            ctx.prefilter(pre, self.th, self.mask_th)       # every object of the frame: one projection pass, one compaction launch
            self._mark(None, 0, t)
        # 2. row layout (by capacity: the plan's)
        order = plan.order
        offs, total = {}, 0
        for o in order:
            offs[o] = total
            total += n_keeps[o]
        self._scored = (total, [d for d in n_devs if d is not None],
                        sum(n for n, d in zip(n_keeps, n_devs) if d is not None))
        if self._pooled is None or self._pooled.shape[0] < total:
            cap = max(int(total * 1.25), 1)
            self._pooled = torch.zeros((cap, 1024), dtype=torch.float32, device=ctx.device)
            self._scores = torch.zeros((cap,), dtype=torch.float32, device=ctx.device)
            self._buf_gen += 1
        # 3. features -> shared MLP + max-pool
        same_n = len({ctx.obj_npts[r["slot"]] for r in res}) == 1
        N0 = ctx.obj_npts[res[0]["slot"]] if res else 0
        if self.fused and same_n and N0 >= 128 and total > 0:
            # bf16 path, one cloud size: projection + gather + features + MLP + max-pool of a scorer's whole hypothesis
            # list in ONE kernel (zs_pool_fused); with a pre-filter every object's kept list and device-side count go in
            # as they are (rows stay laid out by capacity), so a filtered frame takes the same single launch per scorer
            for ws in sorted({res[o]["wslot"] for o in order}):
                members = [o for o in order if res[o]["wslot"] == ws]
                lo, hi = offs[members[0]], offs[members[-1]] + n_keeps[members[-1]]
                if hi > lo:
                    t = self._mark("pool", (hi - lo) * N0)
                    ctx.pool_fused(ws, [(res[o]["slot"], res[o]["poses12"], keeps[o], n_devs[o]) for o in members
                                        if n_keeps[o] > 0], out=self._pooled[lo:hi])
                    self._mark(None, 0, t)
            order_for_loop = []
        elif same_n and all(kp is None for kp in keeps) and total > 0:
            # no pre-filter, one cloud size, two-kernel path (fused=False or the fp32-accurate scorer): the MLP kernel runs
            # once per chunk of the scorer's concatenated hypothesis list (several objects per launch), chunked so that
            # the feature buffer stays bounded
            N = N0
            for ws in sorted({res[o]["wslot"] for o in order}):
                members = [o for o in order if res[o]["wslot"] == ws]
                lo, hi = offs[members[0]], offs[members[-1]] + n_keeps[members[-1]]
                for cs in range(lo, hi, self.chunk):
                    ce = min(cs + self.chunk, hi)
                    feat = self._feat_buf(ce - cs, N, min(self.chunk, hi - lo))
                    t = self._mark("features", (ce - cs) * N)
                    segs = []
                    for o in members:
                        a, b = max(cs, offs[o]), min(ce, offs[o] + n_keeps[o])
                        if a < b:
                            segs.append((res[o]["slot"], res[o]["poses12"][a - offs[o]: b - offs[o]], feat[a - cs: b - cs]))
                    ctx.features_multi(segs)              # the chunk's objects in one launch
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
                feat = self._feat_buf(e - s, N, min(self.chunk, n_keep))
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
                ctx.head(ws, self._pooled[lo:hi], head_tc, out=self._scores[lo:hi])
                self._mark(None, 0, t)
        # 5. per-object top-k in one launch (one CTA per object), written straight into this rank's candidate record;
        #    indices mapped back to global hypothesis indices inside the kernel.  Without a pre-filter the segment table
        #    is the plan's; with device-side counts it is assembled on the device (no read-back).
        t = self._mark("topk", n_obj)
        if all(kp is None for kp in keeps):
            ctx.topk_segments(self._scores, plan.seg, k, out=(S_loc, I_loc))
        else:
            ctx.topk_segments(self._scores, plan.seg_dyn, k, index_map=plan.index_map, out=(S_loc, I_loc))
        if self.rerank:          # the candidates' poses travel with the record: any rank can re-score any candidate
            ctx.gather_poses(self._p12[self._buf], I_loc, plan.pose_seg, out=rec[plan.pose_at:].view(torch.float32))
        self._mark(None, 0, t)
        if local_record:
            return rec.clone()
        # 6. multi-GPU: one all-gather of the records, one merge launch; identical on every rank
        rank, world = self._rank_world()
        if world > 1 and self.forced_rank_world is None:
            t = self._mark("allgather+merge", n_obj)
            if self._gathered is None or self._gathered.shape != (world, plan.rec_ints):
                self._gathered = torch.empty((world, plan.rec_ints), dtype=torch.int32, device=ctx.device)
            allgather_records(rec, self.group, out=self._gathered)
            S, I, P = self.merge_records(self._gathered)
            self._mark(None, 0, t)
        else:
            S, I = S_loc, I_loc
            P = rec[plan.pose_at:].view(torch.float32).view(n_obj, k, 12) if self.rerank else None
            self._result_flat = rec[: 2 * nk]
        if self.rerank:
            S, I = self._rerank(S, I, P)
        return S, I

    def merge_records(self, gathered: torch.Tensor):
        """(world, rec_ints) all-gathered records -> merged (S, I, candidate poses or None) on the device."""
        plan, k = self._plan, self.k
        n_obj, nk = plan.n_obj, plan.n_obj * k
        out = plan.out
        S, I = out[:nk].view(torch.float32).view(n_obj, k), out[nk:].view(n_obj, k)
        self.ctx.merge_topk(gathered, n_obj, k, out=(S, I), poses_out=plan.P)
        self._result_flat = out
        return S, I, plan.P

    def _feat_buf(self, rows: int, N: int, rows_cap: int):
        """Feature rows of a chunk: (rows,N,8) bf16, or split-bf16 planes (rows,2,N,8) on the fp32-accurate path."""
        per = N * 8 * (2 if self.split else 1)
        need = rows * per
        if self._feat is None or self._feat.numel() < need:
            self._feat = torch.empty((max(need, rows_cap * per),), dtype=torch.bfloat16, device=self.ctx.device)
            self._buf_gen += 1
        return self._feat[:need].view(rows, 2, N, 8) if self.split else self._feat[:need].view(rows, N, 8)

    def _rerank(self, S, I, P):
        """Re-score the k candidates of every object with the fp32-accurate scorer and order them by that score
        ((score desc, index asc) again).  ``P``: (n_obj,k,12) candidate poses; empty slots (I < 0) stay empty."""
        ctx, plan, k = self.ctx, self._plan, self.k
        n_obj = plan.n_obj
        t = self._mark("rerank", n_obj * k)
        rr = self._rr
        Nmax = max(ctx.obj_npts[o] for o in range(n_obj))
        if rr is None or rr["rows"] < n_obj * k or rr["N"] < Nmax:
            rows = n_obj * k
            self._buf_gen += 1
            rr = self._rr = dict(rows=rows, N=Nmax,
                                 feat=torch.zeros((rows * Nmax * 16,), dtype=torch.bfloat16, device=ctx.device),
                                 pooled=torch.zeros((rows, 1024), dtype=torch.float32, device=ctx.device),
                                 scores=torch.zeros((rows,), dtype=torch.float32, device=ctx.device),
                                 out=torch.zeros((2 * rows,), dtype=torch.int32, device=ctx.device))
        same_n = len({ctx.obj_npts[o] for o in range(n_obj)}) == 1
        groups = {}
        for o in plan.order:
            groups.setdefault(plan.wslot[o], []).append(o)
        for ws, members in groups.items():
            r0 = plan.rr_row[members[0]] * k
            if same_n:
                N = Nmax
                feat = rr["feat"][: n_obj * k * N * 16].view(n_obj * k, 2, N, 8)
                if ws == plan.wslot[plan.order[0]]:   # first scorer group: every object's candidates in one launch
                    ctx.features_multi([(o, P[o], feat[plan.rr_row[o] * k: plan.rr_row[o] * k + k]) for o in plan.order])
                ctx.pool_f32a(ws, feat[r0: r0 + len(members) * k], out=rr["pooled"][r0: r0 + len(members) * k])
            else:
                for o in members:
                    N = ctx.obj_npts[o]
                    feat = rr["feat"][: k * N * 16].view(k, 2, N, 8)
                    ctx.features_f32a(o, P[o], out=feat)
                    row = plan.rr_row[o] * k
                    ctx.pool_f32a(ws, feat, out=rr["pooled"][row: row + k])
            ctx.head(ws, rr["pooled"][r0: r0 + len(members) * k], False, out=rr["scores"][r0: r0 + len(members) * k])
        nk = n_obj * k
        S2, I2 = rr["out"][:nk].view(torch.float32).view(n_obj, k), rr["out"][nk: 2 * nk].view(n_obj, k)
        ctx.topk_segments(rr["scores"], plan.rr_seg, k, index_map=I.reshape(-1), out=(S2, I2))
        self._result_flat = rr["out"][: 2 * nk]
        self._mark(None, 0, t)
        return S2, I2

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
            _, uv, _, _ = ctx.features(r["slot"], p12, dtype=torch.float32, want_uv=True)
            out, st = ctx.icp_refine(p12, ob["_zs_host"][0], uv, max_dist=icp_max_dist, max_iter=max_iter)
            pending.append((o, len(idx), out, st))
        for o, m, out, st in pending:                       # read back after every object's launches are queued
            poses[o, :m, :3, :] = out.cpu().numpy().astype(np.float64).reshape(m, 3, 4)
            stats[o, :m] = st.cpu().numpy()
        return poses, stats

    # -- public end-to-end call -------------------------------------------------------------
    def score_frame(self, img_u8, depth, cam_K, objects: List[dict], weight_of=lambda o: 0):
        """Host buffers in, host results out: ``(scores (n_obj,k), indices (n_obj,k))`` numpy arrays.  A frame with more
        objects than the context has cloud slots (64) is scored in groups of 64."""
        if len(objects) > MAX_OBJECTS:
            parts = [self.score_frame(img_u8, depth, cam_K, objects[g: g + MAX_OBJECTS], lambda o, g=g: weight_of(g + o))
                     for g in range(0, len(objects), MAX_OBJECTS)]
            return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
        self.upload(img_u8, depth, cam_K, objects, weight_of)
        S, I = self.run_resident()
        return S.cpu().numpy(), I.cpu().numpy()

    def _pinned(self, n_ints: int, slot: int):
        key = (n_ints, slot)
        if key not in self._pin:
            self._pin[key] = torch.empty((n_ints,), dtype=torch.int32, pin_memory=True)
        return self._pin[key]

    def score_frames(self, frames: List[dict], weight_of=lambda o: 0, depth: int = 2):
        """Stream of frames (BASELINE.json config 5: multi-frame scoring between finetune steps).

        ``frames``: dicts with ``img`` (uint8), ``depth``, ``cam_K``, ``objects``.  Uploads, kernels and the
        read-back of each frame's top-k are all asynchronous; the host only blocks when ``depth`` frames are in
        flight, and the pose hypotheses of frame f+1 (most of a frame's bytes) travel on a side stream while the kernels
        of frame f run.  Every device and pinned buffer is owned by this object and reused.  Returns a list of
        ``(scores (n_obj,k), indices (n_obj,k))`` numpy pairs, one per frame.  The free-space pre-filter
        (inconst_ratio_th < 100) keeps its counts on the device on the tensor-core path; only the fp32 parity path
        synchronises once per object to read the kept count.
        """
        stream = torch.cuda.current_stream(self.ctx.device)
        inflight, out = [], []

        def drain():
            ev, host, n_obj = inflight.pop(0)
            ev.synchronize()
            nk = n_obj * self.k
            a = host.numpy()
            out.append((a[:nk].view(np.float32).reshape(n_obj, self.k).copy(), a[nk: 2 * nk].reshape(n_obj, self.k).copy()))

        nxt = self.prefetch_poses(frames[0]["objects"], weight_of) if frames else None
        for f, fr in enumerate(frames):
            self.upload(fr["img"], fr["depth"], fr["cam_K"], fr["objects"], weight_of, prefetched=nxt)
            nxt = self.prefetch_poses(frames[f + 1]["objects"], weight_of) if f + 1 < len(frames) else None    # overlaps this frame's kernels
            S, I = self.run_resident()
            n_obj = S.shape[0]
            host = self._pinned(2 * n_obj * self.k, f % (depth + 1))
            host.copy_(self._result_flat, non_blocking=True)      # S and I are the two halves of one device buffer
            ev = torch.cuda.Event()
            ev.record(stream)
            inflight.append((ev, host, n_obj))
            if len(inflight) >= depth:
                drain()
        while inflight:
            drain()
        return out
