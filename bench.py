"""Benchmark of the Zephyr hypothesis-scoring hot path (BASELINE.json metric: hypotheses scored/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4|c5]
                    [--precision bf16|fp32] [--scaling weak|strong] [--no-rerank]

A "step" is one pass of the hot path over one synthetic frame: for every object, project the model
points under every pose hypothesis, gather, featurise, score with the MLP, take the per-object top-k,
(for N>1) all-gather and merge the k candidates, and re-score the k candidates with the fp32-accurate
scorer.  N=1 workload = BASELINE.json configs[1] (YCB-V-shaped frame, 21 objects x 10,000 hypotheses x
1,000 points).  For N>1 the headline keeps that per-GPU load (weak scaling: objects carry 10,000*N
hypotheses, sharded contiguously); the same run also shards the FIXED C3 and C2 frames over the N ranks
(strong scaling, BASELINE.json configs[2]), lets rank 0 score each whole frame alone and requires the
NCCL-merged result to be bit-identical: the `strong` / `strong_c2` sub-records and
`sharded_equals_single`.  `--scaling strong` makes the fixed frame the headline instead.

One JSON line on stdout (rank 0).  `value` = whole-job hypotheses/s with inputs resident in HBM;
`e2e` = the same through the public host-buffer API (H2D of frame and poses and D2H of the top-k inside the
timed region, median of three repeats; the static model clouds are uploaded once, as the reference loads
them once).  `--impl reference` times the CPU restatement of the reference path (oracle/, torch CPU, all
host threads; driven by the reference's own unmodified networkInference when /root/reference is present)
on one whole object of the same workload per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # name: (intrinsics, n_obj, hypotheses per object per GPU, points per object, description)
    "c1": ("lmo", 1, 1000, 1000, "synthetic 640x480, 1 object x 1,000 hypotheses x 1,000 pts"),
    "c2": ("ycbv", 21, 10000, 1000, "YCB-V-shaped frame: 21 objects x 10,000 hypotheses x 1,000 pts"),
    "c3": ("lmo", 8, 50000, 1000, "LM-O-shaped frame: 8 objects x 50,000 hypotheses x 1,000 pts, all inside DTOID-style box crops"),
    # the same frame BEFORE the detector's crop is applied: the hypothesis mixture is unfiltered, every object carries its
    # DTOID-style box, and box -> mask rasterisation + the mask-overlap pre-filter run inside the timed step
    "c3f": ("lmo", 8, 50000, 1000, "LM-O-shaped frame: 8 objects x 50,000 unfiltered hypotheses + DTOID-style boxes; box->mask and mask pre-filter inside the step"),
    "c4": ("hd", 1, 200000, 4000, "bandwidth stress: 1280x720, 1 object x 200,000 hypotheses x 4,000 pts"),
    # a step = the 32 frames scored between two finetune rounds (online_learning.py: finetune_interval); frames are C2-shaped
    "c5": ("ycbv", 21, 10000, 1000, "online-learning stream: 32 frames x (21 objects x 10,000 hypotheses x 1,000 pts) per step"),
}
FRAMES_PER_STEP = {"c5": 32}
MLP_MACS_PER_POINT = 8 * 64 + 64 * 128 + 128 * 1024
HEAD_MACS = 1024 * 512 + 512 * 256 + 256


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def make_workload(name, n_gpus, seed=1, gpu=True):
    """Synthetic frame for `name`; hypotheses per object scale with the GPU count (weak scaling)."""
    from ossid_code_b200 import synthetic as syn
    intr, n_obj, per_gpu, n_pts, _ = WORKLOADS[name]
    sc = syn.make_scene(seed, intr, n_obj=n_obj, n_pts=n_pts, n_hypo=min(per_gpu, 10000))
    reps = -(-per_gpu * n_gpus // min(per_gpu, 10000))
    rng = np.random.default_rng(seed + 99)
    for ob in sc["objects"]:
        if name == "c3":
            ob["pose_hypos"] = crop_hypotheses(sc, ob, rng, per_gpu * n_gpus, gpu)
        elif reps > 1:   # more hypotheses of the same mixture, freshly drawn
            extra = [syn.make_hypotheses(rng, ob["gt_pose"], min(per_gpu, 10000), sc["cam_K"], sc["H"], sc["W"])
                     for _ in range(reps - 1)]
            ob["pose_hypos"] = np.concatenate([ob["pose_hypos"], *extra])[: per_gpu * n_gpus]
        if name == "c3f":    # detector output for this object: its box (before expandBox) and a confident score
            ob["boxes"] = np.asarray([syn.gt_box(sc, ob, 1.0)], dtype=np.float64)
            ob["box_scores"] = np.asarray([0.9])
    return sc


def crop_hypotheses(sc, ob, rng, n, gpu=True):
    """C3: every hypothesis lies inside the object's DTOID-style box crop (GT box grown by expandBox's 1.2,
    python/ossid/utils/__init__.py:11-16): draw from the usual mixture and keep what passes the reference's
    filterHypoByMask(th=0.5) (python/ossid/utils/zephyr_utils.py:49-71; here the fused zs_mask_count kernel).
    Workload preparation, outside every timed region."""
    from ossid_code_b200 import synthetic as syn, zephyr_utils as glue
    x1, y1, x2, y2 = syn.gt_box(sc, ob, 1.2)
    mask = np.zeros((sc["H"], sc["W"]), np.uint8)
    mask[y1:y2, x1:x2] = 1
    meta, kept, have = glue.K2meta(sc["cam_K"]), [], 0
    while have < n:
        cand = syn.make_hypotheses(rng, ob["gt_pose"], 20000, sc["cam_K"], sc["H"], sc["W"])
        if gpu:
            keep = glue.filterHypoByMask(ob["model_points"], meta, cand, mask, th=0.5, device=torch.cuda.current_device())
        else:                                     # reference arm: the CPU restatement of the same filter
            from oracle import zephyr_oracle as zo
            keep = zo.mask_filter(zo.project_raw(cand, ob["model_points"], meta), torch.from_numpy(mask.astype(np.int64)),
                                  len(ob["model_points"]), th=0.5).numpy()
        cand = cand[np.asarray(keep)]
        kept.append(cand)
        have += len(cand)
    return np.concatenate(kept)[:n]


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples while the timed region runs (B200_PROFILING.md).

    Started before the warm-up (nvidia-smi needs a few hundred ms to produce its first line); samples are
    time-stamped on arrival and only those inside [mark_start, mark_stop] are summarised.
    """
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.rows, self.t0, self.t1 = gpu_index, None, [], None, None

    def _reader(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line))

    def start(self):
        import threading
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True, bufsize=1)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._reader, daemon=True)
        self.thread.start()
        t_end = time.perf_counter() + 5.0
        while not self.rows and time.perf_counter() < t_end:      # wait for the first sample
            time.sleep(0.02)

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_stop(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=10)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        inside = [l for (t, l) in self.rows if self.t0 is not None and self.t0 <= t <= self.t1 + 0.06]
        scope = "timed region"
        if len(inside) < 2:                                        # very short run: fall back to everything under load
            inside, scope = [l for (_, l) in self.rows], "warm-up + timed region"
        sm, mx, pw, reasons = [], [], [], set()
        for line in inside:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


def reference_glue():
    """The glue function that drives the CPU arm: the reference's own, unmodified ``networkInference`` when the
    reference checkout is present (this container), else this repo's mirror of it (the GPU box).  -> (fn, name)."""
    if os.path.exists("/root/reference/python/ossid/utils/zephyr_utils.py"):
        try:
            from oracle import gen_golden, zephyr_oracle as zo
            ref = gen_golden.import_reference(lambda poses, pts, meta: zo.project_raw(poses, pts, meta).numpy())
            return ref.networkInference, "reference networkInference (unmodified, /root/reference)"
        except Exception as exc:                       # missing optional import inside the reference tree
            sys.stderr.write(f"reference glue not importable ({exc}); using the mirror\n")
    from ossid_code_b200 import zephyr_utils as glue
    return glue.networkInference, "mirror of networkInference (ossid_code_b200/zephyr_utils.py; /root/reference absent)"


def cpu_reference_run(sample, n_repeat):
    """Time the reference-style CPU path: networkInference driving the torch-CPU oracle objects, span = featurise +
    score as the reference defines it (zephyr_utils.py:29-37)."""
    from oracle import zephyr_oracle as zo
    from ossid_code_b200 import weights
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    infer, glue_name = reference_glue()
    ds, model = zo.OracleScoreDataset(100.0), zo.OracleScorer(weights.seeded_folded(0))
    times, n_scored = [], 0
    for _ in range(n_repeat):
        data = dict(sample)
        data["pose_hypos"] = sample["pose_hypos"].copy()
        poses, scores, _, _, dt = infer(model, ds, data, return_time=True)
        times.append(dt)
        n_scored = len(scores)
    return times, n_scored, cores, glue_name


def cpu_sample(sc, n_hypo):
    ob = sc["objects"][0]
    return dict(img=sc["img"], depth=sc["depth"], cam_K=sc["cam_K"], model_points=ob["model_points"],
                model_colors=ob["model_colors"], model_normals=ob["model_normals"],
                pose_hypos=ob["pose_hypos"][:n_hypo])


def config_for(args, world, scaling):
    """The `config` object; identical for --impl ours and --impl reference (the reference arm's bounded sample is
    described in its cpu_baseline.sample)."""
    intr, n_obj, per_gpu, n_pts, desc = WORKLOADS[args.workload]
    n_frames = FRAMES_PER_STEP.get(args.workload, 1)
    mult = world if scaling == "weak" else 1
    return {"workload": f"{args.workload}: {desc}", "hypotheses_per_step": n_obj * per_gpu * mult * n_frames,
            "frames_per_step": n_frames, "objects": n_obj, "points_per_object": n_pts, "topk": args.k,
            "inconst_ratio_th": args.inconst_th,
            "kernels": ("fused projection+gather+features+MLP+max-pool kernel" if (args.precision == "bf16" and not args.no_fuse)
                        else "zs_features -> zs_pool (features through HBM)"),
            "parallelism": (f"hypothesis-sharded x{world}, one all-gather of top-k records" if world > 1 else "single GPU"),
            "l2": "feature chunks of 32768 hypotheses x 1000 pts (>= 0.5 GB) exceed the 126 MB L2; no flush needed",
            "weights": "seeded random (no checkpoint is published)"}


def run_reference(args):
    """--impl reference: the CPU restatement on one whole object of the same workload per step; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    intr, n_obj, per_gpu, n_pts, desc = WORKLOADS[args.workload]
    sc = make_workload(args.workload, 1, gpu=False)
    n_s = min(args.cpu_sample if args.cpu_sample else per_gpu, per_gpu)
    sample = cpu_sample(sc, n_s)
    if args.warmup:                                    # untimed warm-up steps run on a tenth of the sample
        cpu_reference_run(cpu_sample(sc, max(n_s // 10, 1)), args.warmup)
    times, n_scored, cores, glue_name = cpu_reference_run(sample, args.steps)
    total = sum(times)
    value = n_scored * len(times) / total
    sample_desc = (f"per step: all {n_s} hypotheses x {n_pts} pts of object 0 of the workload's frame, i.e. one reference "
                   f"networkInference call (warm-up steps: {max(n_s // 10, 1)}); oracle port (torch CPU fp32) driven by {glue_name}")
    line = {
        "impl": "reference", "metric": "hypotheses_scored_per_sec", "value": value, "unit": "hypotheses/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_for(args, args.gpus, args.scaling),
        "cpu_baseline": {"value": value, "unit": "hypotheses/s", "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "hypotheses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL prints its version line) write to fd 1; the contract is ONE JSON line on stdout.
    Point fd 1 at stderr for the duration of the run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


class Runner:
    """One scorer + one workload on this rank: device-resident timing and end-to-end timing."""

    def __init__(self, args, sc, local, world_view=None):
        from ossid_code_b200 import scoring, weights
        self.scoring = scoring
        self.args, self.sc, self.dev = args, sc, torch.device("cuda", local)
        w = [weights.seeded_folded(0), weights.seeded_folded(1)]
        self.fs = scoring.FrameScorer(w, device=local, precision=args.precision, inconst_ratio_th=args.inconst_th,
                                      k=args.k, rerank=not args.no_rerank, fused=not args.no_fuse)
        self.fs.forced_rank_world = world_view      # (0, 1): score the whole frame on this rank alone, no collective
        self.weight_of = (lambda o: o % 2)          # two scorers keyed on object parity, online_learning.py:461-463
        self.frame = None

    def upload(self):
        sc = self.sc
        self.fs.upload(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], self.weight_of)

    def resident(self, n_iter, record_stages=False):
        """n_iter back-to-back run_resident() calls between two CUDA events -> (ms, S, I, launches, stages)."""
        fs = self.fs
        if record_stages:
            fs.stage_events = []
        l0 = fs.ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_iter):
            S, I = fs.run_resident()
        e1.record()
        torch.cuda.synchronize(self.dev)
        stages = fs.stage_times_ms() if record_stages else {}
        fs.stage_events = None
        return e0.elapsed_time(e1), S, I, fs.ctx.launches - l0, stages

    def host_frame(self):
        if self.frame is None:
            sc = self.sc
            pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            objs = [dict(model_points=pin(ob["model_points"]), model_colors=pin(ob["model_colors"]),
                         model_normals=pin(ob["model_normals"]),
                         pose_hypos=pin(ob["pose_hypos"]),          # float64 (M,4,4), as the reference hands them over
                         **{k: ob[k] for k in ("boxes", "box_scores") if k in ob}) for ob in sc["objects"]]
            self.frame = dict(img=pin(sc["img"]), depth=pin(sc["depth"]), cam_K=sc["cam_K"], objects=objs)
        return self.frame

    def e2e(self, n_frames):
        """score_frames over n_frames copies of the frame: wall seconds including every H2D / D2H copy."""
        frame = self.host_frame()
        torch.cuda.synchronize(self.dev)
        t0 = time.perf_counter()
        results = self.fs.score_frames([frame] * n_frames, self.weight_of)
        torch.cuda.synchronize(self.dev)
        return time.perf_counter() - t0, results[-1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = hypotheses per object grow with N (per-GPU load fixed); strong = the frame is fixed "
                         "and sharded over the N ranks")
    ap.add_argument("--strong-workloads", default="c3,c2",
                    help="N>1: fixed frames that are additionally sharded over the ranks and checked against rank 0 "
                         "scoring them alone (sub-records `strong`, `strong_c2`); empty = skip")
    ap.add_argument("--no-rerank", action="store_true", help="skip the fp32-accurate re-scoring of the top-k candidates")
    ap.add_argument("--no-fuse", action="store_true",
                    help="bf16 path: zs_features + zs_pool as two kernels (features through HBM) instead of the fused kernel")
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="hypotheses per step of the CPU arm (0 = one whole object of the workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inconst-th", type=float, default=100.0,
                    help="free-space pre-filter threshold in percent (reference: 100 = off for LM-O, 10 for YCB-V, "
                         "online_learning.py:174,184); the headline keeps it off so that every hypothesis is scored")
    args = ap.parse_args()
    quiet_stdout()

    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from ossid_code_b200 import scoring

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    intr, n_obj, per_gpu, n_pts, desc = WORKLOADS[args.workload]
    scaling = args.scaling if world > 1 else "weak"
    sc = make_workload(args.workload, world if scaling == "weak" else 1)
    run = Runner(args, sc, local)
    fs = run.fs
    n_frames = FRAMES_PER_STEP.get(args.workload, 1)
    total_hyp = sum(len(ob["pose_hypos"]) for ob in sc["objects"]) * n_frames
    local_hyp = sum(scoring.shard_range(len(ob["pose_hypos"]), rank, world)[1]
                    - scoring.shard_range(len(ob["pose_hypos"]), rank, world)[0] for ob in sc["objects"]) * n_frames

    # ---- device-resident arm: `value` ------------------------------------------------------------
    run.upload()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    run.resident(args.warmup)
    barrier()
    sampler.mark_start()
    ms, S, I, launches, stages = run.resident(args.steps * n_frames, record_stages=True)
    barrier()
    sampler.mark_stop()
    ms_total = max_over_ranks(ms)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    # the same K steps replayed from a CUDA graph (single GPU; what score_frames does): no per-stage events inside, so
    # it is reported beside `value`, whose timed region carries the events the roofline numbers come from
    graph_ms = None
    if world == 1 and fs.use_graph:
        run.resident(3 * n_frames)
        graph_ms = run.resident(args.steps * n_frames)[0] / args.steps
        if not fs._graphs:
            graph_ms = None
    value = total_hyp / (ms_step * 1e-3)
    scored = int(fs.last_scored) * n_frames if fs._plan.filtered else total_hyp

    # ---- end-to-end arm through the public host-buffer API: `e2e` ----------------------------------
    # per frame: H2D of the uint8 image, the float32 depth and this rank's pose hypotheses (float64 (M,4,4) blocks as the
    # reference hands them over, cast to float32 rows on the device), D2H of the top-k; the model
    # clouds are static assets (the reference loads them once, online_learning.py:303-311): uploaded by the warm-up only
    run.e2e(max(5, args.warmup))                                  # warm-up frames (build the pinned cloud copies, size the rings)
    e2e_runs = []
    for _ in range(3):
        barrier()
        sec, (Sh, Ih) = run.e2e(args.steps * n_frames)
        e2e_runs.append(max_over_ranks(sec) / args.steps)
    e2e_s = statistics.median(e2e_runs)
    frame = run.host_frame()
    pose_bytes = frame["objects"][0]["pose_hypos"].element_size() * 16
    h2d = n_frames * (frame["img"].numel() + frame["depth"].numel() * 4) + local_hyp * pose_bytes
    d2h = int(Sh.size * 4 + Ih.size * 4) * n_frames
    if not (np.array_equal(Ih, I.cpu().numpy()) and np.array_equal(Sh, S.cpu().numpy())):
        raise SystemExit("end-to-end result differs from the device-resident result")

    # ---- roofline of the dominant kernel + the feature kernel --------------------------------------
    peaks = load_peaks()
    fbytes = 2 if args.precision == "bf16" else 4
    timed_ms = ms_step * args.steps
    roof = {}
    if "pool" in stages:
        st = stages["pool"]
        if fs._plan.filtered:      # launches are sized by capacity; the MLP runs on the hypotheses that passed the pre-filter
            st = dict(st, units=scored * n_pts * args.steps)
        flops = 2.0 * MLP_MACS_PER_POINT * st["units"]
        ach = flops / (st["ms"] * 1e-3) / 1e12
        peak = peaks["tensor_sustained"]
        # what the tensor pipe really executes: K of layer 1 padded 8 -> 16, points padded to whole 256-point pair-tiles
        # (x3 MMAs per product in the fp32-accurate bf16-split mode)
        n_pad = -(-n_pts // 256) * 256
        hw_macs = (16 * 64 + 64 * 128 + 128 * 1024) * (n_pad / n_pts) * (3 if args.precision == "fp32" else 1)
        hw_flops = 2.0 * hw_macs * st["units"]
        sm_mhz = (clocks or {}).get("sm_mhz")
        roof["roofline"] = {"kernel": "zs_pool (shared MLP 8-64-128-1024 + max-pool)", "bound": "tensor",
                            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                            "peak_source": peaks["source"] + ", sustained bf16 (cuBLAS 8192^3 back to back)",
                            "peak_burst": peaks["tensor_burst"], "frac_of_burst": ach / peaks["tensor_burst"],
                            "peak_nominal_dense": 2250.0, "frac_of_nominal": ach / 2250.0,
                            "hardware_tflops_incl_padding": hw_flops / (st["ms"] * 1e-3) / 1e12,
                            "frac_of_clock_scaled_peak": (hw_flops / (st["ms"] * 1e-3)) / (148 * 8192 * sm_mhz * 1e6) if sm_mhz else None,
                            "clock_scaled_peak": "148 SMs x 8192 dense bf16 FLOP/clk x the median SM clock sampled in the timed region",
                            "note": "a fraction above 1 means this kernel sustains more than the measured cuBLAS GEMM does: "
                                    "both are power-limited, and the fused MLP moves fewer operand bytes per MAC "
                                    "(weights resident in shared memory, activations never leave the SM); "
                                    "frac_of_clock_scaled_peak is the interpretable one",
                            "launch_groups": st["calls"],
                            "ms_in_timed_region": st["ms"], "share_of_step": st["ms"] / timed_ms}
    if "features" in stages:
        st = stages["features"]
        hyp_units = st["units"] / n_pts
        byts = hyp_units * (48 + n_pts * (8 * fbytes))       # poses in + features out (mask/uv not written on this path)
        ach = byts / (st["ms"] * 1e-3) / 1e9
        roof["roofline_features"] = {"kernel": "zs_features (projection + gather + residual features)", "bound": "hbm",
                                     "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                                     "traffic": None, "peak_source": peaks["source"], "launch_groups": st["calls"],
                                     "ms_in_timed_region": st["ms"], "share_of_step": st["ms"] / timed_ms}
    for key in ("prefilter", "head", "topk", "rerank", "allgather+merge"):
        if key in stages:
            roof[f"{key}_share_of_step"] = stages[key]["ms"] / timed_ms
            roof[f"{key}_ms_per_step"] = stages[key]["ms"] / args.steps
    # The same feature kernel writing fp32 features (the 1e-4 parity configuration), timed alone on this
    # rank's hypotheses of object 0 replicated to >= 32768 (output 1 GB, far beyond L2): context for the
    # in-step bf16 figure above, which moves half the bytes per point and is issue-bound instead.
    r0 = fs._resident[0]
    reps = max(1, -(-32768 // max(r0["poses12"].shape[0], 1)))
    p_big = r0["poses12"].repeat(reps, 1)[:32768].contiguous()
    buf = torch.empty((p_big.shape[0], n_pts, 8), dtype=torch.float32, device=dev)
    torch.cuda.synchronize(dev)
    time.sleep(2.0)      # "standalone": let the clocks recover from the power-capped tensor-core phase above
    for _ in range(3):
        fs.ctx.features(r0["slot"], p_big, out=buf)
    torch.cuda.synchronize(dev)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(5):
        fs.ctx.features(r0["slot"], p_big, out=buf)
    k1.record()
    torch.cuda.synchronize(dev)
    k_ms = k0.elapsed_time(k1) / 5
    k_bytes = p_big.shape[0] * (48 + n_pts * 32)
    roof["roofline_features_fp32"] = {"kernel": "zs_features, fp32 features, standalone launch of 32768 hypotheses (NOT in the timed step)",
                                      "bound": "hbm", "achieved": k_bytes / (k_ms * 1e-3) / 1e9, "peak": peaks["hbm"],
                                      "unit": "GB/s", "frac": k_bytes / (k_ms * 1e-3) / 1e9 / peaks["hbm"],
                                      "ms_per_launch": k_ms, "hypotheses_per_s": p_big.shape[0] / (k_ms * 1e-3)}
    del buf, p_big

    for tname in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):                     # measured DRAM traffic per launch, valid for the profiled launch shape only
            tr = json.load(open(tpath))
            if tr.get("workload") == args.workload and tr.get("precision") == args.precision and world == 1:
                for key, stage in (("roofline", "zs_pool"), ("roofline_features", "zs_features")):
                    if key in roof and stage in tr:
                        roof[key]["traffic"] = tr[stage]["dram_read_bytes"] + tr[stage]["dram_write_bytes"]
                        roof[key]["traffic_source"] = f"profiles/{tname} (ncu --set full, bytes of the first captured launch: 10,000 hypotheses for the feature kernel, a 32,768-hypothesis chunk for the MLP kernel)"
            break
    if "roofline" in roof and "pool" in stages:
        # HBM side of the dominant kernel (it is tensor-bound; this is what `traffic` is to be compared with): fused = pose in,
        # pooled vector out; two-kernel path = feature rows in, pooled vector out
        hyp_per_launch = stages["pool"]["units"] / n_pts / stages["pool"]["calls"]
        per_hyp = (48 + 4096) if "features" not in stages else (n_pts * 8 * fbytes + 4096)
        roof["roofline"]["algorithmic_hbm_bytes_per_launch"] = per_hyp * hyp_per_launch
        roof["roofline"]["hypotheses_per_launch"] = hyp_per_launch
    if "roofline_features" in roof:
        roof["roofline_features"]["algorithmic_bytes_per_launch"] = \
            (48 + n_pts * 8 * fbytes) * stages["features"]["units"] / n_pts / stages["features"]["calls"]

    cfg = config_for(args, world, scaling)
    assert cfg["hypotheses_per_step"] == total_hyp, (cfg["hypotheses_per_step"], total_hyp)
    line = {
        "metric": "hypotheses_scored_per_sec", "value": value, "unit": "hypotheses/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": cfg, "hypotheses_passing_prefilter": scored,
        "top1": ("fp32-accurate re-rank of the k candidates of every object (unconditional top-1)" if fs.rerank
                 else "ordered by the scorer's own precision"),
        "e2e": {"value": total_hyp / e2e_s, "unit": "hypotheses/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                "ms_per_step_repeats": {"min": min(e2e_runs) * 1e3, "median": e2e_s * 1e3, "max": max(e2e_runs) * 1e3, "n": len(e2e_runs)},
                "vs_resident": e2e_s * 1e3 / ms_step},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if graph_ms is not None:
        line["cuda_graph"] = {"ms_per_step": graph_ms, "value": total_hyp / (graph_ms * 1e-3),
                              "note": "the step replayed from a CUDA graph (FrameScorer's default on one GPU, used by the e2e arm)"}
    line.update(roof)

    # ---- strong scaling of fixed frames + on-hardware identity of the sharded result (N > 1) --------
    if world > 1:
        names = [w for w in args.strong_workloads.split(",") if w]
        if scaling == "strong" and args.workload not in names:
            names.insert(0, args.workload)
        all_equal = True
        for wname in names:
            rec = strong_record(args, wname, local, rank, world, barrier, max_over_ranks)
            if rank == 0:
                all_equal &= bool(rec["sharded_equals_single"])
                line["strong" if wname == "c3" else f"strong_{wname}"] = rec
        if names and rank == 0:
            line["sharded_equals_single"] = all_equal

    if rank == 0 and not args.no_cpu_baseline:
        n_s = min(args.cpu_sample if args.cpu_sample else per_gpu, per_gpu)
        cpu_reference_run(cpu_sample(sc, max(n_s // 10, 1)), 1)
        times, n_scored, cores, glue_name = cpu_reference_run(cpu_sample(sc, n_s), 1)
        line["cpu_baseline"] = {"value": n_scored / statistics.median(times), "unit": "hypotheses/s", "cores": cores,
                                "kind": "port",
                                "sample": f"all {n_s} hypotheses x {n_pts} pts of object 0, one call after a warm-up on {max(n_s // 10, 1)} "
                                          f"(oracle port, torch CPU fp32, driven by {glue_name})"}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def strong_record(args, wname, local, rank, world, barrier, max_over_ranks):
    """Strong scaling of the FIXED frame `wname` (BASELINE.json configs[2] for c3): the same frame sharded over the
    `world` ranks vs rank 0 scoring it alone, plus the bit-for-bit comparison of the two results."""
    import copy
    sub = copy.copy(args)
    sub.workload = wname
    sc = make_workload(wname, 1)
    total = sum(len(ob["pose_hypos"]) for ob in sc["objects"])
    steps = max(args.steps, 5)
    # the sharded run gets `world` times the steps of the single-GPU run: both timed regions then last equally long, so
    # both run at the clocks of a sustained (power-capped) load instead of comparing a short burst with a long run
    steps_n = steps * world
    sharded = Runner(sub, sc, local)
    sharded.upload()
    sharded.resident(max(args.warmup, 3) * world)
    barrier()
    ms, S, I, launches, stages = sharded.resident(steps_n, record_stages=True)
    ms_n = max_over_ranks(ms) / steps_n
    barrier()
    sharded.e2e(5)
    barrier()
    sec, (Sh, Ih) = sharded.e2e(steps_n)
    e2e_n = max_over_ranks(sec) / steps_n
    rec = None
    if rank == 0:                                   # the whole frame on one GPU: reference result and the N=1 time
        alone = Runner(sub, sc, local, world_view=(0, 1))
        alone.upload()
        alone.resident(max(args.warmup, 3))
        ms1, S1, I1, _, _ = alone.resident(steps)
        ms_1 = ms1 / steps
        equal = bool(torch.equal(S1, S) and torch.equal(I1, I) and np.array_equal(Sh, S1.cpu().numpy())
                     and np.array_equal(Ih, I1.cpu().numpy()))
        step_ms = ms_n
        rec = {"workload": f"{wname}: {WORKLOADS[wname][4]}", "hypotheses_per_step": total, "n_gpus": world,
               "steps": steps_n, "steps_single_gpu": steps,
               "ms_per_step": ms_n, "value": total / (ms_n * 1e-3), "unit": "hypotheses/s",
               "ms_per_step_single_gpu": ms_1, "value_single_gpu": total / (ms_1 * 1e-3),
               "speedup_vs_n1": ms_1 / ms_n, "efficiency_vs_n1": ms_1 / (world * ms_n),
               "e2e_ms_per_step": e2e_n * 1e3, "e2e_value": total / e2e_n,
               "sharded_equals_single": equal,
               "stage_ms_per_step_rank0": {k: v["ms"] / steps_n for k, v in stages.items()},
               "launches_per_step_rank0": launches / steps_n,
               "unaccounted_ms_per_step_rank0": step_ms - sum(v["ms"] for v in stages.values()) / steps_n}
        del alone
    barrier()
    del sharded
    torch.cuda.empty_cache()
    return rec


if __name__ == "__main__":
    main()
