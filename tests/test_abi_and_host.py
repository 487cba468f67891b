"""CPU-side checks: the C-ABI library builds, loads and exports what include/zs.h declares; host logic."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from ossid_code_b200 import _lib, scoring, weights, zephyr_shim
from ossid_code_b200.engine import poses_to_rt12

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "zs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    declared = _header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/zs.h but not exported by libzs.so"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes signature table and header disagree"
    bound = _lib.load()
    assert bound.zs_version() == 200
    assert bound.zs_strerror(0) == b"ok" and b"CUDA" in bound.zs_strerror(-2)


def test_library_is_sm100a_only_with_lineinfo(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_weight_blob_size_matches_header():
    src = open(os.path.join(ROOT, "include", "zs.h")).read()
    expr = re.search(r"#define ZS_WEIGHT_FLOATS \((.*)\)", src).group(1)
    assert eval(expr) == _lib.ZS_WEIGHT_FLOATS == sum(t.numel() for t in weights.seeded_folded(0).values())


def test_create_without_gpu_fails_loudly(lib_path):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ossid_code_b200.engine import ZsContext
    with pytest.raises(_lib.ZsError):
        ZsContext(0)
    h = ctypes.c_void_p()
    assert _lib.load().zs_create(ctypes.byref(h), 0) < 0 and not h.value


def test_bn_folding_matches_torch_eval():
    sd = weights.seeded_state_dict(3)
    m = zephyr_shim.PointNet2SSG(8, None, 1)
    m.load_state_dict(sd)
    f = weights.fold_state_dict(m.state_dict())
    x = torch.randn(5, 8, 40)
    with torch.no_grad():
        ref = torch.relu(m.bn1(m.conv1(x)))
        ours = torch.relu(torch.einsum("oc,bcn->bon", f["W1"], x) + f["b1"][None, :, None])
        torch.testing.assert_close(ours, ref, rtol=1e-5, atol=1e-5)
        g = torch.randn(7, 1024)
        torch.testing.assert_close(torch.relu(g @ f["F1"].T + f["c1"]), torch.relu(m.bn_fc1(m.fc1(g))), rtol=1e-4, atol=1e-4)
    pref = weights.fold_state_dict({"model.net." + k: v for k, v in sd.items()})   # Lightning-style prefixes
    assert all(torch.equal(pref[k], f[k]) for k in f)


def test_shim_installs_zephyr_modules():
    zephyr_shim.install()
    from zephyr.datasets.score_dataset import ScoreDataset
    from zephyr.models.pointnet2 import PointNet2SSG
    from zephyr.utils import projectPointsUv
    ds = ScoreDataset([], "", "lmo", type("A", (), {"inconst_ratio_th": 10})(), mode="test")
    assert ds.dim_point == 8 and ds.inconst_ratio_th == 10.0 and callable(projectPointsUv)
    model = PointNet2SSG(ds.dim_point, None, num_class=1).eval()
    assert model.device.type == "cpu"
    with pytest.raises(RuntimeError):
        model({"point_x": torch.zeros(1, 4, 8)})            # no CPU scoring path


def test_poses_to_rt12():
    T = np.tile(np.eye(4), (3, 1, 1))
    T[:, :3, 3] = [[1, 2, 3], [4, 5, 6], [7, 8, 9]]
    p = poses_to_rt12(T, "cpu")
    assert p.dtype == torch.float32 and p.shape == (3, 12)
    assert p[1].tolist() == [1, 0, 0, 4, 0, 1, 0, 5, 0, 0, 1, 6]
    with pytest.raises(ValueError):
        poses_to_rt12(np.zeros((3, 3, 4)), "cpu")


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 1000, 50001):
        for world in (1, 2, 3, 4, 8):
            spans = [scoring.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_merge_topk_ordering():
    s = torch.tensor([[1.0, 5.0, 5.0, float("nan"), 2.0, 9.0]])
    i = torch.tensor([[7, 3, 1, 0, 4, -1]], dtype=torch.int32)
    ms, mi = scoring.merge_topk(s, i, 4)
    assert mi.tolist() == [[1, 3, 4, 7]] and ms.tolist() == [[5.0, 5.0, 2.0, 1.0]]
    ms, mi = scoring.merge_topk(s[:, :2], i[:, :2], 4)
    assert mi.tolist() == [[3, 7, -1, -1]]


def _record(S, I, info):
    return torch.cat([S.to(torch.float32).contiguous().view(torch.int32).reshape(-1), I.to(torch.int32).reshape(-1),
                      torch.tensor(info, dtype=torch.int32).reshape(-1)])


def test_merge_records_applies_the_never_empty_rule_to_the_whole_list():
    """a rank whose slice was entirely rejected by the pre-filter contributes only its never-empty fallback;
    that candidate must not enter the merge when another rank kept something, and when NO rank kept anything only the
    first minimum-violation fallback survives (what a single-GPU run keeps)."""
    k = 3
    # object 0: rank 0 kept 2 hypotheses, rank 1 kept none (its fallback scores highest and must still lose)
    # object 1: no rank kept anything; fallbacks have 9 and 4 violations -> rank 1's survives
    # object 2: no rank kept anything, equal violation counts -> the lower rank (= lower index) survives
    S0 = torch.tensor([[1.0, 0.5, float("-inf")], [7.0, float("-inf"), float("-inf")], [2.0, float("-inf"), float("-inf")]])
    I0 = torch.tensor([[3, 1, -1], [0, -1, -1], [5, -1, -1]])
    S1 = torch.tensor([[9.0, float("-inf"), float("-inf")], [6.0, float("-inf"), float("-inf")], [8.0, float("-inf"), float("-inf")]])
    I1 = torch.tensor([[12, -1, -1], [14, -1, -1], [11, -1, -1]])
    g = torch.stack([_record(S0, I0, [[2, 0], [0, 9], [0, 6]]), _record(S1, I1, [[0, 5], [0, 4], [0, 6]])])
    S, I = scoring.merge_records_reference(g, 3, k)
    assert I.tolist() == [[3, 1, -1], [14, -1, -1], [5, -1, -1]]
    assert S[0, :2].tolist() == [1.0, 0.5] and S[1, 0] == 6.0 and S[2, 0] == 2.0
    # without the info section saying otherwise ({1,0} = unfiltered), everything merges by score
    g2 = torch.stack([_record(S0, I0, [[1, 0]] * 3), _record(S1, I1, [[1, 0]] * 3)])
    S, I = scoring.merge_records_reference(g2, 3, k)
    assert I.tolist() == [[12, 3, 1], [0, 14, -1], [11, 5, -1]]


def test_record_layout_keeps_poses_aligned():
    for n_obj, k in ((21, 8), (1, 1), (3, 5), (8, 64)):
        at = scoring.record_pose_offset(n_obj, k)
        assert at % 4 == 0 and at >= 2 * n_obj * k + 2 * n_obj
        assert scoring.record_ints(n_obj, k, True) == at + 12 * n_obj * k and scoring.record_ints(n_obj, k, False) == at


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    scores = torch.randn(3, 40, generator=g)
    scores[1, 5] = scores[1, 33] = scores[1].max() + 1          # tie across ranks: lower index must win
    lo, hi = scoring.shard_range(40, rank, world)
    k = 4
    loc_s, loc_i = [], []
    for o in range(3):
        s = scores[o, lo:hi]
        order = sorted(range(hi - lo), key=lambda j: (-float(s[j]), j))[:k]
        loc_s.append(s[order])
        loc_i.append(torch.tensor(order, dtype=torch.int32) + lo)
    S, I = scoring.allgather_topk(torch.stack(loc_s), torch.stack(loc_i), k)
    q.put((rank, S.tolist(), I.tolist()))
    dist.destroy_process_group()


def test_allgather_topk_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(30) for p in procs]
    assert res[0][1:] == res[1][1:], "ranks disagree after the merge"
    g = torch.Generator().manual_seed(0)
    scores = torch.randn(3, 40, generator=g)
    scores[1, 5] = scores[1, 33] = scores[1].max() + 1
    for o in range(3):
        exp = sorted(range(40), key=lambda j: (-float(scores[o, j]), j))[:4]
        assert res[0][2][o] == exp
    assert res[0][2][1][:2] == [5, 33]


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (CPU restatement, no GPU needed): exactly one JSON line on stdout with the keys the
    driver reads; the GPU arm prints the same keys plus roofline / clocks / gpu_launches."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "16"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["workload"].startswith("c2")
    # both arms print the same `config` object (the driver's same_config check)
    sys.path.insert(0, root)
    import argparse
    import bench
    ns = argparse.Namespace(workload="c2", k=8, inconst_th=100.0, precision="bf16", no_fuse=False)
    assert d["config"] == bench.config_for(ns, 1, "weak")
