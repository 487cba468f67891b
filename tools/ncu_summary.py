"""Turn the two ncu outputs of the profiling recipe into the tracked summaries under profiles/.

    python tools/ncu_summary.py <launches.csv> <full.ncu-rep> <round-tag> "<profiled command>"

<launches.csv>: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ...`
<full.ncu-rep>: `ncu --set full --clock-control none --import-source on -k regex:... -o ...`
Writes profiles/<tag>_launches_c2.csv (copy), profiles/<tag>_ncu_summary.txt and profiles/<tag>_traffic.json.
"""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag, cmd = sys.argv[1:5]
out_dir = os.path.join(ROOT, "profiles")

rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e6
total = sum(a[1] for a in agg.values())
lines = [f"== ncu launch list: `{cmd}`, first {len(rows) - 1} launches",
         "   ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES with bench.py's share_of_step)"]
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{name[:72]:72s} n={n:4d} total={ms:9.3f} ms avg={ms / n * 1e3:9.1f} us share={ms / total * 100:5.1f}%")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
idx = {k: i for i, k in enumerate(h)}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__cycles_elapsed.avg.per_second", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
lines += ["", f"== ncu --set full --clock-control none --import-source on, same command (source: {os.path.basename(rep)}, "
              f"{os.path.getsize(rep) >> 20} MB, not committed)"]
seen, traffic = set(), {}
for r in rr[2:]:
    kname = r[idx["Kernel Name"]].split("(")[0]
    if kname in seen:
        continue
    seen.add(kname)
    lines.append(f"\n-- {kname}  grid={r[idx['launch__grid_size']]}")
    for w in WANT:
        if w in idx:
            lines.append(f"   {w} [{units[idx[w]]}] = {r[idx[w]]}")
    def to_bytes(metric):
        v, u = float(r[idx[metric]].replace(",", "")), units[idx[metric]].lower()
        return int(v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u])
    key = "zs_pool_fp32a" if "mlp_tc3" in kname else "zs_pool" if "mlp_tc" in kname else "zs_features" if "features" in kname else kname
    traffic[key] = {"kernel": kname.replace("void ", "").replace("<unnamed>::", ""),
                    "dram_read_bytes": to_bytes("dram__bytes_read.sum"), "dram_write_bytes": to_bytes("dram__bytes_write.sum")}
os.makedirs(out_dir, exist_ok=True)
open(os.path.join(out_dir, f"{tag}_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
shutil.copy(launches, os.path.join(out_dir, f"{tag}_launches_c2.csv"))
tj = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch, from one `ncu --set full --clock-control none` capture of "
                  f"`{cmd}` (C2, bf16).  zs_pool = the first captured launch of the MLP kernel: with the fused kernel all "
                  "hypotheses of one scorer (110,000 or 100,000 x 1,000 points, no feature rows in HBM: poses in, pooled "
                  "vectors out, frame and clouds through L2); with --no-fuse one 32,768-hypothesis chunk of bf16 feature rows.  "
                  f"zs_pool_fp32a = the re-rank's 3-term kernel (88 / 80 hypotheses).  See {tag}_ncu_summary.txt.",
      "workload": "c2", "precision": "bf16"}
tj.update(traffic)
json.dump(tj, open(os.path.join(out_dir, f"{tag}_traffic.json"), "w"), indent=2)
print("\n".join(lines))
