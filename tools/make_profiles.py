"""Copies the judged ncu evidence into profiles/: the launch list of the bench command, the raw metric page of the
full capture of aw_align_kernel, a per-region source summary, and profiles/ncu_summary.json (read by bench.py for
roofline.traffic).  usage: python tools/make_profiles.py <tag> <launches.csv> <full.ncu-rep> "<bench command>" <config> <pairs per launch>
The summary is stamped with a hash of the kernel sources; bench.py reports roofline.traffic only when the stamp matches the
sources it runs (a capture of an older kernel says nothing about the current one)."""
import csv, hashlib, io, json, os, shutil, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep, cmd = sys.argv[1:5]
config = sys.argv[5] if len(sys.argv) > 5 else "C2"
pairs_per_launch = int(sys.argv[6]) if len(sys.argv) > 6 else 9472
_h = hashlib.sha1()
for _f in ("aw_wfa.cuh", "aw_common.cuh", "aw_sketch.cuh"):
    _h.update(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "allwave_b200", "csrc", _f), "rb").read())
kernel_source_hash = _h.hexdigest()[:16]
P = os.path.join(ROOT, "profiles")
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
with open(os.path.join(P, f"{tag}_launches_bench.csv"), "w", newline="") as f:
    csv.writer(f).writerows(rows)
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    tot[r[ik].split("(")[0]] += v
    cnt[r[ik].split("(")[0]] += 1
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_align_kernel_raw.csv"), "w").write(raw)
r3 = list(csv.reader(io.StringIO(raw)))
m = dict(zip(r3[0], r3[2]))
u = dict(zip(r3[0], r3[1]))
def gb(name):
    v = float(m[name]); unit = u[name]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}[unit]
reg = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), rep, "25"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_align_kernel_regions.txt"), "w").write(reg)
s = sum(tot.values())
summary = {
    "tag": tag, "command": cmd, "config": config, "kernel_source_hash": kernel_source_hash,
    "launch_shares": {k: {"share": v / s, "launches": cnt[k], "total_ms": v / 1e6} for k, v in tot.most_common()},
    "align_kernel": {
        "pairs_per_launch": pairs_per_launch,
        "name": m.get("Kernel Name", "aw_align_kernel"), "grid": m.get("launch__grid_size"), "block": m.get("launch__block_size"), "regs": m.get("launch__registers_per_thread"),
        "duration_ms": float(m["gpu__time_duration.sum"]) * {"s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(u["gpu__time_duration.sum"], 1e-6),
        "dram_bytes_read": gb("dram__bytes_read.sum"), "dram_bytes_write": gb("dram__bytes_write.sum"),
        "dram_bytes_per_launch": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"),
        "dram_throughput_pct": float(m["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]),
        "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        "warps_active_pct": float(m["sm__warps_active.avg.pct_of_peak_sustained_active"]),
        "inst_executed": float(m["smsp__inst_executed.sum"]),
        "l1_hit_pct": float(m["l1tex__t_sector_hit_rate.pct"]), "l2_hit_pct": float(m["lts__t_sector_hit_rate.pct"]),
        "stalls_per_issue": {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v)
                             for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and float(v or 0) > 0.05},
    },
}
json.dump(summary, open(os.path.join(P, "ncu_summary.json"), "w"), indent=1)
print(json.dumps(summary["align_kernel"], indent=1)[:1500])
