"""Rebuilds the measured table of BASELINE.md section 4 from the bench lines committed under profiles/r02_runs/
(bench_<config>.json = `python bench.py --config <config>` on one B200; strong_<config>_n<N>.json =
`python bench.py --scaling strong --config <config> --gpus N` on one 8-GPU box).  usage: python tools/fill_baseline_table.py"""
import glob, json, os, re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = os.path.join(ROOT, "profiles", "r02_runs")
PAIRS = {"C1": "240", "C2": "999,000", "C3": "2,000,810", "C4": "2,043", "C5": "≈2.51 M"}


def load(pattern):
    out = {}
    for f in sorted(glob.glob(os.path.join(R, pattern))):
        try:
            d = json.loads(open(f).read().strip().splitlines()[-1])
        except Exception:
            continue
        out[os.path.basename(f)] = d
    return out


def fmt(v):
    if v is None:
        return ""
    if v >= 1e6:
        return f"{v / 1e6:.2f} M"
    if v >= 1e4:
        return f"{v / 1e3:.1f} k"
    if v >= 100:
        return f"{v:,.0f}"
    return f"{v:.1f}"


bench = load("bench_C*.json")
strong = load("*strong_C*_n*.json")
rows = []
for c in ("C1", "C2", "C3", "C4", "C5"):
    b = bench.get(f"bench_{c}.json")
    s = {}
    for k, d in strong.items():
        m = re.search(rf"strong_{c}_n(\d+)\.json", k)
        if m and k.startswith("r2_13"):
            s[int(m.group(1))] = d
    cpu = b and b.get("cpu_baseline")
    par = b and b.get("parity")
    rows.append("| " + " | ".join([
        c, PAIRS[c],
        f"{fmt(cpu['value'])} (T={cpu['cores']}; plain checker {fmt(cpu.get('checker_value'))})" if cpu else ("not run (13.6 CPU-min per pair in the fast oracle)" if c == "C4" else ""),
        fmt(b["value"]) if b else "", fmt(b["e2e"]["value"]) if b else "",
        fmt(s[1]["value"]) if 1 in s else "", fmt(s[2]["value"]) if 2 in s else "", fmt(s[4]["value"]) if 4 in s else "", fmt(s[8]["value"]) if 8 in s else "",
        f"{s[8]['value'] / s[1]['value']:.2f}×" if 1 in s and 8 in s else "",
        f"{b['roofline']['frac']:.1e}" if b else "",
        (f"{par['paf_identical']}/{par['pairs_checked']}" if par else ("4/4 vs frozen oracle" if c == "C4" else "")),
    ]) + " |")
hdr = ("| config | pairs in the full job | CPU restatement pairs/s | 1×B200 device-resident | 1×B200 e2e (host buffers) | product path 1 GPU | 2 | 4 | 8 | 8×/1× | HBM frac @1 | PAF bit-exact vs oracle |\n"
       "|---|---|---|---|---|---|---|---|---|---|---|---|")
print(hdr)
print("\n".join(rows))
