"""Aggregates an ncu launch list (--csv --log-file, any of gpu__time_duration.sum / dram__bytes_read.sum / dram__bytes_write.sum)
per kernel name: launches, total time, share, DRAM bytes.  Usage: python tools/launch_list_summary.py list.csv[.gz] [top]"""
import collections, csv, gzip, re, sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
op = gzip.open if path.endswith(".gz") else open
with op(path, "rt", errors="replace") as f:
    lines = [l for l in f if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
ki, mi, ui, vi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:110]


agg = collections.defaultdict(lambda: {"n": set(), "ms": 0.0, "rd": 0.0, "wr": 0.0})
for r in rows[1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1.0)
    except ValueError:
        continue
    a = agg[short(r[ki])]
    a["n"].add(r[ii])
    if r[mi] == "gpu__time_duration.sum":
        a["ms"] += v
    elif r[mi] == "dram__bytes_read.sum":
        a["rd"] += v
    elif r[mi] == "dram__bytes_write.sum":
        a["wr"] += v
tot = sum(a["ms"] for a in agg.values())
rd = sum(a["rd"] for a in agg.values())
wr = sum(a["wr"] for a in agg.values())
print(f"source: {path}; {sum(len(a['n']) for a in agg.values())} launches, {tot:.2f} ms under ncu (cold-cache, serialised: shares, not absolutes), "
      f"DRAM read {rd / 1e9:.2f} GB, write {wr / 1e9:.2f} GB\n")
print("| ms | share | launches | DRAM read GB | DRAM write GB | kernel |\n|--|--|--|--|--|--|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:top]:
    print(f"| {a['ms']:.3f} | {100 * a['ms'] / max(tot, 1e-9):.1f}% | {len(a['n'])} | {a['rd'] / 1e9:.3f} | {a['wr'] / 1e9:.3f} | `{k}` |")
