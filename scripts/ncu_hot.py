"""Print the hottest SASS lines (warp-stall samples) of one kernel in an .ncu-rep.
   python scripts/ncu_hot.py file.ncu-rep <kernel-regex> [launch-index] [top-n]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
idx = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
if rx != "-":
    cmd += ["--kernel-id", f"::regex:{rx}:{idx}"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si = hdr.index("Warp Stall Sampling (All Samples)")
src = hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
items = []
for k, r in enumerate(rows[2:]):
    try:
        v = float(r[si])
    except Exception:
        continue
    st = sorted(((float(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    items.append((v, k, r[src].strip()[:90], st))
tot = sum(v for v, *_ in items)
print("kernel:", rows[0][1][:100], " total samples", tot)
for v, k, s, st in sorted(items, reverse=True)[:top]:
    print(f"{v:8.0f} {100*v/max(tot,1):5.1f}%  #{k:<5d} {s:<90s} {st[0][1]}={st[0][0]:.0f} {st[1][1]}={st[1][0]:.0f}")
