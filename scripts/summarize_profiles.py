"""Turn the ncu artefacts in gpurun_out/ into the tracked summaries under profiles/.
    python scripts/summarize_profiles.py <round-tag>        e.g. r01
Inputs (made by scripts/gpu_profile.sh on the B200 box):
    gpurun_out/launches.csv         ncu --metrics gpu__time_duration.sum launch list of 3 C2 steps
    gpurun_out/prof_decoder.ncu-rep / prof_gru.ncu-rep / prof_gemm.ncu-rep   ncu --set full captures
"""
import csv, os, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
lines = [f"# {tag}: ncu summaries (B200, C2 inference step: B=32, Tx=128, T=200, r=5, precision fp32x3 = 3xTF32 on tcgen05)\n"]

# ---------------- launch list ----------------
lp = os.path.join(GO, f"{tag}_launches.csv")
if not os.path.exists(lp):
    lp = os.path.join(GO, "launches.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    gi_ = [i for i, r in enumerate(data) if "gather_rows" in r[ki]]
    start = gi_[-1] - 1 if gi_ else 2 * (len(data) // 3)      # a step starts with the embedding-table GEMM + the row gather
    step = data[start:]
    tot = sum(float(r[vi].replace(",", "")) for r in step)
    with open(os.path.join(OUT, f"{tag}_launches_step.csv"), "w") as f:
        f.write("kernel,grid,duration_us,share\n")
        for r in step:
            v = float(r[vi].replace(",", ""))
            f.write(f"\"{r[ki][:70]}\",\"{r[gi]}\",{v/1000:.1f},{v/tot:.4f}\n")
    lines.append(f"## Launch list of one steady-state step ({len(step)} launches, ncu serialised/cold-cache: compare SHARES)\n")
    lines.append("| kernel | grid | us | share |\n|---|---|---:|---:|")
    agg = {}
    for r in step:
        nm = r[ki].replace("<unnamed>::", "").replace("void ", "").split("(")[0]
        v = float(r[vi].replace(",", ""))
        a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += v
    for nm, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"| {nm} x{n} | | {v/1000:.1f} | {100*v/tot:.1f}% |")
    lines.append(f"| total | | {tot/1000:.1f} | 100% |\n")

# ---------------- full captures ----------------
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def stalls(rep, idx):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{idx}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return {}
    hdr = rows[1]
    cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = {hdr[i]: 0.0 for i in cols}
    for r in rows[2:]:
        for i in cols:
            try:
                tot[hdr[i]] += float(r[i] or 0)
            except ValueError:
                pass
    s = sum(tot.values()) or 1
    return {k: v / s for k, v in sorted(tot.items(), key=lambda x: -x[1])[:6]}


for name in ("prof_decoder", "prof_gru", "prof_gemm", "prof_dw"):
    rep = os.path.join(GO, f"{tag}_{name}.ncu-rep")
    if not os.path.exists(rep):
        rep = os.path.join(GO, name + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, data = raw(rep)
    lines.append(f"## {name}.ncu-rep (ncu --set full --clock-control none)\n")
    ki = hdr.index("Kernel Name")
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    lines.append("| # | kernel | " + " | ".join(w.split(".")[0].replace("launch__", "").replace("sm__", "") for w, _ in cols) + " |")
    lines.append("|---|---|" + "---:|" * len(cols))
    for n, r in enumerate(data):
        nm = r[ki].replace("<unnamed>::", "").replace("void ", "").split("(")[0][:40]
        vals = []
        for w, i in cols:
            vals.append(f"{r[i]} {units[i]}".strip())
        lines.append(f"| {n} | {nm} | " + " | ".join(vals) + " |")
    lines.append("")
    for n in range(min(len(data), 3 if name != "prof_gemm" else 0)):
        st = stalls(rep, n)
        if st:
            lines.append(f"warp-stall sampling, launch {n}: " + ", ".join(f"{k.replace('stall_', '')} {100*v:.0f}%" for k, v in st.items()) + "\n")

open(os.path.join(OUT, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines)[:6000])
