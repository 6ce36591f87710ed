"""Count the Blackwell-specific SASS mnemonics per kernel of the built library (evidence for tcgen05 / TMEM / TMA use).
    python scripts/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "tacotron_b200", "libtaco_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UBLKCP[.\w]*|UTCBAR[.\w]*|HMMA\.1688\.F32\.TF32|FFMA2|SYNCS[.\w]*|UTCATOMSWS[.\w]*|ST\.E\.64\.STRONG\.GPU|LDG\.E\.128\.STRONG\.SYS|STS?\.\w*CLUSTER\w*|MAPA\w*|UCGABAR\w*)")
cnt = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "")[:110]
        cnt[name] = collections.Counter()
        continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        cnt[name]["(instructions)"] += 1
        for mm in pat.findall(line):
            cnt[name][mm] += 1
print("# cuobjdump -sass tacotron_b200/libtaco_b200.so -- Blackwell-specific mnemonics per kernel")
print("# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (tensor memory), UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk,")
print("# SYNCS = mbarrier, HMMA.1688.F32.TF32 = mma.sync m16n8k8 tf32, FFMA2 = packed fp32x2 FMA, MAPA / ST..CLUSTER / UCGABAR = cluster DSMEM")
for k, c in cnt.items():
    items = [f"{m} x{n}" for m, n in c.items() if m != "(instructions)"]
    if items:
        print(f"{k}\n    {c['(instructions)']} instructions; " + ", ".join(items))
