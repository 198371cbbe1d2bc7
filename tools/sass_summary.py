"""Static SASS summary of the shipped library (profiles/r02_sass_summary.md): per-kernel counts of the instructions the
design rests on, from `cuobjdump -sass`, plus registers / spills from `cuobjdump -res-usage`."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qcsim_b200", "libqcsim_b200.so")
COLS = [("UTMALDG", r"UTMALDG"), ("UTMASTG", r"UTMASTG"), ("SYNCS", r"SYNCS\."), ("DMMA", r"DMMA"), ("DFMA", r"DFMA"), ("DMUL", r"DMUL"), ("DADD", r"DADD"),
        ("LDS.128", r"LDS\.128"), ("LDS.64", r"LDS\.64"), ("STS.128", r"STS\.128"), ("STS.64", r"STS\.64"), ("LDG.E.ENL2.256", r"LDG\.E\.ENL2\.256"),
        ("STG.E.ENL2.256", r"STG\.E\.ENL2\.256"), ("LDL", r"\bLDL"), ("STL", r"\bSTL"), ("BAR", r"\bBAR\."), ("FENCE", r"FENCE\.VIEW\.ASYNC")]


def demangle(name):
    m = re.search(r"(k_[a-z0-9_]+?)(?:E|I|ILi|IL)", name)
    return m.group(1) if m else name


sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, copies, cur = collections.OrderedDict(), collections.Counter(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1))
        counts.setdefault(cur, collections.Counter())
        copies[cur] += 1
        continue
    if cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
        for col, pat in COLS:
            if re.search(pat, line):
                counts[cur][col] += 1
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
regs = {}
name = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = demangle(m.group(1))
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and name:
        r = regs.setdefault(name, [0, 0])
        r[0] = max(r[0], int(m.group(1)))
        r[1] = max(r[1], int(m.group(2)))

out = ["# Static SASS summary of libqcsim_b200.so, round 2 (sm_100a, `cuobjdump -sass` / `-res-usage`; `python tools/sass_summary.py`)", "",
       "Counts of the instructions that matter for the design, per kernel of the shipped library, per copy (`copies` = template instances / translation units that hold the kernel; their counts are averaged).  `UTMALDG` / `UTMASTG` ="
       " `cp.async.bulk.tensor` (TMA) loads / stores, `SYNCS.*` = mbarrier operations (`ARRIVE.TRANS64`, `PHASECHK.TRANS64.TRYWAIT`, `EXCH.64`), `DMMA` ="
       " `mma.sync.m8n8k4.f64` (`DMMA.8x8x4`), `LDG/STG.E.ENL2.256` = two amplitudes per 256-bit access, `FENCE` = `fence.proxy.async` (`FENCE.VIEW.ASYNC.S`),"
       " `LDL` / `STL` = local memory (spills).", "",
       "| kernel | copies | regs | stack | " + " | ".join(c for c, _ in COLS) + " |", "|---|---|---|---|" + "---|" * len(COLS)]
for k, c in counts.items():
    r = regs.get(k, ["?", "?"])
    out.append(f"| `{k}` | {copies[k]} | {r[0]} | {r[1]} | " + " | ".join(str(c[col] // copies[k]) for col, _ in COLS) + " |")
tp = {k: v // max(1, copies["k_tile_pipe"]) for k, v in counts.get("k_tile_pipe", {}).items()}
out += ["", f"`k_tile_pipe`: {tp.get('DMMA', 0)} `DMMA.8x8x4` per round body (6 per panel of 8 items x 4 panels), 16 `LDS.64` + 16 `STS.64` fragment halves, "
        f"{tp.get('UTMALDG', 0)} / {tp.get('UTMASTG', 0)} TMA load / store sites in the producer warp, no local memory."]
open(os.path.join(ROOT, "profiles", "r02_sass_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-8:]))
