"""Compare the sampled amplitude ranges tools/grover_sharded.py saved for different shard counts: the same circuit on
2, 4 and 8 GPUs must give the same amplitudes to 1e-12 (shard-count invariance)."""
import glob, re, sys
import numpy as np
by = {}
for f in glob.glob("gpurun_out/grover*_w*_r*.npy"):
    m = re.match(r".*/(grover\d+_\w+_k\d+)_w(\d+)_r(\d+)\.npy", f)
    by.setdefault((m.group(1), int(m.group(3))), {})[int(m.group(2))] = np.load(f)
worst, pairs = 0.0, 0
for (name, j), d in sorted(by.items()):
    ws = sorted(d)
    for w in ws[1:]:
        worst = max(worst, float(np.max(np.abs(d[w] - d[ws[0]]))))
        pairs += 1
print(f"GROVER_COMPARE ranges={len(by)} pairs={pairs} worlds={sorted({w for d in by.values() for w in d})} max|d|={worst:.3e}")
sys.exit(0 if worst <= 1e-12 and pairs > 0 else 1)
