"""Regenerates the results table of README.md (between the results markers) from the bench lines under profiles/."""
import json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def line(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return None
    with open(path) as f:
        for l in f:
            if l.startswith("{"):
                return json.loads(l)
    return None


def fmt(x, nd=1):
    return "—" if x is None else f"{x:,.{nd}f}"


rows = []
ref = line("r02_bench_reference.json")
full = line("r02_bench_random_30q.json")           # full evidence run (kernel sweeps, cpu_baseline)
one = line("r02_bench_random_30q_final.json") or full  # last measurement of the round (register-triple round planner), bench line only
r01 = line("r01_bench_random_30q.json")
if one:
    rf = one["roofline"]
    rows.append(("random-circuit layer, 30 q, 1 GPU (BASELINE config 2 width)", f"{one['ms_per_step']:.2f} ms / layer", f"{fmt(one['value'])} gate-apps/s",
                 f"round 1: {r01['ms_per_step']:.1f} ms; reference CPU arm {fmt(ref['value'], 2) if ref else '—'} gate-apps/s on {ref['cpu_baseline']['cores'] if ref else '?'} cores; "
                 f"{rf['passes_per_step']} HBM passes and {rf['rounds_per_step']} DMMA rounds per layer; `k_tile_pipe` at {rf['frac']:.2f} of the measured HBM peak, "
                 f"{2 * 24 * rf['rounds_per_step'] * 2 ** 30 / (one['ms_per_step'] * 1e-3) / 1e12:.1f} of {rf['fp64']['peak_tflops_fp64']} fp64 TFLOP/s executed (24 FMA per amplitude and round)"))
    rows.append(("same, end to end (host gate descriptors in, host result out, per layer)", f"{one['e2e']['ms_per_step']:.2f} ms / layer", f"{fmt(one['e2e']['value'])} gate-apps/s", "one GetQubitProbability per layer: only the queued gates that can change it are flushed, the rest keeps fusing across layers"))
    k = (full or one).get("kernels", {})
    if k:
        fr = [v["frac_of_peak"] for n, v in k.items() if "Measure" not in n and "Probability" not in n and "c=0" not in n and "SWAP" not in n]
        rows.append(("single-gate kernels, 30 q (H, RZ, CNOT, CPhase, CCX, dense 4x4 / 8x8)", f"{min(fr):.2f}–{max(fr):.2f} of HBM peak", "", "`kernels` / `kernels_33q` in the bench line; CNOT/SWAP with a qubit-0 partner 0.46 (16-byte granules)"))
        m = k.get("MeasureAll scan (no collapse)")
        if m:
            rows.append(("MeasureAll, 30 q, bit-exact reference outcome", f"{m['ms']:.1f} ms", "", "two passes + block-wide walk of the sequential sum"))
for q in (30, 33):
    d = line(f"r02_bench_qft_{q}q.json")
    o = line(f"r01_bench_qft_{q}q.json")
    if d:
        rows.append((f"QFT, {q} q, 1 GPU" + (" (BASELINE config 3)" if q == 33 else ""), f"{d['ms_per_step']:.1f} ms / transform", f"{fmt(d['value'])} gate-apps/s",
                     f"round 1: {o['ms_per_step']:.1f} ms; 4 TMA-staged passes + 1 reversal pass = {5 * 32 * 2 ** q / 6453.7e9 * 1e3:.1f} ms at the HBM peak"))
for n in (2, 4, 8):
    d = line(f"r02_bench{n}_random.json")
    if d and one:
        ex = d.get("exchange", {})
        rows.append((f"random-circuit layer, {d['config']['qubits']} q, {n} GPUs (2^30 amplitudes per GPU)", f"{d['ms_per_step']:.2f} ms / layer", f"{fmt(d['value'])} gate-apps/s",
                     f"weak-scaling efficiency {d['value'] / (n * (full or one)['value']):.2f} vs the 1-GPU run of the same build ({(full or one)['ms_per_step']:.1f} ms; measured before the last planner change); exchange {fmt(ex.get('GBps_per_direction'), 0)} GB/s per direction"))
    d = line(f"r02_bench{n}_qft30.json")
    if d:
        rows.append((f"QFT, {d['config']['qubits']} q, {n} GPUs", f"{d['ms_per_step']:.1f} ms / transform", f"{fmt(d['value'])} gate-apps/s", ""))
d = line("r02_bench8_qft36.json")
if d:
    ex = d["exchange"]
    rows.append(("QFT, 36 q, 8 GPUs (1 TiB, BASELINE config 5)", f"{d['ms_per_step']:.0f} ms / transform", f"{fmt(d['value'])} gate-apps/s",
                 f"exchange {fmt(ex['GBps_per_direction'], 0)} GB/s per direction = {ex['frac_of_nvlink_900']:.2f} of NVLink 5"))
d = line("r02_bench8_random36.json")
if d:
    rows.append(("random-circuit layer, 36 q, 8 GPUs (BASELINE config 5)", f"{d['ms_per_step']:.0f} ms / layer", f"{fmt(d['value'])} gate-apps/s", ""))
for n in (2, 4, 8):
    path = os.path.join(P, f"r02_grover32_{n}gpu.log")
    if os.path.exists(path):
        m = re.search(r"time=([\d.]+)s.*P\(marked\)=([\d.e+-]+).*diff=([\d.e+-]+)", open(path).read())
        if m:
            rows.append((f"Grover, 32 q, 201 iterations, {n} GPUs (BASELINE config 4)", f"{float(m.group(1)):.1f} s", "", f"P(marked) = {float(m.group(2)):.8f}, {m.group(3)} from the analytic value"))

table = "| Workload | Time | Throughput | Notes |\n|---|---|---|---|\n" + "\n".join("| " + " | ".join(r) + " |" for r in rows)
readme = os.path.join(ROOT, "README.md")
s = open(readme).read()
begin, end = "<!-- results:begin -->", "<!-- results:end -->"
if begin in s:
    s = s[: s.index(begin) + len(begin)] + "\n" + table + "\n" + s[s.index(end):]
else:
    s = s.replace("RESULTS_TABLE", begin + "\n" + table + "\n" + end)
open(readme, "w").write(s)
print(table)
