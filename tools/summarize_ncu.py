import csv, collections, json, subprocess, sys, os
O="gpurun_out"; P="profiles"
def launch_list(path):
    rows=[r for r in csv.reader(open(path)) if len(r)>10]
    hdr=rows[0]; ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); iid=hdr.index("ID")
    d=collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault((int(r[iid]), r[ik].split("(")[0].split("::")[-1]),{})[r[im]]=float(r[iv].replace(",",""))
    return d
def table(d):
    agg=collections.OrderedDict()
    for (i,k),m in d.items():
        a=agg.setdefault(k,{"n":0,"ns":0.0,"rd":0.0,"wr":0.0,"fp64":0.0})
        a["n"]+=1; a["ns"]+=m.get("gpu__time_duration.sum",0); a["rd"]+=m.get("dram__bytes_read.sum",0); a["wr"]+=m.get("dram__bytes_write.sum",0)
        a["fp64"]+=max(m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",0), m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",0))
    tot=sum(a["ns"] for a in agg.values())
    out=["| kernel | launches | avg ms | share of GPU time | DRAM read+write per launch (GB) | fp64 pipe active % (DFMA or DMMA, whichever is higher) |","|---|---|---|---|---|---|"]
    for k,a in sorted(agg.items(), key=lambda t:-t[1]["ns"]):
        out.append(f"| {k} | {a['n']} | {a['ns']/a['n']/1e6:.3f} | {100*a['ns']/tot:.1f} % | {(a['rd']+a['wr'])/a['n']/1e9:.2f} | {a['fp64']/a['n']:.1f} |")
    return "\n".join(out), agg
def full(path, keys):
    txt=subprocess.run(["ncu","-i",path,"--page","raw","--csv"],capture_output=True,text=True).stdout
    rows=list(csv.reader(txt.splitlines())); hdr=rows[0]; units=rows[1]; r=rows[2]
    return {h:(r[i],units[i]) for i,h in enumerate(hdr) if h in keys or ("issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h)}
KEYS=["Kernel Name","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","launch__registers_per_thread","launch__grid_size","launch__block_size","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","sm__throughput.avg.pct_of_peak_sustained_elapsed"]
md=["# Round 2 ncu summaries (B200, 30 qubits = 16 GiB state, `--clock-control none`)","",
    "Launch lists: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active... -c 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-kernel-sweep [--workload qft]` (warm-up is forced to 3 steps; per-launch times under ncu are cold-cache and serialised — the SHARE per kernel is what carries over to the timed run).",""]
traffic={}
for name,f in (("random circuit (BASELINE config 2)","r02_launches_random.csv"),("QFT","r02_launches_qft.csv")):
    t,agg=table(launch_list(os.path.join(O,f)))
    md += [f"## Launch list — {name}","",t,""]
    if "random" in f and "k_tile_pipe" in agg: a=agg["k_tile_pipe"]; traffic["random:30"]={"kernel":"k_tile_pipe","dram_bytes_per_launch":int((a["rd"]+a["wr"])/a["n"]),"algorithmic_bytes_per_launch":32<<30}
    if "qft" in f and "k_qft_pipe" in agg: a=agg["k_qft_pipe"]; traffic["qft:30"]={"kernel":"k_qft_pipe","dram_bytes_per_launch":int((a["rd"]+a["wr"])/a["n"]),"algorithmic_bytes_per_launch":32<<30}
json.dump(traffic, open(os.path.join(P,"traffic.json"),"w"), indent=1)
for title,f in (("k_tile_pipe (fused gate block: TMA ring + DMMA rounds), one launch","r02_k_tile_pipe_30q.ncu-rep"),("k_qft_pipe (TMA-staged radix-8/4/2 QFT pass), one launch","r02_k_qft_pipe_30q.ncu-rep"),("k_bit_reverse (one-pass qubit reversal), one launch","r02_k_bit_reverse_30q.ncu-rep")):
    m=full(os.path.join(O,f),KEYS)
    md += [f"## `ncu --set full` — {title}","","| metric | value | unit |","|---|---|---|"]
    for k,(v,u) in m.items():
        try:
            fv=float(v)
            if "issue_stalled" in k and fv<0.2: continue
            v=f"{fv:.4g}"
        except ValueError: pass
        md.append(f"| {k} | {v} | {u} |")
    md.append("")
open(os.path.join(P,"r02_ncu_summary.md"),"w").write("\n".join(md))
print("\n".join(md)[:6000])
