"""Summarise an .ncu-rep: per-kernel totals, key metrics of the largest launch, top stalls; optional per-instruction stall table.
usage: python tools/ncu_summary.py rep.ncu-rep [kernel-regex-for-source-page]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
ki = h.index("Kernel Name")
tv = lambda r: float(r[h.index("gpu__time_duration.sum")].replace(",", ""))
best, tot, cnt = {}, {}, {}
for r in rows[2:]:
    k = r[ki].split("(")[0]
    tot[k] = tot.get(k, 0) + tv(r); cnt[k] = cnt.get(k, 0) + 1
    if k not in best or tv(r) > tv(best[k]): best[k] = r
exact = ["launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
         "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
         "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
         "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active"]
for k, r in sorted(best.items(), key=lambda kv: -tot[kv[0]]):
    if tot[k] < 0.05: continue
    print(f"===== {k[:60]}  launches={cnt[k]} total={tot[k]:.3f} ms  largest={tv(r):.3f} ms")
    for n in exact:
        if n in h: print(f"  {n:72s} {r[h.index(n)]} {u[h.index(n)]}")
    st = [(float(r[i].replace(",", "")), h[i]) for i in range(len(h)) if "smsp__average_warps_issue_stalled" in h[i] and "per_issue_active" in h[i] and r[i] not in ("", "n/a")]
    print("  stalls: " + ", ".join(f"{n[34:].split('_per_')[0]} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # sections: ["Kernel Name", name] / header row / data rows
    sections = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    for si, start in enumerate(sections):
        end = sections[si + 1] if si + 1 < len(sections) else len(rows)
        name = rows[start][1][:70]
        h = rows[start + 1]; ci = {n: i for i, n in enumerate(h)}
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        total = {s_: 0 for s_ in stalls}; recs = []; samples = 0
        for r in rows[start + 2:end]:
            if len(r) < len(h): continue
            n = int(r[ci["# Samples"]]); samples += n
            d = {s_: int(r[ci[s_]]) for s_ in stalls if int(r[ci[s_]])}
            for s_, v in d.items(): total[s_] += v
            recs.append((n, r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]]), d))
        if not samples: continue
        print("#####", name, "samples", samples, {k[6:]: f"{100*v/samples:.1f}%" for k, v in sorted(total.items(), key=lambda kv: -kv[1]) if v})
        for n, s_, ie, d in sorted(recs, key=lambda x: -x[0])[:top]:
            print(f"{n:7d} {ie:10d} {s_[:56]:56s} " + " ".join(f"{k[6:]}={v}" for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:3]))
