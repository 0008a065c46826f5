#!/usr/bin/env python
"""Summarise an ncu --set full capture of the control-cycle kernel: headline metrics, stall reasons, SASS opcode mix.
Usage: tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [out.md]"""
import collections, csv, io, json, re, subprocess, sys

rep = sys.argv[1]
out_md = sys.argv[2] if len(sys.argv) > 2 else None

def ncu(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout

raw = list(csv.reader(io.StringIO(ncu("raw"))))
hdr, units = raw[0], raw[1]
rows = [r for r in raw[2:] if "control_cycle" in r[hdr.index("Kernel Name")]]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "launch__grid_size", "launch__block_size", "smsp__warps_eligible.avg.per_cycle_active"]
stall = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
if not stall:
    stall = [h for h in hdr if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct")]
summary = {"report": rep, "kernels": []}
lines = []
for r in rows:
    k = {"name": r[hdr.index("Kernel Name")]}
    for w in want:
        if w in hdr:
            try: k[w] = float(r[hdr.index(w)].replace(",", ""))
            except ValueError: k[w] = r[hdr.index(w)]
            k[w + ":unit"] = units[hdr.index(w)]
    st = {}
    for s_ in stall:
        try: st[s_] = float(r[hdr.index(s_)].replace(",", ""))
        except ValueError: pass
    k["stalls"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:8])
    summary["kernels"].append(k)

# opcode mix of the first captured launch
src = list(csv.reader(io.StringIO(ncu("source"))))
secs = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
mix, total = collections.Counter(), 0
if secs:
    start = secs[0]; end = secs[1] if len(secs) > 1 else len(src)
    h2 = src[start + 1]
    iS, iE = h2.index("Source"), h2.index("Instructions Executed")
    for r in src[start + 2:end]:
        if len(r) <= iE: continue
        m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS].strip())
        if not m: continue
        try: e = int(r[iE])
        except ValueError: continue
        mix[m.group(1).split(".")[0]] += e; total += e
summary["opcode_mix_warp_instructions"] = dict(mix.most_common(25))
summary["warp_instructions_total"] = total
k0 = summary["kernels"][0] if summary["kernels"] else {}
grid = k0.get("launch__grid_size", 0); block = k0.get("launch__block_size", 0)
warps = grid * block / 32 if grid and block else 0
summary["warp_instructions_per_warp"] = total / warps if warps else None
txt = json.dumps(summary, indent=1)
print(txt)
if out_md:
    with open(out_md, "w") as f:
        f.write(txt + "\n")
