"""Summarise an `ncu --set full` report (one row per captured launch) as markdown + per-launch DRAM traffic JSON.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_ncu_full.md [profiles/traffic.json]"""
import csv, io, json, subprocess, sys
rep, out_md = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum"]
ik = hdr.index("Kernel Name")
md = [f"# ncu --set full summary of `{rep.split('/')[-1]}`\n",
      "Times under ncu are serialised / cold-cache; use them for shares and per-launch traffic, not as bench values.\n"]
traffic = {}
for r in data:
    name = r[ik]
    md += [f"## {name[:100]}\n", "| metric | value | unit |", "|---|---|---|"]
    for w in want:
        if w in hdr:
            i = hdr.index(w); md.append(f"| {w} | {r[i]} | {units[i]} |")
    md.append("")
    def val(k):
        i = hdr.index(k); v = float(r[i].replace(",", "")); u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    key = "attention" if "attn" in name else "gemm1_gelu" if "<0," in name else "gemm2_combine" if "<1," in name else name[:30]
    traffic[key] = {"dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                    "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum")}
open(out_md, "w").write("\n".join(md))
if len(sys.argv) > 3:
    json.dump({"source": rep.split("/")[-1], "per_launch": traffic}, open(sys.argv[3], "w"), indent=1)
print(json.dumps(traffic, indent=1))
