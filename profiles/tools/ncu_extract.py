"""Extract the metrics quoted in profiles/*.md from an .ncu-rep (uses `ncu -i ... --page raw --csv`)."""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum"]

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
res = []
for r in rows[2:]:
    name = r[ki].split("(")[0].split("::")[-1]
    d = {"kernel": name}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            d[w] = "%s %s" % (r[i], units[i])
    res.append(d)
if len(sys.argv) > 2 and sys.argv[2] == "--json":
    print(json.dumps(res, indent=1))
else:
    for d in res:
        print("\n**%s**\n\n| metric | value |\n|---|---|" % d["kernel"])
        for w in WANT:
            if w in d:
                print("| %s | %s |" % (w, d[w]))
