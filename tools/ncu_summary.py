"""Summarises a `ncu --set full` report: key throughput / occupancy metrics and the top stall locations (needs -lineinfo).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/name.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "launch__cluster_size"]
print("report:", rep)
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print("%-70s %s %s" % (h, vals[i], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr_rows = [i for i, r in enumerate(rows) if "# Samples" in r]
if hdr_rows:
    h2 = rows[hdr_rows[0]]
    data = [r for r in rows[hdr_rows[0] + 1:] if len(r) == len(h2)]
    iS, iSrc = h2.index("# Samples"), h2.index("Source")
    stall = [i for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
    # (reports with several source views list every instruction once per view: keep the first occurrence of an address)
    iA = h2.index("Address") if "Address" in h2 else None
    if iA is not None:
        seen, uniq = set(), []
        for r in data:
            if r[iA] not in seen:
                seen.add(r[iA])
                uniq.append(r)
        data = uniq
    tot = sum(int(r[iS]) for r in data if r[iS].isdigit())
    agg = sorted(((sum(int(r[i]) for r in data if r[i].isdigit()), h2[i]) for i in stall), reverse=True)
    print("\nwarp-state samples: %d; by reason: %s" % (tot, ", ".join("%s %d" % (n, c) for c, n in agg[:8])))
    print("top stall locations (samples, SASS, dominant reason):")
    for s, i in sorted(((int(r[iS]), i) for i, r in enumerate(data) if r[iS].isdigit()), reverse=True)[:12]:
        st = sorted(((int(data[i][j]), h2[j]) for j in stall if data[i][j].isdigit()), reverse=True)[0]
        print("  %5d  %-70s %s" % (s, data[i][iSrc].strip()[:70], st[1]))
