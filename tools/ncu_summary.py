"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-launch table kept under profiles/.

    ncu -i gpurun_out/full.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv > profiles/rN_ncu_full_summary.csv
"""
import csv, sys

KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum.per_cycle_elapsed", "sm__cycles_elapsed.max",
]
import json
FM_KERNELS = ("fm::", "seg::fwd_kernel", "dkdv_kernel", "dq_kernel", "qkv_kernel", "group_sum_kernel", "apply_kernel", "build_kernel", "intra_kernel", "prep_bwd_kernel", "transpose_kernel", "scan_kernel",
              "reduce_kernel", "applyfull_kernel", "finalize_kernel")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [c for c in KEEP if c in ix]
w = csv.writer(sys.stdout)
w.writerow(cols)
for r in rows[2:]:
    if len(r) < len(hdr) or not any(k in r[ix["Kernel Name"]] for k in FM_KERNELS):      # only this library's kernels
        continue
    w.writerow([r[ix[c]] + ("" if c == "Kernel Name" or not units[ix[c]] else " " + units[ix[c]]) for c in cols])

# optional: total DRAM traffic of the captured launches -> JSON next to the table (bench.py's roofline.traffic)
if len(sys.argv) > 2:
    def gb(r, c):
        v, u = float(r[ix[c]]), units[ix[c]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
    fm_rows = [r for r in rows[2:] if len(r) >= len(hdr) and any(k in r[ix["Kernel Name"]] for k in
               FM_KERNELS)]
    tot = sum(gb(r, "dram__bytes_read.sum") + gb(r, "dram__bytes_write.sum") for r in fm_rows)
    json.dump({"launches": len(fm_rows), "dram_bytes_per_step": tot, "config": {"B": 8, "H": 32, "N": 16384, "kept_states": True},
               "source": "ncu --set full, one fwd+bwd step at the headline shape (tools/prof_run.py 8 32 16384 64 1 --bwd --keep)"},
              open(sys.argv[2], "w"), indent=1)
