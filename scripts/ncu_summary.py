"""Summarise .ncu-rep files (ncu -i ... --page raw --csv) into a small table for profiles/."""
import csv, subprocess, sys
WANT = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    print(f"## {rep}")
    print("| " + " | ".join(f"{w} [{units[i]}]" for w, i in idx) + " |")
    print("|" + "---|" * len(idx))
    for r in rows[2:]:
        print("| " + " | ".join(r[i][:70] for _, i in idx) + " |")
    print()
