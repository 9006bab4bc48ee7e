"""Condense an `ncu -i <rep> --page raw --csv` export into the per-launch summary tables kept under profiles/.

usage: python tools/ncu_summary.py raw.csv summary.csv [traffic.json]
The optional third argument writes the mean dram__bytes_read+write per launch (MB -> bytes) of the launches in the
export, in the format bench.py reads for `roofline.traffic`.
"""
import csv, json, sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_elapsed.avg.per_second"]


def main():
    raw, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, body = rows[hdr_i], rows[hdr_i + 1], rows[hdr_i + 2:]
    idx = {c: hdr.index(c) for c in COLS if c in hdr}
    kn, gs, bs = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "grid", "block"] + [f"{c} [{units[idx[c]]}]" for c in idx])
        tot, n = 0.0, 0
        for i, r in enumerate(body):
            if len(r) <= kn:
                continue
            w.writerow([i, r[kn].split("(")[0].replace("hvit::<unnamed>::", ""), r[gs], r[bs]] + [r[idx[c]] for c in idx])
            if "dram__bytes_read.sum" in idx:
                def to_bytes(c):
                    v = float(r[idx[c]].replace(",", ""))
                    u = units[idx[c]].lower()
                    return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0)
                tot += to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
                n += 1
    if len(sys.argv) > 3 and n:
        json.dump({"kernel_family": "igemm_tc (igemm_tc2_kernel / igemm_halo_kernel)", "source": out,
                   "dram_bytes_per_launch": tot / n, "launches": n}, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
