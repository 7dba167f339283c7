"""Summarise `ncu --set full` reports (one launch each) into a table: duration, DRAM bytes / throughput, pipe activity, occupancy.
usage: python scripts/ncu_summary.py out.txt a.ncu-rep b.ncu-rep ...   (runs `ncu -i ... --page raw --csv`; no GPU needed)"""
import csv, io, subprocess, sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%act"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_%act"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%")]


def to_bytes(v, u):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(u, 1)


def main():
    out = open(sys.argv[1], "w")
    for rep in sys.argv[2:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].split("(")[0]
            t_us = float(r[ix["gpu__time_duration.sum"]]) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}[units[ix["gpu__time_duration.sum"]]]
            rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
            wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            out.write(f"== {rep.split('/')[-1]}: {name}\n")
            out.write(f"   time {t_us:.1f} us   dram read {rd / 1e6:.1f} MB  write {wr / 1e6:.1f} MB  -> {(rd + wr) / t_us / 1e6:.2f} TB/s"
                      f" = {(rd + wr) / t_us / 1e6 / 6.5431:.2f} of the measured copy peak (6543 GB/s)\n")
            for k, label in KEYS[3:]:
                if k in ix:
                    out.write(f"   {label:12s} {r[ix[k]]} {units[ix[k]]}\n")
    out.close()
    print(open(sys.argv[1]).read())


if __name__ == "__main__":
    main()
