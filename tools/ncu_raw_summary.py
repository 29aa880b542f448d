"""Short text summary of an .ncu-rep (one kernel per report): the raw-page metrics the profiles/ README quotes.

    python tools/ncu_raw_summary.py gpurun_out/sor_r02r.ncu-rep > profiles/sor_r02r_raw.txt
"""
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "dram__throughput", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate", "sm__throughput", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active", "smsp__average_warp", "smsp__average_warps_issue_stalled", "launch__occupancy_limit", "sm__maximum_warps", "launch__waves",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__cycles_active.avg", "sm__inst_executed_pipe_fp64", "smsp__issue_active", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__throughput", "l1tex__throughput")


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        for h, u, v in zip(hdr, units, vals):
            if h == "Kernel Name" or any(h.startswith(k) for k in KEEP):
                if "not_issued" in h:
                    continue
                print(f"{h} [{u}] = {v}")
        print()


if __name__ == "__main__":
    main()
