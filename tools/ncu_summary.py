"""Prints the metrics DESIGN.md / profiles/ quote from an `ncu --set full` report (run where ncu is installed):

    python tools/ncu_summary.py gpurun_out/prof_k2_c2.ncu-rep "header line" > profiles/rNN_x_ncu_full_....txt
"""
import csv
import io
import re
import subprocess
import sys

WANT = [
    r"^gpu__time_duration\.sum$",
    r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$", r"^dram__bytes_read\.sum\.per_second$",
    r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^l1tex__m_xbar2l1tex_read_bytes\.sum$", r"^l1tex__m_xbar2l1tex_read_bytes\.sum\.per_second$",
    r"^lts__t_sector_hit_rate\.pct$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_sectors_srcunit_tex_op_read\.sum$", r"^lts__t_sectors_srcunit_ltcfabric\.sum$",
    r"^launch__(grid_size|block_size|cluster_size|registers_per_thread|shared_mem_per_block_dynamic|waves_per_multiprocessor)$",
    r"^sm__cycles_elapsed\.avg$", r"^sm__cycles_elapsed\.avg\.per_second$",
    r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(elapsed|active)$",
    r"^sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__inst_executed\.sum$",
    r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    if len(sys.argv) > 2:
        print(sys.argv[2])
        print()
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("kernel:", d.get("Kernel Name", "?")[:110], "| grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for h, u, v in zip(hdr, units, vals):
            if any(re.search(w, h) for w in WANT):
                print(f"{h} [{u}] = {v}")
        print()


if __name__ == "__main__":
    main()
