"""Developer tool: digest of `ncu --page raw --csv` exports (one row per profiled launch): duration, tensor-pipe / SM / DRAM / L2
utilisation, DRAM bytes, occupancy and registers -- the numbers DESIGN.md and profiles/*_summary.md quote.  Usage:
    python tools/ncu_digest.py gpurun_out/r2x_*_ncu_raw.csv > profiles/r2x_ncu_digest.txt"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "duration"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (of elapsed)"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy % (active SMs)"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe % (active SMs)"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts by tensor core %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "registers/thread"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem/CTA")]


def main(paths):
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"== {path}")
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            print("  kernel:", r[col["Kernel Name"]][:110])
            for key, label in WANT:
                if key in col:
                    print(f"    {label:42s} {r[col[key]]} {units[col[key]]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1:])
