"""Summarise .ncu-rep captures (read here with `ncu -i`) into markdown tables for profiles/."""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration under ncu"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed (max)"),
    ("smsp__cycles_active.avg", "SMSP cycles active (avg)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (hmma subpipe) active % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % of peak"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM read bytes"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster size"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
]


def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    return hdr, units, r[2:]


def main():
    for rep in sys.argv[1:]:
        hdr, units, data = rows(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        for d in data:
            name = d[idx["Kernel Name"]]
            print(f"### `{rep.split('/')[-1]}` - kernel `{name[:110]}`\n")
            print("| metric | value |\n|---|---|")
            for k, label in KEYS:
                if k in idx and d[idx[k]] != "":
                    print(f"| {label} (`{k}`) | {d[idx[k]]} {units[idx[k]]} |")
            print()


if __name__ == "__main__":
    main()
