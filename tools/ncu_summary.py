"""Summarise an .ncu-rep (read with the ncu CLI, no GPU needed) into a small text file
for profiles/: duration, DRAM traffic, occupancy, issue rate, pipe utilisation, stall
reasons, and the executed instruction mix by opcode."""
import collections
import csv
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def source(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[1], rows[2:]


KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__issue_active.avg.per_cycle_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
]


def main():
    rep = sys.argv[1]
    units_per_launch = float(sys.argv[2]) if len(sys.argv) > 2 else None
    hdr, units, rows = raw(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        print("kernel:", r[ix["Kernel Name"]][:110])
        for k in KEYS:
            if k in ix:
                print(f"  {k:70s} {r[ix[k]]:>18s} {units[ix[k]]}")
        print("  stall reasons (warps stalled per issue-active cycle):")
        st = [(float(r[i]), h) for h, i in ix.items()
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        for v, h in sorted(st, reverse=True)[:9]:
            print(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {v:6.3f}")
        if units_per_launch:
            rd = float(r[ix["dram__bytes_read.sum"]])
            wr = float(r[ix["dram__bytes_write.sum"]])
            mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            rd *= mul[units[ix["dram__bytes_read.sum"]]]
            wr *= mul[units[ix["dram__bytes_write.sum"]]]
            print(f"  dram bytes per land-cell-timestep: {(rd + wr) / units_per_launch:.4f} "
                  f"(read {rd/1e6:.1f} MB + write {wr/1e6:.1f} MB per launch; algorithmic figure 352 B)")
            inst = float(r[ix["smsp__inst_executed.sum"]])
            print(f"  warp instructions per 32 cell-steps: {inst / (units_per_launch / 32):.1f}")
    sh, sd = source(rep)
    sx = {h: i for i, h in enumerate(sh)}
    ops = collections.Counter()
    samp = collections.Counter()
    for r in sd:
        t = r[sx["Source"]].split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") else t[0]
        op = op if op.startswith("MUFU") else op.split(".")[0]
        ops[op] += int(r[sx["Instructions Executed"]])
        samp[op] += int(r[sx["# Samples"]])
    tot = sum(ops.values())
    print(f"executed instruction mix ({len(sd)} static SASS instructions, {tot:.3e} warp-level executed):")
    for op, n in ops.most_common(18):
        per = f"  per cell-step {n / (units_per_launch / 32):7.1f}" if units_per_launch else ""
        print(f"  {op:10s} {100 * n / tot:5.1f}%{per}   stall samples {100 * samp[op] / max(1, sum(samp.values())):5.1f}%")


if __name__ == "__main__":
    main()
