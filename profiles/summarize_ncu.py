#!/usr/bin/env python3
"""Turn an .ncu-rep (brought back from the GPU box in gpurun_out/) into the text summary committed under
profiles/: key raw metrics, stall reasons, SASS opcode mix and the hottest CUDA source lines.
Usage: python profiles/summarize_ncu.py gpurun_out/X.ncu-rep > profiles/X.txt"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
print("# %s" % rep)
print("kernel: %s   grid %s block %s" % (vals[hdr.index("Kernel Name")], vals[hdr.index("Grid Size")], vals[hdr.index("Block Size")]))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        print("%-70s %s %s" % (k, vals[i], units[i]))
st = []
for i, k in enumerate(hdr):
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
        try:
            st.append((float(vals[i]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
print("\nwarp stall reasons (warps stalled per issue-active cycle):")
print("  " + "  ".join("%s=%.2f" % (k, v) for v, k in sorted(st, reverse=True)[:10]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, h = None, None
lines, ops, ops_s = collections.OrderedDict(), collections.Counter(), collections.Counter()
tot = tots = 0
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        h = r
        iI, iS, iT = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
        continue
    if h is None or len(r) < len(h):
        continue
    try:
        n, s, t = int(r[iI]), int(r[iS]), int(r[iT])
    except ValueError:
        continue
    if r[2] == "-":
        key = (cur, r[0], r[1].strip()[:100])
        a = lines.get(key, (0, 0, 0))
        lines[key] = (a[0] + n, a[1] + s, a[2] + t)
        tot += n
        tots += s
    else:
        sass = r[3]
        op = (sass.split()[1] if sass.startswith("@") else sass.split()[0]).split(".")[0]
        ops[op] += n
        ops_s[op] += s
print("\nSASS opcode mix (%% of warp instructions executed / %% of stall samples), total %.3e warp instructions:" % tot)
osum, ssum = sum(ops.values()), sum(ops_s.values())
print("  " + "  ".join("%s=%.1f/%.1f" % (op, 100.0 * n / osum, 100.0 * ops_s[op] / max(ssum, 1)) for op, n in ops.most_common(18)))
print("\nhottest source lines by warp instructions (%% warp instructions, %% samples, mean active lanes):")
for k, (n, s_, t) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("  %5.1f%% %5.1f%% %5.1f  %s:%s  %s" % (100.0 * n / tot, 100.0 * s_ / max(tots, 1), t / max(n, 1), k[0], k[1], k[2]))
print("\nhottest source lines by stall samples (%% warp instructions, %% samples, mean active lanes):")
for k, (n, s, t) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("  %5.1f%% %5.1f%% %5.1f  %s:%s  %s" % (100.0 * n / tot, 100.0 * s / max(tots, 1), t / max(n, 1), k[0], k[1], k[2]))
