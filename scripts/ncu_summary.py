"""Summarise ncu CSV exports (raw page + source page) of one kernel: key metrics, stall mix, per-opcode totals."""
import collections
import csv
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
d = dict(zip(rows[0], rows[-1]))
u = dict(zip(rows[0], rows[1]))
print("kernel:", d.get("Kernel Name"))
for k in ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct", "smsp__inst_executed.sum",
          "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
          "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__t_sectors_op_red.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]:
    if k in d:
        print("  %-80s %-10s %s" % (k, u.get(k, ""), d[k]))
print("stalls (warps per issue-active cycle):")
for k, v in sorted(d.items(), key=lambda kv: kv[0]):
    if "issue_stalled" in k and "per_issue_active" in k and float(v) > 0.1:
        print("  %-22s %s" % (k.split("issue_stalled_")[1].split("_per")[0], v))
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(collections.Counter)
tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":      # a second kernel section / trailer in the CSV
        continue
    toks = r[ix["Source"]].split()
    if not toks:
        continue
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    n = int(r[ix["# Samples"]] or 0)
    tot += n
    agg[op]["samples"] += n
    agg[op]["exec"] += int(r[ix["Instructions Executed"]] or 0)
    for k in ("stall_no_inst", "stall_short_sb", "stall_long_sb", "stall_mio", "stall_math", "stall_wait", "stall_branch_resolving",
              "stall_barrier", "stall_dispatch", "stall_lg", "stall_selected"):
        agg[op][k] += int(r[ix[k]] or 0)
texec = sum(c["exec"] for c in agg.values())
print("per-opcode (warp-level executed instructions, stall samples); total executed %d, samples %d" % (texec, tot))
for op, c in sorted(agg.items(), key=lambda kv: -kv[1]["exec"])[:16]:
    print("  %-8s exec %12d (%4.1f%%)  samples %7d  %s" % (op, c["exec"], 100.0 * c["exec"] / texec, c["samples"],
          {k[6:]: v for k, v in c.items() if k.startswith("stall_") and v > 0.01 * tot}))
