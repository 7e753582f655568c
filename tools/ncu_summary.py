"""Summarise an .ncu-rep (raw page + source page) into text: key metrics, instruction mix per unit, top stall lines."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]; units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, val = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_op_shared', 'sm__inst_executed_pipe_lsu', 'smsp__warp_issue_stalled', 'lts__t_bytes.sum', 'sm__cycles_elapsed.max',
        'smsp__average_warp_latency_per_inst_issued', 'smsp__average_warps_issue_stalled', 'launch__shared_mem_per_block_dynamic', 'smsp__cycles_active.avg', 'local_load', 'local_store', 'lmem']
for h, u, v in zip(hdr, unit, val):
    if any(h == w or (w in h and ('stalled' in w or 'local' in w or 'lmem' in w or 'average_warp' in w or 'op_shared' in w)) for w in want):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp, tot, lines = collections.Counter(), collections.Counter(), 0, []
for r in rows[2:]:
    if len(r) <= iE: continue
    parts = r[iS].split()
    if not parts: continue
    op = parts[1] if parts[0].startswith('@') else parts[0]
    base = op.split('.')[0]
    n = int(r[iE] or 0); ops[base] += n; tot += n; samp[base] += int(r[iN] or 0)
    lines.append((int(r[iN] or 0), n, r[iS].strip()))
tsamp = sum(samp.values())
print("total warp-instr", tot, "samples", tsamp)
if units:
    print("warp-instr per warp-unit (32 units):", tot / (units / 32))
for k, v in ops.most_common(28):
    print(f"{k:10s} {100*v/tot:6.2f}% instr  {100*samp[k]/max(tsamp,1):6.2f}% samples" + (f"  {v/(units/32):8.1f}/unit" if units else ""))
print("--- top stall lines")
for s_, n, t in sorted(lines, reverse=True)[:25]:
    print(f"{100*s_/max(tsamp,1):5.2f}%  exec {n:10d}  {t}")
