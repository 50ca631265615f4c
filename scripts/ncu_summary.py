"""Summarise an .ncu-rep: key metrics, stall reasons, and the hottest source lines (needs -lineinfo)."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'launch__block_size','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed.sum',
 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warps_eligible.avg.per_cycle_active','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_alu.sum',
 'sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_fp64.sum','smsp__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_uniform.sum','sm__inst_executed_pipe_cbu.sum','sm__inst_executed_pipe_adu.sum',
 'lts__t_bytes.sum','l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum','smsp__thread_inst_executed_per_inst_executed.ratio','launch__shared_mem_per_block_dynamic']
for r in rows[2:]:
    print('--- kernel', r[hdr.index('Kernel Name')][:60])
    for k in keys:
        if k in hdr: print(f"  {k:72s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    st = []
    for i, h in enumerate(hdr):
        if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
            try: st.append((float(r[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError: pass
    for v, n in sorted(st, reverse=True)[:8]: print(f"  STALL(per issue) {n:36s} {v:8.2f}")
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # find header row
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and any('Instructions Executed' in c for c in r))
    h = rows[hi]
    print(h)
