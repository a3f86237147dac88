#!/usr/bin/env python
"""Key metrics per kernel from an .ncu-rep (via `ncu -i rep --page raw --csv`)."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_xu.sum',
        'smsp__inst_executed_pipe_xu.sum', 'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.sum',
        'smsp__cycles_active.avg', 'launch__grid_size', 'launch__waves_per_multiprocessor', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'local_load_bytes', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
for r in rows[2:]:
    print('----', r[idx['Kernel Name']][:90])
    for w in want:
        if w in idx:
            print(f"  {w:72s} {r[idx[w]]:>16s} {units[idx[w]]}")
