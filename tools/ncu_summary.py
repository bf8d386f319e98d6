"""Summarise an ncu report: python tools/ncu_summary.py report.ncu-rep [kernel regex for the source page]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H = rows[0]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    name = r[H.index('Kernel Name')].split('(')[0]
    print("==", name)
    for w in want:
        if w in H:
            print("   %-85s %s %s" % (w, r[H.index(w)], rows[1][H.index(w)]))
    st = [(h, r[H.index(h)]) for h in H if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
    tot = sum(float(v) for _, v in st if v)
    print("   stalls:", ", ".join("%s %.0f%%" % (h.split('stalled_')[1], 100 * float(v) / tot) for h, v in sorted(st, key=lambda x: -float(x[1] or 0))[:8]))
