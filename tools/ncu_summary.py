"""Prints the headline ncu metrics (and top stall reasons) of every kernel in a .ncu-rep: python tools/ncu_summary.py rep..."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'launch__grid_size', 'launch__block_size',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f'== {rep}: {r[hdr.index("Kernel Name")][:100]}')
        for w in WANT:
            if w in hdr:
                print(f'   {w:70s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}')
        st = [(h, i) for i, h in enumerate(hdr)
              if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
        vals = sorted([(float(r[i].replace(',', '')) if r[i] else 0.0, h) for h, i in st], reverse=True)[:5]
        print('   stalls/issue:', ', '.join(f'{h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")}={v:.2f}' for v, h in vals))
