"""Prints the metrics we care about from `ncu -i X.ncu-rep --page raw --csv` output (one block per launch)."""
import csv, sys
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'sm__issue_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg']
STALL = 'smsp__warp_issue_stalled_'


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('-' * 100)
        for w in WANT:
            if w in idx:
                print('%-72s %s %s' % (w, r[idx[w]][:100], units[idx[w]]))
        st = [(float(r[i].replace(',', '') or 0), h) for h, i in idx.items()
              if h.startswith(STALL) and h.endswith('_per_warp_active.pct')]
        for v, h in sorted(st, reverse=True)[:7]:
            print('   stall %-50s %6.1f %%' % (h[len(STALL):-len('_per_warp_active.pct')], v))


if __name__ == '__main__':
    main(sys.argv[1])
