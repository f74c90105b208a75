"""Extract the metrics DESIGN.md / bench.py cite from an .ncu-rep into a small CSV.
   python profiles/ncu_extract.py gpurun_out/prof.ncu-rep profiles/out.csv"""
import csv, subprocess, sys
KEEP = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'sm__cycles_active.avg', 'sm__cycles_active.max']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
stall = [h for h in hdr if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
cols = [h for h in KEEP if h in hdr] + stall
with open(sys.argv[2], 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['metric', 'unit'] + ['launch_%d' % i for i in range(len(rows) - 2)])
    for h in cols:
        i = hdr.index(h)
        w.writerow([h.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('smsp__average_warp_latency_issue_stalled_', 'stall_'),
                    units[i]] + [r[i] for r in rows[2:]])
print('wrote', sys.argv[2], len(cols), 'metrics x', len(rows) - 2, 'launches')
