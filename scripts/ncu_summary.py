"""Summarise an .ncu-rep: key raw metrics + top source lines by stall samples."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'sm__cycles_elapsed.avg',
        # the gather path: requests and sectors between L1 and L2
        'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_sectors.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum', 'lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
        'dram__sectors_read.sum', 'dram__sectors_write.sum']
keys += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        vals = [r[i][:70] for r in data]
        try:
            if all(float(v) < 0.05 for v in vals) and 'stalled' in k: continue
        except ValueError: pass
        print(f"{k} [{units[i]}]: {vals}")
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # find header row
    for n, r in enumerate(rows):
        if 'Source' in r and any('Sampling' in c for c in r):
            h = r; body = rows[n+1:]; break
    else:
        print("no source page"); sys.exit()
    si = h.index('Source'); ci = [i for i, c in enumerate(h) if c.startswith('# Samples') or 'Warp Stall Sampling (All' in c][0]
    tot = sum(float(r[ci] or 0) for r in body if len(r) > ci)
    top = sorted(body, key=lambda r: -float(r[ci] or 0))[:int(sys.argv[2])]
    print("total samples", tot)
    for r in top:
        print(f"{float(r[ci])/tot*100:5.1f}%  {r[si][:110]}")
