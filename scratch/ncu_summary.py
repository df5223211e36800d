"""Summarise an .ncu-rep: per-kernel key metrics + top stalled SASS lines.  usage: ncu_summary.py rep [kernel_regex]"""
import csv, subprocess, sys, io, re
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    name = r[ki][:60]
    vals = []
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            vals.append(f"{w.split('.')[0].replace('__','_')[-28:]}={r[i]}{units[i]}")
    print(name, '|', ' '.join(vals))
if len(sys.argv) > 2:
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + sys.argv[2]],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    which = int(sys.argv[3]) if len(sys.argv) > 3 else -1
    st = starts[which]
    end = starts[which + 1] if which != -1 and which + 1 < len(starts) else len(rows)
    hdr = rows[st + 1]
    si, srci = hdr.index('# Samples'), hdr.index('Source')
    body = [r for r in rows[st + 2:end] if len(r) > si and r[si].isdigit()]
    tot = sum(int(r[si]) for r in body) or 1
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {}
    for r in body:
        for i in stall_cols:
            if r[i].isdigit():
                agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
    print('stall mix:', {k: f"{100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for r in sorted(body, key=lambda r: -int(r[si]))[:18]:
        s2 = sorted([(int(r[i]), hdr[i]) for i in stall_cols if r[i].isdigit()], reverse=True)[:2]
        print(f"{100*int(r[si])/tot:5.1f}%  {r[srci][:64]:64s} {s2}")
