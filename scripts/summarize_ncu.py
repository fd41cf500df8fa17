"""Turn an .ncu-rep into a small text summary for profiles/ (key metrics + hottest SASS lines)."""
import csv, io, re, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = re.compile(r'^(gpu__time_duration.sum|dram__bytes_(read|write).sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|'
                  r'lts__t_bytes.sum|lts__t_sector_hit_rate.pct|sm__warps_active.avg.pct_of_peak_sustained_active|'
                  r'sm__issue_active.avg.pct_of_peak_sustained_elapsed|smsp__inst_executed.sum|launch__(grid_size|block_size|registers_per_thread|'
                  r'shared_mem_per_block_dynamic|occupancy_limit_(shared_mem|registers|warps)|waves_per_multiprocessor)|'
                  r'smsp__average_warps_issue_stalled_(long_scoreboard|barrier|short_scoreboard|wait|membar|lg_throttle|mio_throttle)_per_issue_active.ratio|'
                  r'sm__inst_executed_pipe_(xu|fma|alu|lsu|uniform).sum|sm__throughput.avg.pct_of_peak_sustained_elapsed)$')
with open(out, 'w') as f:
    f.write('# ncu --set full summary of %s\n' % rep.split('/')[-1])
    for k, row in enumerate(rows[2:]):
        name = row[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        f.write('\n## launch %d: %s\n' % (k, name[:100]))
        for i, h in enumerate(hdr):
            if want.match(h):
                f.write('%-88s %s %s\n' % (h, row[i], units[i]))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(src)))
    try:
        hi = next(i for i, r in enumerate(rr) if r and r[0] == 'Address')
        body = [r for r in rr[hi + 1:] if len(r) > 5 and r[2].isdigit()]
        seen, first = set(), []
        for r in body:
            if r[0] in seen:
                break
            seen.add(r[0]); first.append(r)
        tot = sum(int(r[2]) for r in first) or 1
        f.write('\n## hottest SASS instructions by warp-stall samples (launch 0; total %d samples, %d instructions)\n' % (tot, len(first)))
        for i, r in sorted(sorted(enumerate(first), key=lambda x: -int(x[1][2]))[:25]):
            f.write('%5d %6s %5.1f%% exec=%8s  %s\n' % (i, r[2], 100 * int(r[2]) / tot, r[5], r[1].strip()[:100]))
    except StopIteration:
        pass
print('wrote', out)
