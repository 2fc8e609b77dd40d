"""Print an ncu --csv launch list (one line per launch, chosen metrics)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
H = rows[hdr]
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < len(H): continue
    rec = dict(zip(H, r))
    d.setdefault((int(rec['ID']), rec['Kernel Name'].split('(')[0][-34:]), {})[rec['Metric Name']] = rec['Metric Value'].replace(',', '')
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
for (i, k), v in d.items():
    if lo <= i < hi:
        t = float(v.get('gpu__time_duration.sum', 0)) / 1e3
        rd = float(v.get('dram__bytes_read.sum', 0)) / 1e6
        wr = float(v.get('dram__bytes_write.sum', 0)) / 1e6
        extra = ' '.join('%s=%s' % (a.split('.')[0][-18:], b) for a, b in v.items() if a not in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum'))
        print('%3d %-34s %8.1f us  rd %7.1f MB wr %6.1f MB  %5.2f TB/s  %s' % (i, k, t, rd, wr, (rd + wr) / max(t, 1e-9) / 1e6 * 1e0, extra))
