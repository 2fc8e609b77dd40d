import csv, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    d = {}
    for r in rows[1:]:
        d.setdefault(r[iid], {'k': r[ik]})[r[im]] = r[iv]
    print(f)
    for k, v in d.items():
        name = v['k'].replace('void ', '')[:28]
        print(f"{k:>3} {name:28s} {float(v.get('gpu__time_duration.sum',0))/1e3:9.1f} us  rd {float(v.get('dram__bytes_read.sum',0))/1e6:8.1f} MB  wr {float(v.get('dram__bytes_write.sum',0))/1e6:7.1f} MB  inst {float(v.get('smsp__inst_executed.sum',0))/1e6:7.1f} M")
