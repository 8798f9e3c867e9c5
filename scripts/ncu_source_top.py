"""Top stalled SASS instructions of an `ncu --page source --csv` export.  usage: ncu_source_top.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 45
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
num = lambda r, k: int(float(r[ix[k]] or 0))
tot = sum(num(r, '# Samples') for r in data)
print('total samples', tot, 'instructions', len(data))
top = sorted(range(len(data)), key=lambda i: -num(data[i], '# Samples'))[:n]
for i in sorted(top):
    r = data[i]
    st = sorted(((s, num(r, s)) for s in stalls), key=lambda kv: -kv[1])[:3]
    print("%5d %6d %5.1f%%  %-64s %s" % (i, num(r, '# Samples'), 100.0 * num(r, '# Samples') / tot, r[ix['Source']][:64], [x for x in st if x[1]]))
