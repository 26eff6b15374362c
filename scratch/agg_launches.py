import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hdr_i]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict(); tot=0
for r in rows[hdr_i+1:]:
    if len(r) <= vi: continue
    name = re.sub(r'^void ', '', r[ki]); name = re.sub(r'<unnamed>::', '', name); name = re.sub(r'\(.*', '', name)[:78]
    v = float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    a = agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v
print('total us', round(tot), 'launches', sum(a[0] for a in agg.values()))
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 28]:
    print(f'{a[1]:10.1f} us {a[0]:4d}x  {100*a[1]/tot:5.1f}%  {k}')
