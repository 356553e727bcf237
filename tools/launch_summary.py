import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki].split('(')[0][-40:],[]).append(float(r[vi].replace(',','')))
tot=sum(sum(v) for v in agg.values())
for k,v in agg.items(): print(f"{k:42s} n={len(v):4d} total={sum(v)/1e6:9.3f} ms ({100*sum(v)/tot:5.1f}%)  avg={sum(v)/len(v)/1e3:9.1f} us  max={max(v)/1e3:9.1f} us")
