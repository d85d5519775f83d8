import csv, collections, sys
f=sys.argv[1]; half=len(sys.argv)>2
lines=[l for l in open(f) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
if half: rows=rows[len(rows)//2:]
per=collections.OrderedDict()
for row in rows:
    k=row['Kernel Name'].split('(')[0]; v=float(row['Metric Value'].replace(',',''))
    u=row['Metric Unit']
    v = v/1e6 if u=='ns' else (v/1e3 if u=='us' else v)
    per.setdefault(k,[]).append(v)
tot=sum(sum(v) for v in per.values())
for k,v in per.items(): print('%-22s n=%2d total %7.3f ms (%4.1f%%) max %6.3f'%(k,len(v),sum(v),100*sum(v)/tot,max(v)))
print('total %.3f ms'%tot)
