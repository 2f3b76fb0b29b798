import csv, collections, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0,0.0]); seq=[]
for row in r:
    name=row['Kernel Name']
    short=name.split('<')[0].replace('void ','')
    if 'mmvq' in name:
        import re
        m=re.search(r'mmvq_kernel<(\d+), (-?\d+), (\d+)>', name)
        short='mmvq<n=%s,t=%s,mode=%s>'%m.groups() if m else short
        short+=' grid=%s'%row.get('Grid Size','')
    try: v=float(row['Metric Value'].replace(',',''))
    except: continue
    unit=row['Metric Unit']
    us = v/1000.0 if unit in('ns','nsecond') else (v if unit in ('us','usecond') else v*1000)
    agg[short][0]+=1; agg[short][1]+=us; seq.append((short,us))
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{k:62s} n={v[0]:5d} total={v[1]:10.1f}us avg={v[1]/v[0]:8.2f}us share={100*v[1]/tot:5.1f}%")
print("total us", tot, "launches", len(seq))
if len(sys.argv)>2:
    for s in seq[int(sys.argv[2]):int(sys.argv[2])+40]: print("   %-60s %8.2f"%s)
