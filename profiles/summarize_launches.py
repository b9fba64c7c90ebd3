import csv, collections, re, sys
f=sys.argv[1]
lines=[l for l in open(f) if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0,0.0]); seq=[]
for row in r:
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    if unit=='ns': v/=1000
    elif unit=='ms': v*=1000
    short=re.sub(r'\(.*','',row['Kernel Name']).replace('void ','')
    agg[short][0]+=1; agg[short][1]+=v; seq.append((short,round(v,1),row.get('Grid Size')))
tot=sum(v[1] for v in agg.values())
print("total us",round(tot,1),"launches",len(seq))
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{k:45s} n={n:4d} total={t:10.1f}us avg={t/n:8.1f}us share={t/tot:6.1%}")
n=int(sys.argv[2]) if len(sys.argv)>2 else 0
for s in seq[:n]: print(s)
