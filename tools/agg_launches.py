import csv, re, collections, sys
def load(path):
    with open(path) as f:
        lines=[l for l in f if not l.startswith('==')]
    rows=[]
    for row in csv.DictReader(lines):
        if row.get('Metric Name')=='gpu__time_duration.sum':
            v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
            v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
            rows.append((re.sub(r'\(.*','',row['Kernel Name']).replace('void ','').replace('tocvp::',''), v, row.get('Grid Size','')))
    return rows
for p in sys.argv[1:]:
    rows=load(p); agg=collections.defaultdict(lambda:[0,0.0]); tot=sum(r[1] for r in rows)
    for n,v,g in rows: agg[n][0]+=1; agg[n][1]+=v
    print(f"== {p}: {len(rows)} launches, {tot/1e3:.2f} ms")
    for n,(c,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:14]:
        print(f"{v/1e3:9.2f} ms {100*v/tot:5.1f}%  x{c:5d}  avg {v/c:8.1f} us  {n[:80]}")
