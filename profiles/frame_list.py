import csv,re,sys
fn=sys.argv[1]
rows=[]
lines=[l for l in open(fn) if not l.startswith('==')]
for row in csv.DictReader(lines):
    if row.get('Metric Name')=='gpu__time_duration.sum':
        v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
        v = v/1e6 if u=='ns' else v/1e3 if u=='us' else v*1e3 if u=='s' else v
        rows.append((int(row['ID']), row['Kernel Name'], v))
idx=[i for i,(id_,k,v) in enumerate(rows) if 'clear_image' in k]
a,b=idx[-3],idx[-2]
tot=0
for id_,k,v in rows[a:b]:
    kk=re.sub(r'\(.*','',k)[:70]
    print(f"  {kk:72s} {v:8.4f} ms"); tot+=v
print("  total", tot)
