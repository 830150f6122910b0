import csv, collections, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches_r1.csv'
with open(path) as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
seq=[]
for row in r:
    name=row['Kernel Name']; v=float(row['Metric Value'].replace(',',''))
    unit=row['Metric Unit']
    if unit=='ns': v/=1e3
    elif unit=='ms': v*=1e3
    short=re.sub(r'\(.*','',name).replace('void ','').replace('mimamo::','')
    seq.append((short,v,row.get('Grid Size'),row.get('Block Size')))
agg=collections.defaultdict(lambda:[0,0.0])
for s in seq: agg[s[0]][0]+=1; agg[s[0]][1]+=s[1]
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda x:-x[1][1]):
    print('%-44s n=%4d  %10.1f us  %5.1f%%'%(k[:44],v[0],v[1],100*v[1]/tot))
print('total us',tot, 'launches', len(seq))
if '--seq' in sys.argv:
    lo,hi=int(sys.argv[-2]),int(sys.argv[-1])
    for s in seq[lo:hi]:
        print('%-34s %8.1f us grid=%s'%(s[0][:34],s[1],s[2]))
