"""Dev helper: per-region stall-sample accumulation from an ncu source-page CSV."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
lo=int(sys.argv[2]) if len(sys.argv)>2 else 0
hi=int(sys.argv[3]) if len(sys.argv)>3 else 10**9
keys=sys.argv[4].split(',') if len(sys.argv)>4 else ['SYNCS','LDTM','STTM','UTC','BAR','WARPSYNC']
h=next(i for i,r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
ix={c:i for i,c in enumerate(rows[h])}
data=[]
for n,r in enumerate(rows[h+1:]):
    try:
        s=int(r[ix['# Samples']]); e=int(r[ix['Instructions Executed']])
    except: continue
    data.append((n,s,e,r[ix['Source']]))
tot=sum(d[1] for d in data)
print('total samples',tot,'instrs',len(data))
acc=0;cnt=0
for n,s,e,src in data:
    if n<lo or n>hi: continue
    acc+=s;cnt+=1
    if any(k in src for k in keys):
        print(n,f'acc {100*acc/tot:5.2f}% ({cnt:3d} instr) self {100*s/tot:5.2f}%',e,src[:90])
        acc=0;cnt=0
