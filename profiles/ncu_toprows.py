import csv,sys
r=csv.reader(sys.stdin); next(r); hdr=next(r)
ia=hdr.index('Instructions Executed'); isrc=hdr.index('Source'); ist=hdr.index('# Samples')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and '(' not in h]
rows=[]
for n,row in enumerate(r):
    if len(row)<=ia: continue
    st=sorted(((int(row[i] or 0),hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    rows.append((n,int(row[ist]),int(row[ia]),row[isrc].strip(),st))
lo,hi=int(sys.argv[1]),int(sys.argv[2])
sel=[x for x in rows if lo<=x[0]<hi]
for x in sorted(sel,key=lambda x:-x[1])[:int(sys.argv[3])]: print(x)
