import csv,sys,collections
r=csv.reader(sys.stdin)
next(r); hdr=next(r)
ia=hdr.index('Instructions Executed'); isrc=hdr.index('Source'); iw=hdr.index('L1 Wavefronts Shared'); ist=hdr.index('# Samples')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and '(' not in h]
ops=collections.Counter(); samp=collections.Counter(); tot=0; wf=0; st=collections.Counter(); rows=[]
for row in r:
    if len(row)<=ia: continue
    s=row[isrc].split()
    if not s: continue
    op=s[0] if not s[0].startswith('@') else s[1]
    op=op.split('.')[0]
    n=int(row[ia]); ops[op]+=n; tot+=n; samp[op]+=int(row[ist]); wf+=int(row[iw] or 0)
    for i in stall_cols: st[hdr[i]]+=int(row[i] or 0)
    rows.append((int(row[ist]), n, row[isrc].strip()))
N=int(sys.argv[1]) if len(sys.argv)>1 else 1<<20
print('total/query',tot/N, 'smem wf/query', wf/N)
for k,v in ops.most_common(24): print(f'{k:10s} {v/N:7.2f} /query   samples {samp[k]}')
print({k:v for k,v in st.most_common(10)})
if len(sys.argv)>2:
    for i,(a,b,c) in enumerate(rows): print(i,a,b,c)
