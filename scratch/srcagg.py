import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1])))
cur=None; out=[]; hdr=None; kern=None; first=None
for r in rows:
    if len(r)==2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)==2 and r[0]=='Function Name':
        if first is None: first=r[1]
        kern=r[1]; continue
    if r and r[0]=='Line No': hdr=r; continue
    if kern!=first: continue
    if hdr and len(r)==len(hdr) and r[0].isdigit():
        d=dict(zip(hdr,r))
        out.append((cur,int(r[0]),r[1].strip()[:80],int(d['Instructions Executed']),int(d['# Samples'])))
tot=sum(o[3] for o in out); ts=sum(o[4] for o in out)
print('kernel',first,'total inst',tot,'samples',ts)
byfile=collections.Counter(); sf=collections.Counter()
for o in out: byfile[o[0]]+=o[3]; sf[o[0]]+=o[4]
for k,v in byfile.most_common(): print(k, v, '%.1f%%'%(100*v/tot), 'samples %.1f%%'%(100*sf[k]/ts))
N=int(sys.argv[2]) if len(sys.argv)>2 else 40
print('--- top lines by samples')
for o in sorted(out,key=lambda o:-o[4])[:N]: print('%-20s %4d inst=%5.2f%% samp=%5.2f%% %s'%(o[0],o[1],100*o[3]/tot,100*o[4]/ts,o[2]))
