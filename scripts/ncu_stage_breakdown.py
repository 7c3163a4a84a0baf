import csv,sys,subprocess
rep,kern=sys.argv[1],sys.argv[2]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name','regex:'+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
data=rows[2:]
keys=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
marks=[]
prev=None
for n,r in enumerate(data):
    src=r[idx['Source']]
    toks=src.split()
    op=toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '')
    base=op.split('.')[0]
    if base in ('BAR','UTCHMMA','SYNCS','UBLKCP','LDTM','STTM','REDG','RED','ATOMG','STG','LDGSTS'):
        if base!=prev: marks.append(n)
        prev=base
    
marks=[0]+marks+[len(data)]
tot=sum(int(r[idx['# Samples']] or 0) for r in data)
toti=sum(int(r[idx['Instructions Executed']] or 0) for r in data)
print('total samples',tot,'inst',toti)
for a,b in zip(marks[:-1],marks[1:]):
    s=sum(int(r[idx['# Samples']] or 0) for r in data[a:b])
    i=sum(int(r[idx['Instructions Executed']] or 0) for r in data[a:b])
    if s<tot*0.004: continue
    st={k:sum(int(r[idx[k]] or 0) for r in data[a:b]) for k in keys}
    top=sorted(st.items(),key=lambda x:-x[1])[:4]
    print(f"{a:5d}-{b:5d} {data[a][idx['Source']][:34]:34s} {100*s/tot:6.2f}% inst {100*i/toti:6.2f}%  "+' '.join(f"{k[6:]}={100*v/max(s,1):.0f}%" for k,v in top))
