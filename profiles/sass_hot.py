#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel in an ncu source-page CSV.
usage: python profiles/sass_hot.py file.csv [kernel_index] [top_n]"""
import csv, sys, collections, re
rows=list(csv.reader(open(sys.argv[1])))
ki=int(sys.argv[2]) if len(sys.argv)>2 else 0
topn=int(sys.argv[3]) if len(sys.argv)>3 else 40
kern=[]; cur=None
for r in rows:
    if r and r[0]=="Kernel Name": cur={'name':r[1],'rows':[]}; kern.append(cur); continue
    if r and r[0]=="Address": cur['hdr']=r; continue
    if cur is not None and len(r)>10: cur['rows'].append(r)
k=kern[ki]; h=k['hdr']; ix={n:i for i,n in enumerate(h)}
rs=k['rows']
tot=sum(int(r[ix['# Samples']] or 0) for r in rs)
ex=sum(int(r[ix['Instructions Executed']] or 0) for r in rs)
print('kernel',ki,'lines',len(rs),'samples',tot,'warp-inst',ex)
st=collections.Counter()
for r in rs:
    for n in h:
        if n.startswith('stall_') and 'Not Issued' not in n: st[n]+=int(r[ix[n]] or 0)
print({k:v for k,v in st.most_common(8)})
opc=collections.Counter(); ops=collections.Counter()
for r in rs:
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)',r[ix['Source']]); op=m.group(2) if m else '?'
    opc[op]+=int(r[ix['Instructions Executed']] or 0); ops[op]+=int(r[ix['# Samples']] or 0)
print('exec by opcode:',[(o,c) for o,c in opc.most_common(16)])
print('samples by opcode:',[(o,c) for o,c in ops.most_common(12)])
order=sorted(range(len(rs)), key=lambda i:-int(rs[i][ix['# Samples']] or 0))[:topn]
for i in sorted(order):
    r=rs[i]
    print(i, r[ix['# Samples']], r[ix['Instructions Executed']], 'lsb',r[ix['stall_long_sb']],'bar',r[ix['stall_barrier']],'noi',r[ix['stall_no_inst']],'wait',r[ix['stall_wait']],'ssb',r[ix['stall_short_sb']],'|', r[ix['Source']][:80])
