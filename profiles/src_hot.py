#!/usr/bin/env python
"""Per-source-line stall samples from an .ncu-rep (needs -lineinfo and --import-source on).
usage: python profiles/src_hot.py prof.ncu-rep [kernel_index] [top_n]"""
import csv, io, subprocess, sys, collections
rep=sys.argv[1]; ki=int(sys.argv[2]) if len(sys.argv)>2 else 0; topn=int(sys.argv[3]) if len(sys.argv)>3 else 40
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],stdout=subprocess.PIPE,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
kern=-1; hdr=None; cur_file=None; data=collections.defaultdict(list)
seen_fn=[]
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name':
        if r[1] not in seen_fn: seen_fn.append(r[1])
        kern=seen_fn.index(r[1]); continue
    if r[0]=='Line No': hdr=r; continue
    if hdr and len(r)==len(hdr) and r[0] not in ('',):
        data[kern].append((cur_file,r))
# NOTE: one report may hold several launches of the same function; they are merged here
ix={}
for i,n in enumerate(hdr):
    ix.setdefault(n,i)
rs=data[ki]
si=ix['# Samples']
tot=sum(int(r[si] or 0) for f,r in rs)
print('function:',seen_fn[ki][:90]); print('total samples',tot)
stalls=[n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
agg=collections.Counter()
for f,r in rs:
    for n in stalls: agg[n]+=int(r[ix[n]] or 0)
print({k[6:]:v for k,v in agg.most_common(9)})
top=sorted(rs,key=lambda fr:-int(fr[1][si] or 0))[:topn]
for f,r in sorted(top,key=lambda fr:(fr[0],int(fr[1][0]))):
    st={n[6:]:int(r[ix[n]] or 0) for n in stalls if int(r[ix[n]] or 0)>0.08*max(1,int(r[si] or 0))}
    print('%-16s %4s %6s %5.1f%% inst=%-9s'%(f,r[0],r[si],100*int(r[si] or 0)/tot,r[ix['Instructions Executed']]), st, '|', r[1].strip()[:70])
