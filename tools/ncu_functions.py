import csv, sys, collections, re
path=sys.argv[1]
cur=None; lines={}
with open(path,newline='') as f:
    for r in csv.reader(f):
        if not r: continue
        if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
        if r[0]=='Function Name': continue
        if r[0]=='Line No': hdr=r; ii=hdr.index('Instructions Executed'); si=hdr.index('# Samples'); continue
        if r[0]=='': continue
        try: inst=float(r[ii]); samp=float(r[si])
        except ValueError: continue
        lines[(cur,int(r[0]))]=(inst,samp,r[1])
ti=sum(v[0] for v in lines.values()); ts=sum(v[1] for v in lines.values())
# function ranges from source files
import os
def funcs(fn):
    src=open('/root/repo/lossless-audio-codec_b200/csrc/'+fn).read().split('\n')
    out=[]  # (start,name)
    for i,l in enumerate(src,1):
        m=re.match(r'^(?:template.*\n)?(?:__device__|__global__|static|inline).*?(\w+)\s*\(', l)
        if l.startswith('__device__') or l.startswith('__global__'):
            m=re.search(r'(\w+)\s*\(', l.split('__forceinline__')[-1])
            if m: out.append((i,m.group(1)))
    return out
for fn in ('lacb_encode.cuh','lacb_common.cuh','lacb_enc_kernels.cuh','lacb_dec_kernels.cuh'):
    fs=funcs(fn)
    agg=collections.OrderedDict()
    for (f_,l),(i,s,_) in lines.items():
        if f_!=fn: continue
        name='?'
        for st,nm in fs:
            if st<=l: name=nm
        a=agg.setdefault(name,[0,0]); a[0]+=i; a[1]+=s
    print('==',fn)
    for nm,(i,s) in sorted(agg.items(), key=lambda kv:-kv[1][0]):
        if i/ti>0.002: print(f"  {nm:28s} inst {100*i/ti:5.2f}%  samp {100*s/ts:5.2f}%")
