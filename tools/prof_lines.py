import re,csv,glob,subprocess,sys,os
rep=sys.argv[1]; binp=sys.argv[2]; pat=sys.argv[3]
os.system('rm -rf /tmp/cub && mkdir /tmp/cub && cd /tmp/cub && cuobjdump -xelf all /root/repo/%s >/dev/null 2>&1 && for f in *.cubin; do nvdisasm -g -c $f > $f.txt 2>/dev/null; done' % binp)
os.system('ncu -i %s --page source --csv 2>/dev/null > /tmp/hub_src.csv' % rep)
txt=open(glob.glob('/tmp/cub/*.txt')[0]).read()
m=re.search(r'\.text\.[^\n]*'+pat+r'[^\n]*\n', txt)
start=m.end()
nxt=txt.find('.section', start+10)
body=txt[start: nxt if nxt>0 else None]
cur=None; lines=[]
for ln in body.split('\n'):
    mm=re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if mm:
        cur=(mm.group(1).split('/')[-1], int(mm.group(2))); continue
    mo=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if mo: lines.append((int(mo.group(1),16), cur, mo.group(2)))
rows=list(csv.reader(open('/tmp/hub_src.csv')))
hdr=rows[1]; data=rows[2:]
isamp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed')
base=int(data[0][0],16)
from collections import defaultdict
agg=defaultdict(lambda:[0,0])
off2line={o:l for o,l,_ in lines}
for r in data:
    off=int(r[0],16)-base
    l=off2line.get(off)
    agg[l][0]+=int(r[isamp]); agg[l][1]+=int(r[iex])
srcs={}
def text(l):
    if not l: return ''
    for root in ['/root/repo/graphblast_b200/csrc/graphblas/backend/cuda/kernels/','/root/repo/include/graphblas/']:
        f=root+l[0]
        if os.path.exists(f):
            if f not in srcs: srcs[f]=open(f).read().split('\n')
            return srcs[f][l[1]-1].strip()[:80]
    return ''
tot_s=sum(v[0] for v in agg.values()); tot_e=sum(v[1] for v in agg.values())
print('total inst',tot_e,'samples',tot_s)
thr=float(sys.argv[4]) if len(sys.argv)>4 else 1.0
for l,v in sorted(agg.items(), key=lambda kv: (kv[0] is None, kv[0])):
    if 100*v[1]/tot_e>=thr or 100*v[0]/tot_s>=thr:
        print(l, 's %5.1f%%'%(100*v[0]/tot_s), 'i %5.1f%%'%(100*v[1]/tot_e), text(l))
