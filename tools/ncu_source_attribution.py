"""Per-source-line attribution of an `ncu --set full --import-source on` capture (developer tool).

    ncu -i rep.ncu-rep --page source --csv > source.csv                       # SASS-level page of ONE kernel
    cuobjdump -xelf all build/sed_frontend.o && nvdisasm --print-line-info *.cubin > lines.txt    # needs -lineinfo
    python tools/ncu_source_attribution.py source.csv lines.txt Li512EfLi0E [units]

joins the SASS addresses of the report with the line table of the same binary (third argument: a substring of the
kernel's mangled name) and prints, per source line, instructions, stall samples and shared-memory wavefronts per
unit (default: 74074 frame pairs = 148 clips of 10 s at hop 160)."""
import csv,re,sys,collections
rep_csv, lines_txt, kern = sys.argv[1], sys.argv[2], sys.argv[3]
txt=open(lines_txt).read().split('\n')
starts=[i for i,l in enumerate(txt) if l.startswith('.text.')]
a=b=None
for si,s in enumerate(starts):
    if kern in txt[s]:
        a=s; b=starts[si+1] if si+1<len(starts) else len(txt)
addr2line={}; cur=None
for ln in txt[a:b]:
    m=re.search(r'//## File "([^"]+)", line (\d+)',ln)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+',ln)
    if m: addr2line[int(m.group(1),16)]=cur
rows=list(csv.reader(open(rep_csv)))
hdr=rows[1]
iA=hdr.index('Address'); iI=hdr.index('Instructions Executed'); iS=hdr.index('# Samples'); iW=hdr.index('L1 Wavefronts Shared'); iWi=hdr.index('L1 Wavefronts Shared Ideal'); iSrc=hdr.index('Source')
base=None
agg=collections.defaultdict(lambda:[0,0,0,0])
tot=[0,0,0,0]
for r in rows[2:]:
    if len(r)<len(hdr): continue
    ad=int(r[iA],16) if r[iA].startswith('0x') else int(r[iA])
    if base is None: base=ad
    key=addr2line.get(ad-base)
    vals=[int(float(r[iI] or 0)),int(float(r[iS] or 0)),int(float(r[iW] or 0)),int(float(r[iWi] or 0))]
    for i in range(4): agg[key][i]+=vals[i]; tot[i]+=vals[i]
pairs=float(sys.argv[4]) if len(sys.argv)>4 else 74074.0
print('total inst %d (%.0f/pair) samples %d wavefronts %d (%.0f/pair) ideal %d'%(tot[0],tot[0]/pairs,tot[1],tot[2],tot[2]/pairs,tot[3]))
regions=eval(open(sys.argv[5]).read()) if len(sys.argv)>5 else None
for k,v in sorted(agg.items(), key=lambda kv:(str(kv[0][0]) if kv[0] else '', kv[0][1] if kv[0] else 0)):
    if v[0]>0.002*tot[0] or v[1]>0.004*tot[1]:
        print(k, 'inst/pair %.1f  samples %.1f%%  wf/pair %.1f (ideal %.1f)'%(v[0]/pairs,100*v[1]/tot[1],v[2]/pairs,v[3]/pairs))
