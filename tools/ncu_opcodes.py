import csv, sys, re, collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
iA=hdr.index("Address"); iS=hdr.index("Source"); iN=hdr.index("Instructions Executed"); iSm=hdr.index("# Samples")
iW=hdr.index("L1 Wavefronts Shared")
steps=int(sys.argv[2]) if len(sys.argv)>2 else 148*16*834
ops=collections.Counter(); samp=collections.Counter(); wf=collections.Counter()
tot=0; totS=0
for r in data:
    try: n=int(r[iN])
    except: continue
    m=re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
    op=m.group(2) if m else '?'
    ops[op]+=n; tot+=n
    s=int(r[iSm] or 0); samp[op]+=s; totS+=s
    try: wf[op]+=int(r[iW] or 0)
    except: pass
print("total inst/warp-step %.0f"%(tot/steps))
for op,n in ops.most_common(40):
    print("%-10s %8.1f /warp-step   samples %5.1f%%  smem wf/warp-step %.0f"%(op, n/steps, 100*samp[op]/max(totS,1), wf[op]/steps))
