import csv, sys, re, collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
iS=hdr.index("Source"); iN=hdr.index("Instructions Executed"); iSm=hdr.index("# Samples"); iA=hdr.index("Address")
W=int(sys.argv[2]) if len(sys.argv)>2 else 100
steps=int(sys.argv[3]) if len(sys.argv)>3 else 148*16*834
totS=sum(int(r[iSm] or 0) for r in data)
for s0 in range(0,len(data),W):
    seg=data[s0:s0+W]
    ops=collections.Counter(); n=0; sm=0
    for r in seg:
        try: k=int(r[iN])
        except: continue
        m=re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS]); op=m.group(2) if m else '?'
        ops[op]+=k; n+=k; sm+=int(r[iSm] or 0)
    top=" ".join("%s:%.0f"%(o,c/steps) for o,c in ops.most_common(7))
    print("%5d %s exec/warp-step %6.1f samples %4.1f%%  %s"%(s0, seg[0][iA], n/steps, 100*sm/totS, top))
