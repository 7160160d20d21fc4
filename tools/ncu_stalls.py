"""ncu source-page CSV (`ncu -i x.ncu-rep --page source --csv --kernel-name regex:K > k.csv`) -> warp-stall reasons and opcode mix of
the kernel (all captured launches of it together).  Usage: python tools/ncu_stalls.py k.csv"""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
hdr=rows[hi]; c={h:i for i,h in enumerate(hdr)}
stalls=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot=collections.Counter(); ops=collections.Counter(); opsamp=collections.Counter()
n_exec=0; nst=0; nex=0
def I(x):
    try: return int(x)
    except: return 0
for r in rows[hi+1:]:
    if len(r)<len(hdr) or r[0]=="Address": continue
    nst+=1
    for s in stalls: tot[s]+=I(r[c[s]])
    src=r[c["Source"]].split()
    op=src[1] if src[0].startswith("@") else src[0]
    op=op.split(".")[0]
    e=I(r[c["Instructions Executed"]])
    if e: nex+=1
    ops[op]+=e; opsamp[op]+=I(r[c["# Samples"]]); n_exec+=e
S=sum(tot.values())
print("total samples",S,"executed warp instr",n_exec, "static instrs", nst, "with exec>0", nex)
for k,v in tot.most_common(12): print("  %-24s %6.1f%%"%(k,100*v/S))
print("by opcode (executed share, sample share)")
for k,v in ops.most_common(24): print("  %-10s %6.1f%% %6.1f%%"%(k,100*v/n_exec,100*opsamp[k]/max(1,sum(opsamp.values()))))
