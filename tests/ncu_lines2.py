import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
tiles=float(sys.argv[2])
cur='';hdr=None;out=[]
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]
    elif len(r)>10 and r[0]=='Line No': hdr=r
    elif hdr and len(r)==len(hdr) and r[0].isdigit():
        ie=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples')
        stalls={hdr[i]:int(r[i] or 0) for i in range(len(hdr)) if hdr[i].startswith('stall_') and 'Not Issued' not in hdr[i] and r[i].isdigit()}
        out.append((int(r[ie] or 0), int(r[isamp] or 0), cur, int(r[0]), r[1].strip()[:100], stalls))
tot=sum(o[0] for o in out); ts=sum(o[1] for o in out)
print('total',tot, 'per tile', tot/tiles, 'samples', ts)
n=int(sys.argv[3]) if len(sys.argv)>3 else 30
for e,s,f,ln,src,st in sorted(out,key=lambda o:-o[1])[:n]:
    top=", ".join(f"{k[6:]}={v}" for k,v in sorted(st.items(),key=lambda kv:-kv[1])[:2] if v)
    print(f"smp {100*s/ts:4.1f}% ins {e/tiles:6.0f}/tile {f}:{ln:4d} {src} [{top}]")
