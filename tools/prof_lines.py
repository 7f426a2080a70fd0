"""Aggregate an ncu `--page source --csv --print-source cuda,sass` dump per CUDA source line (top stall samples)."""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
def num(x):
    try: return int(x)
    except Exception: return 0
cur=None; agg=[]; hdr=None
for row in csv.reader(open(path)):
    if not row: continue
    if row[0]=='File Path': cur=row[1].split('/')[-1]; continue
    if row[0]=='Function Name': continue
    if row[0]=='Line No': hdr=row; continue
    if hdr and row[0].strip().isdigit():
        d=dict(zip(hdr,row))
        agg.append((num(d.get('# Samples')), cur, int(row[0]), row[1].strip()[:88], num(d.get('Instructions Executed')), d.get('Avg. Threads Executed'), d))
tot=sum(a[0] for a in agg) or 1
print('total samples',tot)
agg.sort(key=lambda a:-a[0])
for s,f,l,src,inst,avgthr,d in agg[:topn]:
    st={k:num(v) for k,v in d.items() if k.startswith('stall_') and '(Not' not in k and num(v)>0}
    top=sorted(st.items(), key=lambda x:-x[1])[:3]
    print(f"{100*s/tot:5.1f}% {f}:{l:<4} inst={inst:<9} thr={avgthr:<3} {src}  {top}")
