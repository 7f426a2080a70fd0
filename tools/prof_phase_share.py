"""Instruction / stall-sample shares of tile_kernel per phase, from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass
--kernel-name regex:tile_kernel --launch-count 1 > dump.csv`:  python tools/prof_phase_share.py dump.csv [first_line last_line file]
(the line ranges below name the phases of blu_kernels.cu / blu_core.cuh as of the end of round 2)."""
import csv, collections, sys
hdr=None; rows=[]; cur=None
for row in csv.reader(open(sys.argv[1], errors='replace')):
    if not row: continue
    if row[0]=='File Path': cur=row[1].split('/')[-1]; continue
    if row[0]=='Function Name': continue
    if row[0]=='Line No': hdr=row; continue
    if hdr and row[0].strip().isdigit() and len(row)==len(hdr):
        # line-level rows: the Address column is empty
        ia=hdr.index('Address')
        if row[ia].strip() not in ('-',''): continue
        ii=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples')
        try: inst=int(row[ii] or 0)
        except: inst=0
        try: smp=int(row[isamp] or 0)
        except: smp=0
        rows.append((inst,cur,int(row[0]),row[1].strip()[:110],smp))
tot=sum(r[0] for r in rows); tots=sum(r[4] for r in rows)
print('total inst',tot,'samples',tots)
rng=[(440,460,'pack'),(461,483,'slow'),(484,512,'next_head/row_end'),(513,592,'classify_round'),(595,615,'drain_topq'),(615,700,'write_record/flush..'),(700,770,'seg init/loop top'),(770,835,'B loop+geom prefix'),(835,897,'warp15 duty'),(897,972,'phase D'),(972,1045,'E head/stats'),(1045,1150,'E decide'),(1150,1205,'E push/old'),(1205,1300,'F/next')]
c=collections.Counter(); cs=collections.Counter()
for i,f,l,s,smp in rows:
    key=None
    if f=='blu_kernels.cu':
        for a,b,n in rng:
            if a<=l<b: key=n; break
        else: key='other k'
    elif f=='blu_core.cuh':
        if 586<=l<=592: key='core bits64_at'
        elif 594<=l<=609: key='core swar'
        elif 612<=l<=706: key='core parse_row_lean'
        elif 709<=l<=716: key='core load_u32_unaligned'
        elif 736<=l<=761: key='core same_qid_lean'
        else: key='core other %d'%(l//100*100)
    else: key=f
    c[key]+=i; cs[key]+=smp
for k,v in c.most_common(30): print(f'{100*v/tot:5.1f}% inst {100*cs[k]/tots:5.1f}% samples  {k}')
print()
rows.sort(key=lambda r:-r[0])
for i,f,l,s,smp in rows[:45]: print(f'{100*i/tot:4.1f}% {100*smp/tots:4.1f}%s {f}:{l} {s}')
print()
lo,hi=(int(sys.argv[2]),int(sys.argv[3])) if len(sys.argv)>4 else (0,-1)
for i,f,l,s,smp in sorted(rows,key=lambda r:r[2]):
    if len(sys.argv)>4 and f==sys.argv[4] and lo<=l<=hi and i: print(f'{100*i/tot:4.2f}% {100*smp/tots:4.2f}%s {l} {s[:100]}')
