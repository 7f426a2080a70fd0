"""Instruction / stall-sample breakdown of an ncu source dump by kernel phase / helper function (line ranges located
by text markers in the sources)."""
import csv, sys
src = sys.argv[1]
def num(x):
    try: return int(x)
    except Exception: return 0
cur=None; hdr=None; agg={}
for row in csv.reader(open(src)):
    if not row: continue
    if row[0]=='File Path': cur=row[1].split('/')[-1]; continue
    if row[0]=='Function Name': continue
    if row[0]=='Line No': hdr=row; continue
    if hdr and row[0].strip().isdigit():
        d=dict(zip(hdr,row)); agg[(cur,int(row[0]))]=(num(d['Instructions Executed']), num(d['# Samples']), num(d['Thread Instructions Executed']))
def line_of(pat, f):
    txt=open('/root/repo/blutils_b200/csrc/'+f).read()
    i=txt.find(pat)
    return None if i<0 else txt.count('\n',0,i)+1
K='blu_kernels.cu'; C='blu_core.cuh'
marks={K:[('ptx wrappers','uint32_t smem_u32'),('longrun window code','struct WindowIndex'),('stream helpers','struct CarryRun'),('pack32','uint32_t pack8'),('classify_unit_slow','void classify_unit_slow'),('next_head','int next_head'),
          ('row_end_search','int row_end_search'),('phase B classify','void classify_round('),('write_record/flush','void write_record('),('tile prologue',') tile_kernel('),('window setup','    while (true) {\n        const uint8_t* const win'),('phase B call','---- phase B'),('geometry','---- row geometry'),('phase D rows','---- phase D'),
          ('phase E runs','---- phase E: one warp per query run'),('phase E decide (lane 0)','---- lane 0: what happens'),('phase E top rows','---- the run\'s top rows'),('where next','---- where next'),('longrun','// long-run kernel')],
       C:[('probe','uint32_t probe_taxid'),('parse_i64','bool parse_i64'),('parse_f64','uint32_t parse_f64'),('num_class/dfa/check_float','int num_class'),('light_parse_row','LightRow light_parse_row'),('ctz/next_tab/all_digits','int blu_ctz64'),('parse_row_masked','LightRow parse_row_masked'),('bit helpers','uint32_t blu_funnel_r'),('float_shape_ok','bool float_shape_ok'),('parse_row_fast','bool parse_row_fast'),('lean helpers','int blu_popc64'),('swar digits','uint32_t swar4('),('parse_row_lean','bool parse_row_lean'),('load_u32_unaligned','uint32_t load_u32_unaligned(const uint8_t* win, int pos) {'),('same_first_field','bool same_first_field'),('same_qid_lean','bool same_qid_lean'),('TopRow/heavy','struct TopRow'),('parse shorts','bool parse_u32_short'),('split_top_row','uint32_t split_top_row('),('top_row_from_info','bool top_row_from_info('),('join/heavy_masked','uint32_t join_top_row'),('consensus','struct QueryOut')]}
ti=sum(v[0] for v in agg.values()) or 1; ts=sum(v[1] for v in agg.values()) or 1
print('total warp-inst',ti,'samples',ts)
for f,ms in marks.items():
    ms=[(n,line_of(p,f)) for n,p in ms]; ms=sorted([m for m in ms if m[1]], key=lambda m:m[1])
    for (n,a),(n2,b) in zip(ms, ms[1:]+[('end',10**9)]):
        i=sum(v[0] for k,v in agg.items() if k[0]==f and a<=k[1]<b); s=sum(v[1] for k,v in agg.items() if k[0]==f and a<=k[1]<b); t=sum(v[2] for k,v in agg.items() if k[0]==f and a<=k[1]<b)
        if i: print(f"{f:15s} {n:26s} inst {100*i/ti:5.1f}%  samples {100*s/ts:5.1f}%  thr/inst {t/max(i,1):5.1f}")
oth=sum(v[0] for k,v in agg.items() if k[0] not in (K,C)); print('other files inst %.1f%%'%(100*oth/ti))
