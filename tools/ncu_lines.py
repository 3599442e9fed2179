"""Developer tool: join an ncu report's per-SASS-instruction counters with the
line info of the built cubin and print instruction / stall shares per source
line and per engine function.   python tools/ncu_lines.py REPORT.ncu-rep [kernel-substring]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
ksub = sys.argv[2] if len(sys.argv) > 2 else 'Li96ELi12ELi8ELi64'
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30

raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
for h, v in zip(hdr, vals):
    if h in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
             'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
             'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
             'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
             'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
             'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
             'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
             'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
             'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
             'lts__t_bytes.sum', 'launch__grid_size') or 'issue_stalled' in h and h.endswith('.ratio'):
        try:
            if 'issue_stalled' in h and float(v) < 0.3: continue
        except ValueError: pass
        print('%-90s %s' % (h, v))

sass_csv = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                          capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass_csv)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) >= len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}

tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'subpixal_b200/_native/libsubpixal_b200.so')],
               cwd=tmp, capture_output=True)
sass = subprocess.run(['nvdisasm', '-gi', '-c', os.path.join(tmp, 'spx_cc.sm_100a.cubin')],
                      capture_output=True, text=True).stdout.split('\n')
start = end = None
for i, l in enumerate(sass):
    if l.startswith('.text.'):
        if start is not None and end is None: end = i
        if ksub in l: start, end = i, None
sec = sass[start:end]
cur, seq, chain, fresh = None, [], [], True
for l in sec:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        loc = (m.group(1).split('/')[-1], int(m.group(2)))
        if fresh: cur, chain, fresh = loc, [loc], False
        else: chain.append(loc)
        m2 = re.search(r'inlined at "([^"]+)", line (\d+)', l)
        if m2: chain.append((m2.group(1).split('/')[-1], int(m2.group(2))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append((cur, m.group(2), list(chain))); fresh = True
assert len(seq) == len(data), (len(seq), len(data))
src = {}
for f in os.listdir(os.path.join(ROOT, 'subpixal_b200/csrc')):
    src[f] = open(os.path.join(ROOT, 'subpixal_b200/csrc', f)).read().split('\n')
funcs = [(i + 1, m.group(1)) for i, l in enumerate(src['spx_cc_engine.cuh'])
         for m in [re.search(r'SPX_HD [\w:<> ]+? (\w+)\(', l)] if m]
def fn(line):
    name = None
    for a, n in funcs:
        if a <= line: name = n
    return name
agg = collections.defaultdict(lambda: [0, 0]); byop = collections.Counter(); tot = ts = 0
byfn = collections.defaultdict(lambda: [0, 0])
PASSES = ('c1_pass','c2_pass','x1_pass','x2_pass','x3_pass','ic1_pass','ic2_pass','x2_ref_pass','stats_thread','stats_finalize','peak_stage','process_pair')
bypass = collections.defaultdict(lambda: collections.Counter())
MATH = ('FADD','FMUL','FFMA','FADD2','FMUL2','FFMA2'); MEM = ('LDS','STS','LDG','STG','LDC','LD','ST')
for (loc, txt, chain), r in zip(seq, data):
    ins = int(r[ci['Instructions Executed']] or 0); s = int(r[ci['# Samples']] or 0)
    agg[loc][0] += ins; agg[loc][1] += s; tot += ins; ts += s
    op = [o for o in txt.split() if not o.startswith('@')][0].split('.')[0]
    byop[op] += ins
    key = ('engine:' + str(fn(loc[1]))) if loc and loc[0] == 'spx_cc_engine.cuh' else (loc[0] if loc else None)
    byfn[key][0] += ins; byfn[key][1] += s
    pname = 'other'
    for f_, ln in chain:
        if f_ == 'spx_cc_engine.cuh' and fn(ln) in PASSES and fn(ln) != 'process_pair': pname = fn(ln); break
    else:
        if any(f_ == 'spx_cc.cu' for f_, _ in chain): pname = 'kernel/exec'
    bypass[pname]['inst'] += ins; bypass[pname]['smp'] += s
    bypass[pname]['math' if op in MATH else 'mem' if op in MEM else 'ovh'] += ins
print('total warp instructions %d' % tot)
print('opcodes: ' + ', '.join('%s %.1f%%' % (o, 100 * c / tot) for o, c in byop.most_common(16)))
for k, (ins, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    line = src.get(k[0], [''] * 99999)[k[1] - 1].strip()[:78] if k and k[0] in src else ''
    print('%-18s %4s %5.1f%% inst %5.1f%% smp | %s' % (k[0] if k else None, k[1] if k else '', 100 * ins / tot, 100 * s / ts, line))
print('--- by function')
for k, (ins, s) in sorted(byfn.items(), key=lambda kv: -kv[1][0])[:16]:
    print('%-28s %5.1f%% inst %5.1f%% samples' % (k, 100 * ins / tot, 100 * s / ts))
print('--- by pass (inclusive of inlined callees): inst% | samples% | math/mem/overhead share inside the pass')
for k, c in sorted(bypass.items(), key=lambda kv: -kv[1]['inst']):
    n = max(1, c['inst'])
    print('%-14s %5.1f%% inst %5.1f%% smp | math %4.1f%% mem %4.1f%% ovh %4.1f%%' % (k, 100 * c['inst'] / tot, 100 * c['smp'] / ts, 100 * c['math'] / n, 100 * c['mem'] / n, 100 * c['ovh'] / n))
