"""Developer tool: list the executed SASS of one engine pass with counters.
python tools/ncu_sass.py REPORT.ncu-rep PASSNAME [kernel-substring]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, pname = sys.argv[1], sys.argv[2]
ksub = sys.argv[3] if len(sys.argv) > 3 else 'Li96ELi12ELi8ELi64'
sass_csv = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass_csv))); hdr = rows[1]; data = [r for r in rows[2:] if len(r) >= len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'subpixal_b200/_native/libsubpixal_b200.so')], cwd=tmp, capture_output=True)
sass = subprocess.run(['nvdisasm', '-gi', '-c', os.path.join(tmp, 'spx_cc.sm_100a.cubin')], capture_output=True, text=True).stdout.split('\n')
start = end = None
for i, l in enumerate(sass):
    if l.startswith('.text.'):
        if start is not None and end is None: end = i
        if ksub in l: start, end = i, None
src = open(os.path.join(ROOT, 'subpixal_b200/csrc/spx_cc_engine.cuh')).read().split('\n')
funcs = [(i + 1, m.group(1)) for i, l in enumerate(src) for m in [re.search(r'SPX_HD [\w:<> ]+? (\w+)\(', l)] if m]
def fn(line):
    name = None
    for a, n in funcs:
        if a <= line: name = n
    return name
seq, chain, fresh = [], [], True
for l in sass[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        loc = (m.group(1).split('/')[-1], int(m.group(2)))
        if fresh: chain, fresh = [loc], False
        else: chain.append(loc)
        m2 = re.search(r'inlined at "([^"]+)", line (\d+)', l)
        if m2: chain.append((m2.group(1).split('/')[-1], int(m2.group(2))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append((m.group(2), list(chain))); fresh = True
n = 0
for (txt, chain), r in zip(seq, data):
    if any(f == 'spx_cc_engine.cuh' and fn(ln) == pname for f, ln in chain):
        ins = int(r[ci['Instructions Executed']] or 0)
        inner = chain[0]
        print('%9d %5s | %-70s | %s:%d' % (ins, r[ci['# Samples']], txt[:70], inner[0][:14], inner[1]))
        n += 1
print(n, 'instructions')
