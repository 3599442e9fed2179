"""Developer tool: static opcode histogram of one kernel in an object file.
   python tools/sass_mix.py OBJ KERNEL_SUBSTRING"""
import collections, re, subprocess, sys
out = subprocess.run(['cuobjdump', '-sass', sys.argv[1]], capture_output=True, text=True).stdout
cur, hist = None, collections.Counter()
for l in out.split('\n'):
    m = re.search(r'Function : (\S+)', l)
    if m:
        cur = m.group(1); continue
    if cur and sys.argv[2] in cur:
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', l)
        if m: hist[m.group(1)] += 1
tot = sum(hist.values())
print(tot, ' '.join('%s:%d' % kv for kv in hist.most_common(22)))
