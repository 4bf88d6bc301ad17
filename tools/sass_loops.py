"""List the loops of a SASS dump that contain shuffles (instruction mix per loop body)."""
import re, sys
from collections import Counter
ins = []
for l in open(sys.argv[1]):
    m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)', t)
    if m and 'DIV' not in t:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            body = ins[addr[tgt]:i + 1]
            ops = [x[1].split()[1] if x[1].startswith('@') else x[1].split()[0] for x in body]
            c = Counter(o.split('.')[0] for o in ops)
            if c.get('SHFL', 0) >= 2 and c.get('VOTE', 0):
                print(hex(tgt), hex(a), 'n=%d' % len(body), ' '.join('%s:%d' % kv for kv in sorted(c.items(), key=lambda kv: -kv[1])))
