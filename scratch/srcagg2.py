"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per source line / file / stall reason (first kernel or the given one).
usage: srcagg2.py src.csv [N] [kernel-substring]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else None
cur = kern = first = hdr = None
out = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) == 2 and r[0] == 'Function Name':
        kern = r[1]
        if first is None and (want is None or want in kern): first = kern
        continue
    if r and r[0] == 'Line No': hdr = r; continue
    if kern != first or not hdr or len(r) != len(hdr) or not r[0].isdigit(): continue
    ix = {n: i for i, n in enumerate(hdr)}
    g = lambda n: int(r[ix[n]]) if r[ix[n]].lstrip('-').isdigit() else 0
    stalls = {n[6:]: g(n) for n in hdr if n.startswith('stall_') and 'Not Issued' not in n}
    out.append((cur, int(r[0]), r[1].strip()[:90], g('Instructions Executed'), g('# Samples'), g('Thread Instructions Executed'), stalls))
tot = sum(o[3] for o in out) or 1; ts = sum(o[4] for o in out) or 1; tt = sum(o[5] for o in out)
print('kernel', first); print('warp inst', tot, 'thread inst', tt, 'lanes/inst %.2f' % (tt / tot), 'samples', ts)
bf = collections.Counter(); sf = collections.Counter(); st = collections.Counter()
for o in out:
    bf[o[0]] += o[3]; sf[o[0]] += o[4]
    for k, v in o[6].items(): st[k] += v
for k, v in bf.most_common(): print('  %-28s inst %5.1f%%  samples %5.1f%%' % (k, 100 * v / tot, 100 * sf[k] / ts))
print('stall samples:', ', '.join('%s %.1f%%' % (k, 100 * v / ts) for k, v in st.most_common(8)))
print('--- top lines by samples')
for o in sorted(out, key=lambda o: -o[4])[:N]:
    top = max(o[6].items(), key=lambda kv: kv[1])
    print('%-22s %4d inst=%5.2f%% samp=%5.2f%% lanes=%4.1f %-10s %s' % (o[0], o[1], 100 * o[3] / tot, 100 * o[4] / ts, o[5] / max(o[3], 1), top[0], o[2]))
