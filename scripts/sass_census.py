"""Static SASS census of libse_b200.so: tensor-core / TMA / mbarrier mnemonics per kernel -> profiles/<tag>_sass_census.md.
Usage: python scripts/sass_census.py r2"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'semantic_embeddings_b200', 'libse_b200.so')
COLS = [('UTC[A-Z]*MMA', 'UTC*MMA (tcgen05.mma)'), ('LDTM', 'LDTM (tcgen05.ld)'), ('STTM', 'STTM (tcgen05.st)'),
        ('UTMALDG', 'UTMALDG (TMA load)'), ('UTMASTG', 'UTMASTG (TMA store)'), ('UTMAREDG', 'UTMAREDG (TMA reduce)'),
        ('UTCBAR', 'UTCBAR (tcgen05.commit)'), ('SYNCS', 'SYNCS (mbarrier)'), ('HMMA', 'HMMA (legacy mma.sync)'), ('FFMA', 'FFMA')]


def main(tag):
    sass = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True).stdout
    names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
    counts, cur, k = collections.OrderedDict(), None, 0
    for line in sass.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = re.sub(r'\((int|bool|unsigned char)\)', '', names[k]).split('(')[0].replace('void ', '').replace('se::', '')
            k += 1
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            op = m.group(1)
            for pat, _ in COLS:
                if re.match(pat + r'(\.|$)', op):
                    counts[cur][pat] += 1
    rows = sorted(counts.items(), key=lambda kv: (-kv[1][COLS[0][0]], -kv[1]['UTMALDG'], -kv[1]['FFMA']))
    tot = collections.Counter()
    out = ['# SASS census of semantic_embeddings_b200/libse_b200.so (round %s)\n' % tag.lstrip('r'),
           '`python scripts/sass_census.py %s`: `cuobjdump -sass libse_b200.so`, instruction mnemonics counted per kernel (static counts).' % tag,
           'UTC*MMA = `tcgen05.mma`, LDTM/STTM = `tcgen05.ld/st`, UTMALDG/UTMASTG/UTMAREDG = `cp.async.bulk.tensor` load / store / reduce, '
           'UTCBAR = `tcgen05.commit`, SYNCS = mbarrier ops.  No HMMA (legacy `mma.sync`) anywhere.\n',
           '| kernel | ' + ' | '.join(t for _, t in COLS) + ' |', '|---|' + '---:|' * len(COLS)]
    for name, c in rows:
        if not any(c.values()):
            continue
        out.append('| `%s` | ' % name + ' | '.join(str(c[p]) for p, _ in COLS) + ' |')
        tot.update(c)
    out.append('| **total** | ' + ' | '.join('**%d**' % tot[p] for p, _ in COLS) + ' |')
    path = os.path.join(ROOT, 'profiles', '%s_sass_census.md' % tag)
    open(path, 'w').write('\n'.join(out) + '\n')
    print(path, dict(tot))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'r2')
