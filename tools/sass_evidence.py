"""profiles/r2_sass_tcgen05_tma.md: per-kernel counts of the SASS mnemonics that prove which hardware paths the shipped library
uses (cuobjdump -sass on pyipm_b200/libb200ipm.so; runs without a GPU)."""
import collections
import re
import subprocess
import sys

COLS = [('UTCIMMA', 'UTCIMMA (tcgen05.mma)'), ('LDTM', 'LDTM (tcgen05.ld)'), ('UTCBAR', 'UTCBAR (tcgen05.commit)'),
        ('UBLKCP', 'UBLKCP (cp.async.bulk)'), ('UTMALDG', 'UTMALDG (cp.async.bulk.tensor)'), ('DMMA', 'DMMA'),
        ('SYNCS', 'SYNCS (mbarrier)'), ('UCGABAR', 'UCGABAR (cluster barrier)'), ('LDGSTS', 'LDGSTS (cp.async)'),
        ('PRMT', 'PRMT')]
KEEP = ('oz_syrk_kernel', 'oz_slice_kernel', 'ldlt_mini_kernel', 'ldlt_panel_kernel', 'ldlt_tile_kernel', 'gemm_nt_sub64',
        'gemm_nt_dmma_kernel', 'ldlt_fwd256_kernel', 'ldlt_bwd256_kernel', 'ldlt_blockinv_kernel')


def main(so, out):
    txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
    names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', txt)), capture_output=True, text=True).stdout.split('\n')
    counts, cur, k = collections.OrderedDict(), None, 0
    excerpts = collections.defaultdict(list)
    for line in txt.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = names[k].replace('(int)', '').split('(')[0].replace('void ', '') if k < len(names) else m.group(1)
            k += 1
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            op = m.group(1)
            for key, _ in COLS:
                if op.startswith(key):
                    counts[cur][key] += 1
                    if key in ('UCGABAR', 'UTMALDG', 'UTCIMMA', 'LDTM') and len(excerpts[cur]) < 6:
                        excerpts[cur].append(line.strip()[:110])
    with open(out, 'w') as f:
        f.write('# SASS evidence (cuobjdump -sass pyipm_b200/libb200ipm.so, sm_100a; tools/sass_evidence.py): tensor-core / TMEM / TMA / '
                'cluster instructions per kernel\n\n')
        f.write('| kernel | ' + ' | '.join(t for _, t in COLS) + ' |\n|---|' + '---:|' * len(COLS) + '\n')
        for name, c in counts.items():
            if any(s in name for s in KEEP):
                f.write('| `%s` | ' % name + ' | '.join(str(c[k]) for k, _ in COLS) + ' |\n')
        f.write('\nThread-block clusters: `ldlt_fwd256_kernel` / `ldlt_bwd256_kernel` (8 CTAs per link, partial results exchanged through '
                'distributed shared memory) and `ldlt_mini_kernel` (4 CTAs, barrier between the loads and the in-place stores) carry the '
                'cluster-barrier instructions.\n')
        for name in ('b200::oz_syrk_kernel<64, 6>', 'b200::gemm_nt_sub64_tma_kernel', 'b200::ldlt_fwd256_kernel', 'b200::ldlt_mini_kernel'):
            if excerpts.get(name):
                f.write('\n## excerpt: `%s`\n```\n%s\n```\n' % (name, '\n'.join(excerpts[name])))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'pyipm_b200/libb200ipm.so', sys.argv[2] if len(sys.argv) > 2 else 'profiles/r2_sass_tcgen05_tma.md')
