"""Mnemonic counts of the hot kernels' SASS: cuobjdump -sass mlff_b200/lib/libso3k.so > /tmp/sass.txt ; python profiles/sass_summary.py /tmp/sass.txt"""
import re
import sys
from collections import Counter

txt = open(sys.argv[1]).read()
parts = re.split(r'\n\s*Function : ', txt)
# the instantiations the default model (nl = 3, H = 4, F = 132) and the training sweep launch
want = [('k_filter_tcILi3E', 'k_filter_tc<3>'), ('k_edge_bwd2_tcILi3E', 'k_edge_bwd2_tc<3>'),
        ('k_alpha_aggregate_scanILb1E', 'k_alpha_aggregate_scan<true>'), ('k_qk_bwd_streamILi5E', 'k_qk_bwd_stream<5>'),
        ('k_alpha_bwdILi5ELi4ELi33E', 'k_alpha_bwd<5,4,33>'), ('k_featurise', 'k_featurise'), ('k_pairsILi0E', 'k_pairs<0> (cell list)'),
        ('k_pair_records', 'k_pair_records'), ('9k_tr_gemmE', 'k_tr_gemm'), ('12k_tr_gemm_tnE', 'k_tr_gemm_tn'),
        ('k_tr_alphaINS_4Dual', 'k_tr_alpha<Dual>'), ('k_tr_aggregateINS_4Dual', 'k_tr_aggregate<Dual>'), ('k_tr_qkbarINS_4Dual', 'k_tr_qkbar<Dual>')]
groups = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'VOTE', 'RED', 'ATOM',
          'ATOMG', 'FFMA', 'DFMA', 'MUFU', 'BAR']
out = ['SASS of the hot kernels in mlff_b200/lib/libso3k.so (sm_100a; `cuobjdump -sass`, built by mlff_b200/_lib.py): mnemonic counts.',
       'UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (the TMA unit\'s 1-D bulk copy),',
       'UTMALDG / UTMASTG = cp.async.bulk.tensor (tensor-map TMA: not used -- every tile this library moves is one contiguous block, DESIGN.md section 4),',
       'SYNCS = mbarrier, RED / ATOM(G) = global atomics.', '']
for key, label in want:
    for p in parts[1:]:
        name = p.split('\n', 1)[0].strip()
        if key not in name:
            continue
        ops = []
        for l in p.split('\n'):
            m = re.search(r'/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', l)
            if m:
                ops.append(m.group(1))
        c = Counter()
        for op in ops:
            base = op.split('.')[0]
            if base in groups:
                c[base] += 1
        ldg = [op for op in ops if op.startswith('LDG')]
        stg = [op for op in ops if op.startswith('STG')]
        out.append(f'{label}   [{name[:100]}]')
        out.append(f'  {len(ops)} instructions: ' + ', '.join(f'{g} {c[g]}' for g in groups if c[g]))
        out.append(f'  128-bit global loads {sum(".128" in o for o in ldg)} of {len(ldg)}, 128-bit global stores {sum(".128" in o for o in stg)} of {len(stg)}')
        seen = set()
        for l in p.split('\n'):
            m = re.search(r'(UTCHMMA|UBLKCP|LDTM|STTM|UTCBAR)\S*', l)
            if m and m.group(0) not in seen and len(seen) < 8:
                seen.add(m.group(0))
                out.append('    ' + re.sub(r'\s+', ' ', l.strip())[:160])
        out.append('')
        break
print('\n'.join(out))
