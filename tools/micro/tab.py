import json, sys
rows=[json.loads(l) for l in open(sys.argv[1])]
best={}
for r in rows:
    if r['kernel'].startswith('library'):
        print(f"{r['shape']:8s} LIB            us={r['us']:7.2f} gbs={r['gbs']:.0f}")
    else:
        flag = '' if r['bad']==0 else ' BAD'
        print(f"{r['shape']:8s} {r['variant']:22s} C={r['C']:2d} ctas={r['ctas']:4d} maxcl={r['max_active_clusters']:3d} us={r['us']:7.2f} gbs={r['gbs']:5.0f}{flag}")
