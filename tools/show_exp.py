import json, sys
for l in open(sys.argv[1] if len(sys.argv) > 1 else '/root/repo/gpurun_out/exp_sim.jsonl'):
    try: r=json.loads(l)
    except Exception: print(l.strip()); continue
    if 'us' in r: print(f"{r['variant']:14s} N={r['N']:6d} D={r['D']} C={r['C']:2d} {r['mode']:6s} {r['us']:7.2f}us graph={r.get('us_graph',-1):7.2f}us {r['gbs']:7.1f} GB/s {r['frac_6448']:.3f} err={r['max_abs_err']:.2g} bad={r['bad']} st={r['stages_env']}")
    else: print(r)
