"""Prints the interesting numbers of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
def show(name, e):
    print(f"== {name}: ms_per_step {e.get('ms_per_step'):.4f}  value {e.get('value'):.4g}  step_frac {e.get('step_frac_of_hbm_peak')}  serial {e.get('serial_ms_per_step') or (e.get('timing') if isinstance(e.get('timing'), dict) else {}).get('serial_ms_per_step')}")
    for k, v in (e.get('kernels') or {}).items():
        print(f"   {k:12s} {v['ms']*1e3:7.2f} us  frac {v.get('frac_of_hbm_peak')}  (in round {v['ms_in_round_between_event_nodes']*1e3:.2f} us)")
    r = e.get('roofline') or {}
    print(f"   roofline: {r.get('kernel')} frac {r.get('frac')}")
show(d['config']['name'], d)
for c in d.get('configs') or []:
    if 'error' in c: print('==', c['name'], 'ERROR', c['error']); continue
    show(c["name"], c)
    if c.get('loss_sweep'): print('   loss_sweep', c['loss_sweep'])
for k in ('e2e', 'api', 'parity', 'clocks'):
    if d.get(k): print(k, json.dumps(d[k])[:400])
