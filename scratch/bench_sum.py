#!/usr/bin/env python
"""usage: bench_sum.py bench.log -- short summary of a bench.py JSON line"""
import json, sys
d = None
for ln in open(sys.argv[1]):
    if ln.startswith('{'):
        d = json.loads(ln)
print("value %.0f  e2e %.0f  ms/step %.2f  launches %s" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']))
r = d['roofline']; print("roofline:", r['kernel'], "%.0f GB/s frac %.3f traffic %s" % (r['achieved'], r['frac'], r['traffic']))
for k, v in d['kernels'].items():
    print(f"  {k:20s} n={v['launches']:4d} avg={v['avg_ms']*1e3:7.1f}us share={v['share_of_step']*100:5.1f}% gbs={(v['gbs'] or 0):7.0f}")
for k, v in d['methods'].items():
    print(f"  {k:14s} it={v['iterations']:3d} ms={v['ms']:7.3f} Mpx.it/s={v['mpix_iter_s']:9.0f} frac={v['frac_of_hbm_peak']:.3f}")
for k, v in d.get('methods_f64', {}).items():
    print(f"  f64 {k:14s} it={v['iterations']:3d} ms={v['ms']:7.3f} Mpx.it/s={v['mpix_iter_s']:9.0f} frac={v['frac_of_hbm_peak']:.3f}")
b = d.get('batch') or {}
for k, v in b.items():
    if k != 'config': print("  batch", k, {kk: (round(vv['pairs_per_s'], 1) if isinstance(vv, dict) else vv) for kk, vv in v.items()})
