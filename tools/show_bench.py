#!/usr/bin/env python3
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.3f} {d['unit']}  ms/step {d['ms_per_step']:.4f}  n_gpus {d['n_gpus']} launches {d.get('gpu_launches')} clocks {d.get('clocks')}")
print("roofline", {k: v for k, v in d['roofline'].items() if k in ('kernel', 'achieved', 'frac', 'kernel_ms', 'fit_frac_x_read_once')})
print("e2e", d.get('e2e'))
print("cpu", d.get('cpu_baseline'))
for k, v in d.get('matd_ops_32768x32768_f64', {}).items():
    print(f"  {k:45s} {v['ms']:9.3f} ms {v['gbs']:8.1f} GB/s  {v['frac_hbm']:.3f}")
for k, v in d.get('matmul_sweep', {}).items():
    print("  ", k, v)
if 'ops_error' in d: print(d['ops_error'])
