import json, sys
for line in open(sys.argv[1]):
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print(f"value={d['value']:.4g} {d['unit']}  ms/step={d['ms_per_step']:.4f}  sweeps/s={d.get('sweeps_per_sec', 0):.1f}  "
          f"k_sweep={r.get('avg_launch_us', 0):.1f}us k_vars={r.get('variable_kernel_avg_us', 0):.1f}us "
          f"frac={r.get('frac', 0):.3f} frac_moved={r.get('frac_moved') or 0:.3f}  e2e={e.get('value', 0):.4g}  fma={(d.get('fma_mode') or {}).get('value', 0):.4g} "
          f"(k_sweep {(d.get('fma_mode') or {}).get('k_sweep_avg_us', 0):.1f}us)  parity={(d.get('parity_n') or {}).get('status')}  clocks={d.get('clocks')}  final={d.get('final')}")
