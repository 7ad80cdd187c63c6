import sys, json
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as ol
r = ol.reference()
inp = ol.ref_inputs(355, 11303, 2, 10000, lookups=136_000_000, hm=b"large")
sd = r.grid_init_do_not_profile(inp, 1)
out = {}
for n in (34_000_000, 68_000_000, 136_000_000):
    i2 = ol.ref_inputs(355, 11303, 2, 10000, lookups=n, hm=b"large")
    v = r.run_event_based_simulation(i2, sd, 1)
    out[n] = (int(v), int(v) % 999983)
    print(n, out[n], flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'multi_gpu_checksums_raw.json'), 'w'))
