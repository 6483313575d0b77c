import sys, time, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from quickchem_b200 import capi, synth
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 90
b = capi.Booster(os.path.join(ROOT, 'build', 'oh_booster_100x18.model'))
x = synth.quick_features(synth.raw_fields(grid))
n = x.shape[0]
m = capi.DMatrix.device(n, 27); m.upload(x); m.seal()
out = capi.DeviceArray(n)
def run(tag, **params):
    for k, v in params.items(): capi.lib().qcoh_set_param(k.encode(), str(v).encode())
    for _ in range(2): b.predict_device(m, out, exp10=False)
    capi.synchronize()
    ts = []
    for _ in range(5):
        capi.flush_l2(); capi.synchronize()
        capi.timer_start(); b.predict_device(m, out, exp10=False); ts.append(capi.timer_stop())
    r = out.get()
    print(f"{tag:28s} {min(ts):8.3f} ms  (median {sorted(ts)[2]:.3f})  {n/min(ts)/1e6:.1f} Mcells/s", flush=True)
    for k in params: capi.lib().qcoh_set_param(k.encode(), b"0" if k not in ("duo",) else b"-1")
    return r
ref = run("default (8-byte nodes)", duo=0)
grid_cfgs = ((4, 6, 0xA), (4, 6, 0xE), (3, 6, 0x6), (6, 5, 0x2A), (8, 4, 0xEE))
r = run('duo default (6 trees, 4 TEX)', duo=1)
print('   bit-exact:', np.array_equal(r.view(np.uint32), ref.view(np.uint32)))
for ilp, minb, mask in grid_cfgs:
    r = run(f"duo ilp={ilp} minb={minb} mask={mask:#x}", duo=1, ilp=ilp, minb=minb, duo_mask=mask)
    print("   bit-exact:", np.array_equal(r.view(np.uint32), ref.view(np.uint32)), "maxdiff", float(np.abs(r-ref).max()))
