"""Mutate model files (truncate / flip / insert / overwrite length fields) and feed them to the loaders: every
file must be accepted or rejected with an error, never crash.  python tools/fuzz_loader.py [seed] [iterations]"""
import numpy as np, os, shutil, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quickchem_b200 import capi, synth, xgbmodel
f = synth.random_forest_structure(3, 4, seed=2)
d = tempfile.mkdtemp()
xgbmodel.write_legacy_binary(f, d+'/m.model'); xgbmodel.write_json(f, d+'/m.json'); xgbmodel.write_ubj(f, d+'/m.ubj')
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
n_ok=n_err=0
for ext in ('model','json','ubj'):
    raw = bytearray(open(d+'/m.'+ext,'rb').read())
    for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 1500):
        b = bytearray(raw)
        mode = rng.integers(4)
        if mode == 0:
            b = b[:rng.integers(0, len(b))]
        elif mode == 1:
            for _ in range(rng.integers(1, 6)):
                b[rng.integers(len(b))] = rng.integers(256)
        elif mode == 2:
            i = rng.integers(len(b)); b[i:i] = bytes(rng.integers(0,256, rng.integers(1,9), dtype=np.uint8))
        else:
            i = rng.integers(len(b)-8); b[i:i+4] = (int(rng.integers(0, 2**31))).to_bytes(4,'little')
        p = d+'/fz.'+ext
        open(p,'wb').write(bytes(b))
        try:
            bo = capi.Booster(p, parse_only=True); bo.info(); bo.flat(); n_ok+=1
        except capi.QcohError:
            n_err+=1
shutil.rmtree(d, ignore_errors=True)
print('ok', n_ok, 'rejected', n_err)
