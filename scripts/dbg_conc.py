import sys, threading
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from kpal_b200 import _cabi
from oracle import c_oracle
def synthetic_profiles(seed, n, k, lam_lo=0.3, lam_hi=8.0):
    rng = np.random.default_rng(seed)
    lam = np.exp(rng.uniform(np.log(lam_lo), np.log(lam_hi), n))
    return np.stack([rng.poisson(l, 4 ** k) for l in lam]).astype(np.int64)
sets = [synthetic_profiles(seed, 150, 7) for seed in (1, 2, 3, 4)]
want = [_cabi.distance_matrix(p, do_scale=True) for p in sets]
for i in range(4):
    o = c_oracle.distance(sets[i][0], sets[i][1], do_scale=True)
    p = _cabi.pair_distance(sets[i][0], sets[i][1], do_scale=True)
    print('seq', i, 'oracle', o, 'matrix', want[i][1,0], 'pair', p)
res = {}
def work(i):
    for r in range(5):
        m = _cabi.distance_matrix(sets[i], do_scale=True)
        p = _cabi.pair_distance(sets[i][0], sets[i][1], do_scale=True)
        res[(i, r)] = (float(np.max(np.abs(m - want[i]) / np.maximum(want[i], 1e-300))), p)
ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
[t.start() for t in ts]; [t.join() for t in ts]
for key in sorted(res): print('conc', key, res[key], 'matrix10', want[key[0]][1,0])
