import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np
import test_density_gpu as T
from pav_b200.pavlib import density
meta, ref, tig, gold = T._load('small_rev_cluster')
res = density.density_windows([(ref, tig, meta['rev'], meta['srs'])], k=meta['k'])[0]
g = gold['KERN_REV'].to_numpy(); v = res['KERN_REV']
bad = np.flatnonzero(np.abs(v - g) > 1e-11 * np.abs(g) + 1e-211)
print('n bad', len(bad), 'rows', bad[:80].tolist())
sm = gold['STATE_MER'].to_numpy()
print('state2 rows', np.flatnonzero(sm == 2).tolist()[:5], '...', np.flatnonzero(sm == 2).tolist()[-5:], 'N', len(sm))
for j in bad[:12].tolist() + bad[-6:].tolist():
    print(j, 'gold %.17e gpu %.17e rel %.3e' % (g[j], v[j], abs(v[j] - g[j]) / g[j]), 'j%20', j % 20)
