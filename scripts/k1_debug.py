import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-encoder_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, b2enc, b2oracle as o
R, w, h, n = [int(a) for a in sys.argv[1:5]]
rng = np.random.default_rng(1)
cur = rng.integers(0, 256, (n, h, w), dtype=np.uint8); ref = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
mv, cost, _ = b2enc.me_fullpel(cur, ref, R)
z = np.zeros((h // 2, w // 2), np.uint8)
ok = True
for i in range(n):
    c = o.OFrame(w, h).load(cur[i], z, z); r = o.OFrame(w, h).load(ref[i], z, z)
    mvo, co = o.me_fullpel(c, r, R)
    ok &= np.array_equal(co, cost[i]) and np.array_equal(mvo, mv[i])
print("R", R, w, h, n, "PARITY", ok)
