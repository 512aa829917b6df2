"""Time BASELINE config 5: 5-fold cross-validation at N=20k, P=10 (seed 1005)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np  # noqa: E402

from bigkrls_b200 import crossvalidate_bigKRLS  # noqa: E402
import krls_oracle as o  # noqa: E402

N, P = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (20000, 10)
X, y = o.synthetic(N, P, 1005)
for rep in range(2):
    t0 = time.time()
    cv = crossvalidate_bigKRLS(y, X, seed=1, Kfolds=5)
    print(f"rep {rep}: {time.time() - t0:.3f} s; MSE_is {np.mean(cv['MSE_is']):.6f} MSE_oos {np.mean(cv['MSE_oos']):.6f}")
