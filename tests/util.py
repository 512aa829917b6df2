import csv
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def mtcars():
    rows = list(csv.reader(open(os.path.join(GOLDEN, "mtcars.csv"))))
    names = [r[0] for r in rows[1:]]
    M = np.array([[float(v) for v in r[1:]] for r in rows[1:]])
    return names, M[:, 0].copy(), np.asfortranarray(M[:, 1:])


def corolla_golden():
    g = json.load(open(os.path.join(GOLDEN, "mtcars_corolla_kernel.json")))
    g.pop("_source")
    return g


def relerr(a, b):
    """max|a-b| / max|b|  (per-field relative error)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))
