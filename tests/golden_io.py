"""Load the fixtures written by tests/golden/make_golden.py."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = json.loads(str(z["meta"]))
        self.arr = {k: z[k] for k in z.files if k != "meta"}

    def tensor(self, key):
        """-> (ndarray, labels)"""
        return self.arr[key], list(self.meta[key]["labels"])

    def chain(self, key):
        """-> (list of (ndarray, labels), meta dict)"""
        m = self.meta[key]
        return [self.tensor("%s.%d" % (key, i)) for i in range(m["n"])], m

    def scalar(self, key):
        return self.arr[key]


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))
