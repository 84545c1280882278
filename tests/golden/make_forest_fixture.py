"""Generates tests/golden/forest_xy.npz from the reference's shipped query data (config C1).

Run in the build container (needs /root/reference):  python tests/golden/make_forest_fixture.py
Restates, for the forest workload only, what the reference's loaders do before the hot path:
  * file order  sorted(os.listdir(query_path))  => query_10.txt first         QuerySampler.py:172-186
  * line format "col,upper,lower#...@card"                                    QuerySampler.py:157-170
  * encoding    per numeric column [upper_norm, lower_norm]*1000, default (0, 1000)
                                                                              QuerySampler.py:200-221
  * label       log2(card)                                                    QuerySampler.py:188-198
  * split       random.seed(10); random.shuffle(indices); 60/20/20            util.py:271-293
forest.csv is NOT shipped with the reference (readme.md:37), so the per-column min/max the encoder
takes from the data (QuerySampler.py:50-51) are the UCI Covertype ranges of columns A..J.  Both the
oracle and the CUDA path consume the SAME X, so parity is unaffected by this choice.
Only the encoded arrays are stored (float32-exact values are not assumed: stored as float64).
"""
import os
import random
from pathlib import Path

import numpy as np

QUERY_DIR = "/root/reference/Queries/forest_data"
COLS = "ABCDEFGHIJ"
RANGES = {"A": (1859, 3858), "B": (0, 360), "C": (0, 66), "D": (0, 1397), "E": (-173, 601),
          "F": (0, 7117), "G": (0, 254), "H": (0, 254), "I": (0, 254), "J": (0, 7173)}


def encode(line):
    preds, card = line.split("@")[0].strip(), int(line.split("@")[1].strip())
    x = np.zeros(2 * len(COLS))
    x[1::2] = 1000.0
    npred = 0
    for p in preds.split("#"):
        f = p.split(",")
        c = COLS.index(f[0].strip())
        lo_r, hi_r = RANGES[COLS[c]]
        upper, lower = float(f[1]), float(f[2])
        x[2 * c] = (upper - lo_r) / (hi_r - lo_r) * 1000
        x[2 * c + 1] = (lower - lo_r) / (hi_r - lo_r) * 1000
        npred += 1
    return x, card, npred


def main():
    xs, cards, npreds = [], [], []
    for name in sorted(os.listdir(QUERY_DIR)):
        with open(os.path.join(QUERY_DIR, name)) as fh:
            for line in fh:
                if line.strip():
                    x, card, k = encode(line)
                    xs.append(x); cards.append(card); npreds.append(k)
    X = np.array(xs)
    Y = np.log2(np.array(cards, dtype=np.float64))
    idx = list(range(len(X)))
    random.seed(10)
    random.shuffle(idx)
    X, Y, P = X[idx], Y[idx], np.array(npreds)[idx]
    ntr, nte = int(0.6 * len(X)), int(0.2 * len(X))
    np.savez_compressed(Path(__file__).with_name("forest_xy.npz"),
                        x_train=X[:ntr], y_train=Y[:ntr], x_test=X[ntr:ntr + nte], y_test=Y[ntr:ntr + nte],
                        num_predicates_test=P[ntr:ntr + nte])
    print("train", X[:ntr].shape, "test", X[ntr:ntr + nte].shape)


if __name__ == "__main__":
    main()
