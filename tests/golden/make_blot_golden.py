"""Freeze outputs of the blot oracle (oracle/blot_oracle.py) as a regression fixture.

    python tests/golden/make_blot_golden.py

NOT a parity pin against the reference: drizzlepac (the blot's actual implementation) is
absent from this image and from /root/reference, so no reference output exists for this
row (see the oracle's header).  The file only guards the restatement itself - and with it
the CUDA kernel, which is tested bit-for-bit against the restatement - against drift.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import blot_oracle as bo            # noqa: E402

PHASES = ((0.0, 0.0), (0.5, 0.0), (0.0, 0.5), (0.5, 0.5))


def cases():
    rng = np.random.default_rng(20260)
    out = []
    for k, (h, w, ny, nx) in enumerate(((24, 30, 16, 20), (9, 7, 12, 12), (40, 40, 32, 32), (5, 33, 8, 30))):
        yy, xx = np.mgrid[0:h, 0:w]
        src = (50 * np.exp(-((xx - 0.45 * w) ** 2 + (yy - 0.55 * h) ** 2) / 12.0)
               + rng.standard_normal((h, w))).astype(np.float32)
        th = 0.03 * (k + 1)
        sc = (w + 2.0) / nx
        f = lambda x, y, th=th, sc=sc, k=k: (sc * (np.cos(th) * x - np.sin(th) * y) - 1.0 + 1e-4 * k * x * y,
                                             sc * (np.sin(th) * x + np.cos(th) * y) - 0.7 + 2e-4 * y * y)
        out.append((src, bo.fit_map(f, nx, ny), ny, nx))
    return out


def main():
    data = {}
    for k, (src, coef, ny, nx) in enumerate(cases()):
        data["src%d" % k] = src
        data["map%d" % k] = coef
        data["shape%d" % k] = np.array([ny, nx])
        for interp in ("poly5", "poly3", "linear", "nearest"):
            data["%s%d" % (interp, k)] = bo.blot_phases(src, coef, ny, nx, phases=PHASES, scale=0.75,
                                                        misval=-2.0, interp=interp)
    np.savez_compressed(os.path.join(HERE, "blot_oracle.npz"), **data)
    print("wrote blot_oracle.npz with %d arrays" % len(data))


if __name__ == "__main__":
    main()
