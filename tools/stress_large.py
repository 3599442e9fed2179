"""Developer tool (GPU box): random ragged shapes of the 170 / 256 px slab classes (and a few
of the 84 / 128 px classes) through batch.cc_fit_ragged against the CPU oracle - shapes whose
width is not a multiple of the 32-column chunks, very narrow / very flat cutouts, masked pixels,
every cc_type, with and without the surface."""
import os, sys, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import subpixal_oracle as orc
from subpixal_b200 import batch

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
shapes = [(256, 256), (129, 129), (171, 171), (256, 3), (3, 256), (200, 31), (33, 250), (170, 129),
          (255, 255), (131, 33), (160, 97), (256, 65), (65, 256), (192, 192), (224, 160)]
shapes += [(int(rng.integers(3, 257)), int(rng.integers(129, 257))) for _ in range(10)]
shapes += [(int(rng.integers(129, 257)), int(rng.integers(3, 257))) for _ in range(10)]
shapes += [(int(rng.integers(65, 129)), int(rng.integers(65, 129))) for _ in range(4)]
stacks = []
for k, (ny, nx) in enumerate(shapes):
    st, _ = orc.random_pairs(1, ny, nx, seed=900 + k, masked=(k % 3 == 0))
    stacks.append(st[0])
flat = np.concatenate([s.ravel() for s in stacks]).astype(np.float32)
off = np.concatenate([[0], np.cumsum([s.size for s in stacks])])[:-1]
worst = 0.0
for cc_type in ("NCC", "ZNCC", "CC"):
    for want in (False, True):
        res = batch.cc_fit_ragged(torch.from_numpy(flat).cuda(), off, np.array(shapes), cc_type, want_icc=want)
        sh, pk, st = (r.cpu().numpy() for r in res[:3])
        for k, (ny, nx) in enumerate(shapes):
            with warnings.catch_warnings(), np.errstate(all="ignore"):
                warnings.simplefilter("ignore")
                dx, dy, icc, _ = orc.find_displacement(*stacks[k], cc_type=cc_type, full_output=True)
            assert int(np.argmax(icc)) == pk[k], (cc_type, want, ny, nx, int(np.argmax(icc)), pk[k])
            d = max(abs(sh[k, 0] - dx), abs(sh[k, 1] - dy))
            worst = max(worst, d)
            assert d < 5e-5, (cc_type, want, ny, nx, d)
            if want:
                io = res[4]
                got = res[3][int(io[k]):int(io[k + 1])].cpu().numpy().reshape(2 * ny, 2 * nx)
                rel = np.linalg.norm(got - icc) / np.linalg.norm(icc)
                assert rel < 1e-5, (cc_type, ny, nx, rel)
print("stress ok: %d shapes x 3 cc_types x 2, max |shift - oracle| = %.2e px" % (len(shapes), worst))
