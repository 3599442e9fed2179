"""Developer tool: compare tools/host_emu (CPU emulation of the CUDA pass
sequence) with the oracle.  Not used by the product; kept under tests/ because it
imports the oracle.  Build the emulator first:
    g++ -std=c++17 -DSPX_HOST_EMU -O2 -x c++ -o /tmp/host_emu tools/host_emu.cu -lpthread"""
import os, struct, subprocess, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import subpixal_oracle as orc

def run(stacks, cc_type, no_icc=0):
    n, _, ny, nx = stacks.shape
    with open('/tmp/emu_in.bin', 'wb') as f:
        f.write(struct.pack('5i', n, ny, nx, cc_type, no_icc))
        f.write(np.ascontiguousarray(stacks, dtype=np.float32).tobytes())
    subprocess.check_call(['/tmp/host_emu', '/tmp/emu_in.bin', '/tmp/emu_out.bin'])
    buf = open('/tmp/emu_out.bin', 'rb').read()
    o = 0
    sh = np.frombuffer(buf, np.float64, 2 * n, o).reshape(n, 2); o += 16 * n
    pk = np.frombuffer(buf, np.int32, n, o); o += 4 * n
    st = np.frombuffer(buf, np.int32, n, o); o += 4 * n
    icc = np.frombuffer(buf, np.float32, n * 4 * ny * nx, o).reshape(n, 2 * ny, 2 * nx)
    return sh, pk, st, icc

def check(stacks, label):
    for t, name in enumerate(('CC', 'NCC', 'ZNCC')):
        sh, pk, st, icc = run(stacks, t)
        worst = [0, 0, 0]
        for k, s in enumerate(stacks):
            with warnings.catch_warnings(), np.errstate(all='ignore'):
                warnings.simplefilter('ignore')
                dx, dy, ricc, _ = orc.find_displacement(*s, cc_type=name, full_output=True)
            am = int(np.argmax(ricc))
            l2 = np.linalg.norm(icc[k].astype(np.float64) - ricc) / max(np.linalg.norm(ricc), 1e-300)
            worst[0] = max(worst[0], abs(sh[k, 0] - dx), abs(sh[k, 1] - dy))
            worst[1] = max(worst[1], l2)
            worst[2] += int(am != pk[k])
        print('%-10s %-5s max|dshift|=%.2e  max relL2=%.2e  argmax mismatches=%d  status=%s'
              % (label, name, worst[0], worst[1], worst[2], sorted(set(st.tolist()))))

if __name__ == '__main__' and len(sys.argv) == 1:
    for (n, ny, nx) in [(4, 32, 32), (3, 64, 64), (3, 20, 37), (3, 33, 33), (3, 48, 48), (2, 5, 7), (1, 128, 128), (1, 100, 90)]:
        stacks, _ = orc.random_pairs(n, ny, nx, seed=ny * 1000 + nx, masked=(ny == 48))
        check(stacks, '%dx%d' % (ny, nx))

def degenerate():
    g = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'degenerate.npz'))
    order = ["nan_slot0", "nan_slot1", "nan_slot2", "nan_slot3", "nan_slot4", "zero_phases", "inf_slot3", "bogus_type"]
    tmap = {'CC': 0, 'NCC': 1, 'ZNCC': 2, 'bogus': 7, 'zncc': 2}
    for name, (dx, dy, am, nanfrac) in zip(g['names'], g['results']):
        case, t = str(name).split(':')
        st = g['stacks'][order.index(case)] if case != 'lower_case' else g['stacks'][-1]
        sh, pk, status, icc = run(st[None], tmap[t])
        ok = abs(sh[0, 0] - dx) < 1e-5 and abs(sh[0, 1] - dy) < 1e-5 and pk[0] == am and abs(np.mean(np.isnan(icc)) - nanfrac) < 1e-12
        print('%-18s %s got (%.4f, %.4f, %d, nan=%.2f) want (%.4f, %.4f, %d, nan=%.2f) status=0x%x'
              % (name, 'OK ' if ok else 'BAD', sh[0, 0], sh[0, 1], pk[0], np.mean(np.isnan(icc)), dx, dy, am, nanfrac, status[0]))

if len(sys.argv) > 1 and sys.argv[1] == 'deg':
    degenerate()
    stacks, _ = orc.random_pairs(1, 256, 256, seed=5)
    check(stacks, '256x256')


def lazy():
    """Statistics-free NCC path (surface not requested) against the exact path and the oracle."""
    for (n, ny, nx, masked) in [(6, 32, 32, False), (4, 64, 64, False), (3, 48, 48, True), (3, 20, 37, False),
                                (1, 128, 128, False), (2, 100, 90, False)]:
        stacks, _ = orc.random_pairs(n, ny, nx, seed=ny * 1000 + nx + 7, masked=masked)
        a = run(stacks, 1, 1)
        sh_l, pk_l, st_l = a[0].copy(), a[1].copy(), a[2].copy()
        b = run(stacks, 1, 0)
        worst = 0.0
        for k, s in enumerate(stacks):
            with warnings.catch_warnings(), np.errstate(all='ignore'):
                warnings.simplefilter('ignore')
                dx, dy = orc.find_displacement(*s, cc_type='NCC')
            worst = max(worst, abs(sh_l[k, 0] - dx), abs(sh_l[k, 1] - dy))
        print('%dx%d lazy vs exact: max|d|=%.2e peak equal=%s status %s/%s ; lazy vs oracle max|d|=%.2e'
              % (ny, nx, np.abs(sh_l - b[0]).max(), np.array_equal(pk_l, b[1]), sorted(set(st_l.tolist())),
                 sorted(set(b[2].tolist())), worst))
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'degenerate.npz'))
    order = ["nan_slot0", "nan_slot1", "nan_slot2", "nan_slot3", "nan_slot4", "zero_phases", "inf_slot3", "bogus_type"]
    for name, (dx, dy, am, nanfrac) in zip(g['names'], g['results']):
        case, t = str(name).split(':')
        if t != 'NCC':
            continue
        st = g['stacks'][order.index(case)]
        sh, pk, status, _ = run(st[None], 1, 1)
        ok = abs(sh[0, 0] - dx) < 1e-5 and abs(sh[0, 1] - dy) < 1e-5 and pk[0] == am
        print('%-18s %s lazy got (%.4f, %.4f, %d) want (%.4f, %.4f, %d) status=0x%x'
              % (name, 'OK ' if ok else 'BAD', sh[0, 0], sh[0, 1], pk[0], dx, dy, am, status[0]))
    # constant / single-pixel cases: the detector must hand over to the exact path
    base, _ = orc.random_pairs(1, 32, 32, seed=3)
    cases = {}
    c = base[0].copy(); c[1:] = 0; c[1:, 5, 5] = 3.0; cases['one pixel per phase'] = c
    c = base[0].copy(); c[1:] = 2.5; cases['constant phases'] = c
    c = base[0].copy(); c[0] = 7.0; cases['constant ref'] = c
    c = base[0].copy(); c[1:3] = 0; cases['first phase pair zero'] = c
    c = base[0].copy() * 1e-20; cases['tiny values'] = c
    for name, c in cases.items():
        a = run(c[None], 1, 1); la = (a[0].copy(), a[1].copy(), a[2].copy())
        b = run(c[None], 1, 0)
        same = np.array_equal(la[0], b[0], equal_nan=True) and la[1][0] == b[1][0] and la[2][0] == b[2][0]
        print('%-24s %s lazy (%.4f, %.4f, %d, 0x%x) exact (%.4f, %.4f, %d, 0x%x)' % (
            name, 'OK ' if same else 'BAD', la[0][0, 0], la[0][0, 1], la[1][0], la[2][0],
            b[0][0, 0], b[0][0, 1], b[1][0], b[2][0]))

if len(sys.argv) > 1 and sys.argv[1] == 'lazy':
    lazy()
