"""CPU test of the CUDA pass sequence through the host emulator (tools/host_emu.cu: the
SAME engine source, spx_cc_engine.cuh, compiled with g++ and run serially; thread-block
clusters are emulated with one host thread per CTA and a pthread barrier).  It keeps the
index arithmetic, the cluster hand-off protocols (2 / 4 / 8 CTAs per pair, DSMEM row
split, mirrored R rows) and the statistics-free NCC path checkable without a GPU, and
re-runs the passes with permuted thread orders: a pass whose result depends on the order
has an intra-pass data hazard (the stand-in for compute-sanitizer racecheck, which is
closed on this pool).  The emulator is test infrastructure; the product never runs it."""
import os
import shutil
import struct
import subprocess
import warnings

import numpy as np
import pytest

from conftest import ROOT
from oracle import subpixal_oracle as orc

EXE = os.path.join("/tmp", "spx_host_emu_%d" % os.getuid())


@pytest.fixture(scope="module")
def emu():
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    src = os.path.join(ROOT, "tools", "host_emu.cu")
    deps = [src] + [os.path.join(ROOT, "subpixal_b200", "csrc", f)
                    for f in os.listdir(os.path.join(ROOT, "subpixal_b200", "csrc"))]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call([gxx, "-std=c++17", "-DSPX_HOST_EMU", "-O1", "-x", "c++", "-o", EXE,
                               src, "-lpthread"])

    def run(stacks, cc_type, no_icc=0, env=None):
        n, _, ny, nx = stacks.shape
        base = "/tmp/spx_emu_%d" % os.getpid()
        with open(base + ".in", "wb") as f:
            f.write(struct.pack("5i", n, ny, nx, cc_type, no_icc))
            f.write(np.ascontiguousarray(stacks, dtype=np.float32).tobytes())
        subprocess.check_call([EXE, base + ".in", base + ".out"], env=dict(os.environ, **(env or {})))
        buf = open(base + ".out", "rb").read()
        o = 0
        sh = np.frombuffer(buf, np.float64, 2 * n, o).reshape(n, 2).copy(); o += 16 * n
        pk = np.frombuffer(buf, np.int32, n, o).copy(); o += 4 * n
        st = np.frombuffer(buf, np.int32, n, o).copy(); o += 4 * n
        icc = np.frombuffer(buf, np.float32, n * 4 * ny * nx, o).reshape(n, 2 * ny, 2 * nx).copy()
        os.remove(base + ".in"); os.remove(base + ".out")
        return sh, pk, st, icc
    return run


def _oracle(stack, t):
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        return orc.find_displacement(*stack, cc_type=t, full_output=True)


# (pairs, ny, nx, masked): single-CTA classes, the 2-CTA (128), 4-CTA (170) and 8-CTA (256)
# cluster classes, full-tile fast paths and ragged shapes of each
CASES = [(2, 32, 32, False), (2, 64, 64, False), (1, 20, 37, True), (1, 5, 7, False),
         (1, 128, 128, False), (1, 100, 90, True), (1, 170, 170, False), (1, 150, 129, False),
         (1, 256, 256, False), (1, 200, 256, True)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d" % (c[1], c[2]))
def test_emulated_passes_match_the_oracle(emu, case):
    n, ny, nx, masked = case
    stacks, _ = orc.random_pairs(n, ny, nx, seed=1000 * ny + nx, masked=masked)
    for t, name in enumerate(("CC", "NCC", "ZNCC")):
        if max(ny, nx) > 128 and name == "CC":
            continue                                    # keep the CPU suite short
        sh, pk, st, icc = emu(stacks, t)
        for k in range(n):
            dx, dy, ricc, _ = _oracle(stacks[k], name)
            assert int(np.argmax(ricc)) == pk[k] and st[k] == 0
            assert abs(sh[k, 0] - dx) < 5e-5 and abs(sh[k, 1] - dy) < 5e-5
            assert np.linalg.norm(icc[k] - ricc) <= 1e-5 * np.linalg.norm(ricc)


@pytest.mark.parametrize("case", [(1, 64, 64, False), (1, 128, 128, False), (1, 170, 160, False),
                                  (1, 256, 256, False)], ids=lambda c: "%dx%d" % (c[1], c[2]))
def test_results_do_not_depend_on_the_thread_order(emu, case):
    """Reversed / warp-interleaved execution order inside every pass gives bit-identical
    results: no intra-pass hazards (in-place row stage, cluster row split)."""
    n, ny, nx, masked = case
    stacks, _ = orc.random_pairs(n, ny, nx, seed=77 + ny, masked=masked)
    ref = emu(stacks, 2)
    for order in ("1", "2"):
        got = emu(stacks, 2, env={"SPX_EMU_ORDER": order})
        assert all(np.array_equal(a, b) for a, b in zip(ref, got)), order


def test_cluster_and_slab_forms_of_the_large_classes_agree(emu):
    """170 / 256 px: the cluster forms (shared memory, DSMEM) and the global-slab forms (the
    device default: column passes fused chunk by chunk through the scratch S, two CTAs of 8
    warps per SM) run the same arithmetic in the same order - bit-identical results, for
    full and ragged shapes, with and without a requested surface, in any thread order."""
    for ny, nx, masked in ((170, 170, False), (256, 256, False), (150, 129, True), (200, 256, True),
                           (131, 33, False)):
        stacks, _ = orc.random_pairs(1, ny, nx, seed=ny, masked=masked)
        for t, no_icc in ((1, 0), (1, 1), (2, 0)):
            a = emu(stacks, t, no_icc=no_icc)
            b = emu(stacks, t, no_icc=no_icc, env={"SPX_EMU_SLAB": "1"})
            assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])), (ny, nx, t, no_icc)
            if not no_icc:
                assert np.array_equal(a[3], b[3])
        c = emu(stacks, 2, env={"SPX_EMU_SLAB": "1", "SPX_EMU_ORDER": "2"})
        assert all(np.array_equal(x, y) for x, y in zip(b, c)), (ny, nx)


def test_statistics_free_ncc_path(emu):
    """No surface requested + NCC: the statistics sweep is skipped when both standard
    deviations are provably positive; degenerate inputs fall back to the exact path."""
    for n, ny, nx in ((3, 32, 32), (2, 64, 64), (1, 128, 128), (1, 48, 40), (1, 256, 256)):
        stacks, _ = orc.random_pairs(n, ny, nx, seed=5 * ny + nx, masked=(ny == 48))
        lazy = emu(stacks, 1, no_icc=1)
        exact = emu(stacks, 1, no_icc=0)
        assert np.array_equal(lazy[1], exact[1]) and np.array_equal(lazy[2], exact[2])
        assert np.max(np.abs(lazy[0] - exact[0])) < 5e-6
    base, _ = orc.random_pairs(1, 32, 32, seed=3)
    cases = []
    c = base[0].copy(); c[1:] = 0; c[1:, 5, 5] = 3.0; cases.append(c)      # one pixel per phase
    c = base[0].copy(); c[1:] = 2.5; cases.append(c)                       # constant phases
    c = base[0].copy(); c[0] = 7.0; cases.append(c)                        # constant ref
    c = base[0].copy(); c[1:3] = 0; cases.append(c)                        # first phase pair empty
    cases.append(base[0] * np.float32(1e15))                               # huge values
    c = base[0].copy(); c[2, 4, 4] = np.nan; cases.append(c)               # NaN in a phase
    stacks = np.stack(cases)
    lazy, exact = emu(stacks, 1, no_icc=1), emu(stacks, 1, no_icc=0)
    assert np.array_equal(lazy[0], exact[0], equal_nan=True) and np.array_equal(lazy[1], exact[1])


def _decoy_case(n, decoy, rng):
    """A pair whose global peak lies on an odd surface row (second phase pair) far away
    from the best sample of the first phase pair (a decoy present in image00 / image10
    only): the fit samples of the first pair are then NOT in the window kept on chip."""
    c = (n - 1) / 2.0
    st = np.zeros((5, n, n), np.float32)
    st[0] = 0.02 * rng.standard_normal((n, n)) + orc.gaussian_star(n, c, c, 2.0, amp=10.0)
    for k, (a, b) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
        im = orc.gaussian_star(n, c - 9.0 - a / 2, c - 7.5 - b / 2, 2.0, amp=30.0)
        if b == 0:
            im = im + orc.gaussian_star(n, c + 3.0 - a / 2, c + 3.0, 2.0, amp=decoy)
        st[1 + k] = im
    return st


def test_surface_kept_in_shared_memory(emu):
    """No surface requested: IC2 works in place, the fit reads W plus the saved 7x7 window
    of the first phase pair - or repeats the first pair when the window misses."""
    rng = np.random.default_rng(1)
    for n in (32, 64, 42):
        stacks = np.stack([_decoy_case(n, 29.8, rng), _decoy_case(n, 0.0, rng)])
        for t, name in enumerate(("CC", "ZNCC", )):
            t = {"CC": 0, "ZNCC": 2}[name]
            kept = emu(stacks, t, no_icc=1)
            full = emu(stacks, t, no_icc=0)
            assert np.array_equal(kept[0], full[0]) and np.array_equal(kept[1], full[1])
            assert np.array_equal(kept[2], full[2]) and np.all(kept[2] == 0)
            for k in range(2):
                dx, dy, ricc, _ = _oracle(stacks[k], name)
                assert int(np.argmax(ricc)) == kept[1][k] and (kept[1][k] // (2 * n)) % 2 == 1
                assert abs(kept[0][k, 0] - dx) < 5e-5 and abs(kept[0][k, 1] - dy) < 5e-5


def test_staged_tile_forms_agree_with_the_plain_forms(emu):
    """TMA-staged forms of the 32 / 64 px classes (tiles of the next pair are copied into
    shared memory as soon as the last C1 pass of the current pair is over): bit-identical
    to the plain forms for every cc_type, with and without a requested surface - a tile
    overwritten too early (e.g. before a statistics-free attempt is abandoned and the pair
    restarted) would show up here."""
    for n in (32, 64):
        stacks, _ = orc.random_pairs(4, n, n, seed=9 * n)
        stacks[1, 1:] = 2.5                                    # constant phases: lazy attempt abandoned
        stacks[2, 3, 4, 4] = np.nan
        for t in range(3):
            for no_icc in (0, 1):
                plain = emu(stacks, t, no_icc=no_icc)
                staged = emu(stacks, t, no_icc=no_icc, env={"SPX_EMU_STAGE": "1"})
                assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(plain[:3], staged[:3])), (n, t)
                if not no_icc:
                    assert np.array_equal(plain[3], staged[3], equal_nan=True)
