// Developer tool (NOT part of the product): runs the exact pass sequence of the
// CUDA cross-correlation kernel (spx::process_pair) serially on the CPU so the
// index arithmetic can be debugged in a container without a GPU.
//   g++ -std=c++17 -DSPX_HOST_EMU -O2 -x c++ -o /tmp/host_emu tools/host_emu.cu
//   /tmp/host_emu in.bin out.bin
// in.bin : int32 n, ny, nx, cc_type, no_icc (1: surface not requested -> the
//          statistics-free NCC path may run) ; float32 [n,5,ny,nx]
// out.bin: float64 [n,2] shifts ; int32 [n] peak ; int32 [n] status ; float32 [n,2ny,2nx]
#ifndef SPX_KEEP_SURFACE
#define SPX_KEEP_SURFACE 1      // the emulator exercises the on-chip surface form (build knob of the product)
#endif
#include <pthread.h>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include "../subpixal_b200/csrc/spx_cc_engine.cuh"

using namespace spx;
#ifndef EMU128_INPLACE
#define EMU128_INPLACE true
#endif

template <class C>
struct HostExec {
    // SPX_EMU_ORDER=1: reversed, =2: warps interleaved - a pass whose result depends on
    // the thread order has an intra-pass data hazard
    template <class F> void par(F f) {
        static const int order = getenv("SPX_EMU_ORDER") ? atoi(getenv("SPX_EMU_ORDER")) : 0;
        if (order == 1) { for (int t = C::T - 1; t >= 0; --t) f(t); }
        else if (order == 2) { for (int l = 0; l < 32; ++l) for (int w = C::NW - 1; w >= 0; --w) f(w * 32 + l); }
        else { for (int t = 0; t < C::T; ++t) f(t); }
    }
    template <class F> void single(F f) { f(); }
    template <class F> StatPartial reduce_stats(F f) {
        StatPartial tot = stat_zero();
        for (int t = 0; t < C::T; ++t) { StatPartial s = f(t); stat_merge(tot, s); }
        return tot;
    }
    template <class F> void parts(int n, F f) { for (int t = 0; t < n; ++t) f(t); }
    template <class F, class A> void reduce_best_into(Best *dst, F f, A after) {
        for (int t = 0; t < C::T; ++t) { Best b = f(t); if (best_better(b, *dst)) *dst = b; }
        for (int l = 0; l < 32; ++l) after(l);
    }
    template <class LF, class FF> void warp_fit(LF lane_f, FF final_f) {
        FitSums tot = fit_zero();
        for (int l = 0; l < 32; ++l) { FitSums f = lane_f(l); fit_add(tot, f); }
        final_f(tot);
    }
    // staged tiles (C::STAGE): the TMA copy is emulated by a memcpy at the moment the kernel
    // would issue it, so a tile that is overwritten too early shows up as a wrong result
    const float *next_src = nullptr; float *tiles = nullptr; size_t tile_floats = 0;
    void tiles_wait() {}
    void tiles_release() { if (next_src) for (size_t i = 0; i < tile_floats; ++i) tiles[i] = next_src[i]; }
    pthread_barrier_t *bar = nullptr;          // cluster barrier (two host threads = two CTAs)
    void cluster_sync() { pthread_barrier_wait(bar); }
    void cluster_arrive() {}                             // split phase: the wait is the barrier
    void cluster_wait() { pthread_barrier_wait(bar); }
    void global_fence() {}
    void or_status(int *p, int v) { *p |= v; }
    float load_icc(const float *p) { return *p; }
};

static bool g_no_icc = false;

template <class C>
void run(int n, int ny, int nx, int cc_type, const float *stack, double *shifts, int *peak,
         int *status, float *icc) {
    constexpr int CL = C::CL;
    std::vector<cf> W[kMaxCluster], S[kMaxCluster], R[kMaxCluster];
    std::vector<uint16_t> rrow[kMaxCluster];
    PairState ps[kMaxCluster];
    Xchg xc[kMaxCluster];
    FitWin win[kMaxCluster];
    Ctx<C> c[kMaxCluster];
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, CL);
    for (int rk = 0; rk < CL; ++rk) {
        W[rk].resize(C::W_ELEMS); S[rk].resize(C::S_ELEMS + 1); R[rk].resize(C::R_ELEMS);
        rrow[rk].resize(C::RH);
        c[rk].W = W[rk].data(); c[rk].S = S[rk].data(); c[rk].R = R[rk].data();
        c[rk].rrow = rrow[rk].data(); c[rk].ps = &ps[rk];
        c[rk].n_rrow = build_rrow<C>(rrow[rk].data(), rk);
        c[rk].rank = rk; c[rk].col_lo = rk * C::NXL;
        c[rk].xl = &xc[rk]; c[rk].win = &win[rk];
        for (int k = 0; k < 4; ++k) (&xc[rk].det[0][0])[k] = 0u;
    }
    for (int rk = 0; rk < CL; ++rk)
        for (int o = 0; o < CL; ++o) {
            c[rk].Wall[o] = W[o].data(); c[rk].Rall[o] = R[o].data(); c[rk].xall[o] = &xc[o];
        }
    const size_t tile = size_t(ny) * nx;
    std::vector<float> staged(C::STAGE ? 5 * tile : 1);
    auto body = [&](int rk) {
        HostExec<C> ex;
        ex.bar = &bar;
        if (C::STAGE) {
            ex.tiles = staged.data(); ex.tile_floats = 5 * tile;
            ex.next_src = stack; ex.tiles_release();             // the first pair
        }
        for (int p = 0; p < n; ++p) {
            c[rk].dslot = p & 1;
            c[rk].icc = icc + size_t(p) * 4 * tile;
            PairJob job;
            for (int s = 0; s < 5; ++s) job.base[s] = stack + (size_t(p) * 5 + s) * tile;
            if (C::STAGE) {
                for (int s = 0; s < 5; ++s) job.base[s] = staged.data() + s * tile;
                ex.next_src = p + 1 < n ? stack + size_t(p + 1) * 5 * tile : nullptr;
            }
            job.pitch = nx; job.H = ny; job.W = nx; job.x1 = 0; job.y1 = 0;
            job.ny = ny; job.nx = nx; job.cc_type = cc_type; job.status0 = 0;
            job.flags = 0; job.patch_icc = !g_no_icc;
            process_pair<C>(c[rk], ex, job, shifts + 2 * p, peak + p, status + p);
        }
    };
    if (CL == 1) {
        body(0);
    } else {
        std::vector<std::thread> th;
        for (int rk = 1; rk < CL; ++rk) th.emplace_back(body, rk);
        body(0);
        for (auto &t : th) t.join();
    }
    pthread_barrier_destroy(&bar);
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    int hdr[5];
    if (!f || fread(hdr, 4, 5, f) != 5) return 3;
    const int n = hdr[0], ny = hdr[1], nx = hdr[2], cc_type = hdr[3];
    g_no_icc = hdr[4] != 0;
    std::vector<float> stack(size_t(n) * 5 * ny * nx);
    if (fread(stack.data(), 4, stack.size(), f) != stack.size()) return 4;
    fclose(f);
    std::vector<double> shifts(2 * n);
    std::vector<int> peak(n), status(n);
    std::vector<float> icc(size_t(n) * 4 * ny * nx);
    const int m = ny > nx ? ny : nx;
    // 170 / 256 px classes: 4- / 8-CTA cluster forms by default, SPX_EMU_SLAB=1: global-slab forms
    const bool slab = getenv("SPX_EMU_SLAB") && atoi(getenv("SPX_EMU_SLAB")) != 0;
    // SPX_EMU_STAGE=1: the TMA-staged forms of the 32 / 64 px classes (full-size tiles only)
    const bool stage = getenv("SPX_EMU_STAGE") && atoi(getenv("SPX_EMU_STAGE")) != 0;
    if (m <= 20) run<CcCfg<32, 4, 8, 20, 64, 12, true, true, true>>(n, ny, nx, cc_type, stack.data(), shifts.data(), peak.data(), status.data(), icc.data());
    else if (m > 32 && m <= 42) run<CcCfg<64, 8, 8, 42, 128, 4, true, true, true>>(n, ny, nx, cc_type, stack.data(), shifts.data(), peak.data(), status.data(), icc.data());
    else if (m > 64 && m <= 84) run<CcCfg<128, 8, 16, 84, 512, 1, true, true, true>>(n, ny, nx, cc_type, stack.data(), shifts.data(), peak.data(), status.data(), icc.data());
    else if (m > 128 && m <= 170 && slab) run<CcCfg<256, 16, 16, 170, 256, 2, false, false>>(n, ny, nx, cc_type, stack.data(), shifts.data(), peak.data(), status.data(), icc.data());
    else if (m > 128 && m <= 170) run<CcCfg<256, 16, 16, 170, 512, 1, true, true, true, 4>>(n, ny, nx, cc_type, stack.data(), shifts.data(), peak.data(), status.data(), icc.data());
    else if (m <= 32 && stage && ny == 32 && nx == 32) run<CcCfg<48, 6, 8, 32, 128, 4, true, true, true, 1, true>>(n, ny, nx, cc_type, stack.data(), shifts.data(), peak.data(), status.data(), icc.data());
    else if (m <= 64 && m > 42 && stage && ny == 64 && nx == 64) run<CcCfg<96, 12, 8, 64, 512, 1, true, true, true, 1, true>>(n, ny, nx, cc_type, stack.data(), shifts.data(), peak.data(), status.data(), icc.data());
    else if (m <= 32) run<CcCfg<48, 6, 8, 32, 128, 4, true, true, true>>(n, ny, nx, cc_type, stack.data(), shifts.data(), peak.data(), status.data(), icc.data());
    else if (m <= 64) run<CcCfg<96, 12, 8, 64, 256, 2, true, true, true>>(n, ny, nx, cc_type, stack.data(), shifts.data(), peak.data(), status.data(), icc.data());
    else if (m <= 128) run<CcCfg<192, 12, 16, 128, 512, 1, true, true, EMU128_INPLACE, 2>>(n, ny, nx, cc_type, stack.data(), shifts.data(), peak.data(), status.data(), icc.data());
    else if (slab) run<CcCfg<384, 24, 16, 256, 256, 2, false, false>>(n, ny, nx, cc_type, stack.data(), shifts.data(), peak.data(), status.data(), icc.data());
    else run<CcCfg<384, 24, 16, 256, 384, 1, true, true, true, 8>>(n, ny, nx, cc_type, stack.data(), shifts.data(), peak.data(), status.data(), icc.data());
    f = fopen(argv[2], "wb");
    fwrite(shifts.data(), 8, shifts.size(), f);
    fwrite(peak.data(), 4, n, f);
    fwrite(status.data(), 4, n, f);
    fwrite(icc.data(), 4, icc.size(), f);
    fclose(f);
    return 0;
}
