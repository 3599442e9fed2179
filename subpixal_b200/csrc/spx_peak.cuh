// Sub-pixel peak fit: device-side restatement of centroid.find_peak
// (reference centroid.py:158-236) for one surface, executed by ONE thread in
// float64 (the reference upcasts to float64 at centroid.py:98 and fits with
// numpy.linalg.lstsq at centroid.py:207).
//
// The least-squares problem is solved through the 6x6 normal equations in
// coordinates centred on the fit box.  The space of quadratic polynomials is
// translation invariant, so this is the same LS solution as the reference's
// absolute-coordinate lstsq, only better conditioned.
#pragma once
#include "spx_common.cuh"
#include "../../include/subpixal_b200.h"

namespace spx {

struct PeakOut { double x, y; int status; };

// Solve A c = b for symmetric positive definite 6x6 A (Cholesky, in place).
// Returns false when A is not numerically positive definite.
SPX_HD bool chol6(double (&A)[6][6], double (&b)[6]) {
    for (int j = 0; j < 6; ++j) {
        double d = A[j][j];
        for (int k = 0; k < j; ++k) d -= A[j][k] * A[j][k];
        if (!(d > 1e-12 * fabs(A[j][j]) + 1e-300)) return false;
        d = sqrt(d);
        A[j][j] = d;
        for (int i = j + 1; i < 6; ++i) {
            double s = A[i][j];
            for (int k = 0; k < j; ++k) s -= A[i][k] * A[j][k];
            A[i][j] = s / d;
        }
    }
    for (int i = 0; i < 6; ++i) {          // L y = b
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= A[i][k] * b[k];
        b[i] = s / A[i][i];
    }
    for (int i = 5; i >= 0; --i) {         // L^T c = y
        double s = b[i];
        for (int k = i + 1; k < 6; ++k) s -= A[k][i] * b[k];
        b[i] = s / A[i][i];
    }
    return true;
}

// at(j, i) -> double value, good(j, i) -> bool (true = use point).
// (imax, jmax): integer pre-fit pixel; (cx, cy): the `coord` the reference
// would return on its fall-back paths.
template <class At, class Good>
SPX_HD PeakOut fit_peak(At at, Good good, int ny, int nx, int imax, int jmax,
                        double cx, double cy, int wx, int wy) {
    PeakOut o; o.x = cx; o.y = cy; o.status = 0;
    if (wx * wy < 6) { o.status = SPX_S_FEW_POINTS; return o; }    // :160-162
    int x1 = imax - wx / 2; if (x1 < 0) x1 = 0;                    // :165-168
    int x2 = x1 + wx; if (x2 > nx) x2 = nx;
    int y1 = jmax - wy / 2; if (y1 < 0) y1 = 0;
    int y2 = y1 + wy; if (y2 > ny) y2 = ny;
    if (imax == x1 || imax == x2 || jmax == y1 || jmax == y2) {    // :171-172
        o.x = double(imax); o.y = double(jmax);
        o.status = SPX_S_PEAK_EDGE; return o;
    }
    if (x2 - x1 < wx) {                                            // :175-184
        if (x1 == 0) { x2 = x1 + wx; if (x2 > nx) x2 = nx; }
        if (x2 == nx) { x1 = x2 - wx; if (x1 < 0) x1 = 0; }
    }
    if (y2 - y1 < wy) {
        if (y1 == 0) { y2 = y1 + wy; if (y2 > ny) y2 = ny; }
        if (y2 == ny) { y1 = y2 - wy; if (y1 < 0) y1 = 0; }
    }
    if ((x2 - x1) * (y2 - y1) < 6) { o.status = SPX_S_FEW_POINTS; return o; }

    const double ux = 0.5 * double(x1 + x2 - 1), vy = 0.5 * double(y1 + y2 - 1);
    double A[6][6], b[6];
    for (int r = 0; r < 6; ++r) { b[r] = 0.0; for (int c = 0; c < 6; ++c) A[r][c] = 0.0; }
    int npts = 0;
    for (int j = y1; j < y2; ++j) {
        for (int i = x1; i < x2; ++i) {
            if (!good(j, i)) continue;
            const double u = double(i) - ux, v = double(j) - vy;
            const double phi[6] = {1.0, u, v, u * v, u * u, v * v}; // :194
            const double d = at(j, i);
            for (int r = 0; r < 6; ++r) {
                b[r] += phi[r] * d;
                for (int c = 0; c <= r; ++c) A[r][c] += phi[r] * phi[c];
            }
            ++npts;
        }
    }
    if (npts < 6) { o.status = SPX_S_FEW_POINTS; return o; }       // :198-204
    if (!chol6(A, b)) { o.status = SPX_S_SINGULAR; return o; }
    const double c10 = b[1], c01 = b[2], c11 = b[3], c20 = b[4], c02 = b[5];
    const double det = 4.0 * c02 * c20 - c11 * c11;                // :216
    if (det <= 0.0 || ((c20 > 0.0 && c02 >= 0.0) || (c20 >= 0.0 && c02 > 0.0))) {
        o.x = 0.5 * double(x1 + x2); o.y = 0.5 * double(y1 + y2);  // :217-225
        o.status = SPX_S_NO_MAXIMUM; return o;
    }
    const double xm = (c01 * c11 - 2.0 * c02 * c10) / det + ux;    // :227-228
    const double ym = (c10 * c11 - 2.0 * c01 * c20) / det + vy;
    if (xm > 0.0 && xm < double(nx) - 1.0 && ym > 0.0 && ym < double(ny) - 1.0) {
        o.x = xm; o.y = ym; return o;                              // :230-231
    }
    // NaN coefficients end up here as well (all comparisons false)
    o.status = (xm != xm || ym != ym) ? SPX_S_FIT_NONFINITE : SPX_S_FIT_OUTSIDE;
    return o;
}

// (value, flat index) ordering of numpy.argmax: largest value, first index;
// NaN beats every number (numpy propagates the first NaN).
struct Best { float v; int idx; };
SPX_HD Best best_init() { Best b; b.v = -INFINITY; b.idx = 0x7fffffff; return b; }
SPX_HD bool best_better(const Best &a, const Best &b) {
    const bool an = a.v != a.v, bn = b.v != b.v;
    if (an || bn) return an && (!bn || a.idx < b.idx);
    return a.v > b.v || (a.v == b.v && a.idx < b.idx);
}

}  // namespace spx
