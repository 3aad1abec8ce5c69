// Savitzky-Golay smoothing of the control sequence (host side, double precision): the filter pass the Python twin runs on
// its action sequence when `filterSeq` is set (/root/reference/scripts/src/controllers/controller_base.py:281-291:
// scipy.signal.savgol_filter(seq, 10, 9, deriv=0, delta=1.0, axis=0), default mode "interp").  Restated from the
// published algorithm (Savitzky & Golay 1964; scipy's conventions for even windows and for the edges): every output is the
// value of a least-squares polynomial of degree `polyorder` through `window` consecutive points - the window centred on
// the point in the interior (evaluated half a sample off-centre for an even window, the window starting half-1 samples
// before the point), the first / last window for the first / last half points.  Least squares by Householder QR.
#include <cmath>
#include <vector>

#include "../../include/mppi_b200.h"

namespace {

// QR of the n x m Vandermonde matrix of the abscissae xs (n >= m): Householder vectors in `v`, R in `r` (m x m, row major)
struct VanderQR {
    int n, m;
    std::vector<double> v, r, beta;
    bool build(const std::vector<double> &xs, int m_)
    {
        n = (int)xs.size();
        m = m_;
        std::vector<double> a((size_t)n * m);
        for (int i = 0; i < n; i++) {
            double pw = 1.0;
            for (int k = 0; k < m; k++) { a[(size_t)i * m + k] = pw; pw *= xs[i]; }
        }
        v.assign((size_t)n * m, 0.0);
        r.assign((size_t)m * m, 0.0);
        beta.assign(m, 0.0);
        for (int k = 0; k < m; k++) {
            double norm = 0.0;
            for (int i = k; i < n; i++) norm += a[(size_t)i * m + k] * a[(size_t)i * m + k];
            norm = std::sqrt(norm);
            if (norm == 0.0) return false;
            const double alpha = a[(size_t)k * m + k] > 0 ? -norm : norm;
            for (int i = k; i < n; i++) v[(size_t)i * m + k] = a[(size_t)i * m + k];
            v[(size_t)k * m + k] -= alpha;
            double vv = 0.0;
            for (int i = k; i < n; i++) vv += v[(size_t)i * m + k] * v[(size_t)i * m + k];
            beta[k] = vv > 0 ? 2.0 / vv : 0.0;
            for (int j = k; j < m; j++) {               // apply the reflector to the remaining columns
                double dot = 0.0;
                for (int i = k; i < n; i++) dot += v[(size_t)i * m + k] * a[(size_t)i * m + j];
                dot *= beta[k];
                for (int i = k; i < n; i++) a[(size_t)i * m + j] -= dot * v[(size_t)i * m + k];
            }
            for (int j = k; j < m; j++) r[(size_t)k * m + j] = a[(size_t)k * m + j];
        }
        return true;
    }
    void apply_qt(std::vector<double> &y) const        // y <- Q^T y
    {
        for (int k = 0; k < m; k++) {
            double dot = 0.0;
            for (int i = k; i < n; i++) dot += v[(size_t)i * m + k] * y[i];
            dot *= beta[k];
            for (int i = k; i < n; i++) y[i] -= dot * v[(size_t)i * m + k];
        }
    }
    void apply_q(std::vector<double> &y) const         // y <- Q y
    {
        for (int k = m - 1; k >= 0; k--) {
            double dot = 0.0;
            for (int i = k; i < n; i++) dot += v[(size_t)i * m + k] * y[i];
            dot *= beta[k];
            for (int i = k; i < n; i++) y[i] -= dot * v[(size_t)i * m + k];
        }
    }
    // polynomial coefficients of the least-squares fit through (xs_i, y_i)
    bool solve(std::vector<double> y, std::vector<double> &coef) const
    {
        apply_qt(y);
        coef.assign(m, 0.0);
        for (int k = m - 1; k >= 0; k--) {
            double s = y[k];
            for (int j = k + 1; j < m; j++) s -= r[(size_t)k * m + j] * coef[j];
            if (r[(size_t)k * m + k] == 0.0) return false;
            coef[k] = s / r[(size_t)k * m + k];
        }
        return true;
    }
    // weights c with  sum_i c_i y_i = value of the fitted polynomial at abscissa 0  (first row of the pseudo-inverse)
    bool weights_at_zero(std::vector<double> &c) const
    {
        std::vector<double> w(n, 0.0);                 // R^T w = e_0 (forward substitution), then c = Q [w; 0]
        for (int k = 0; k < m; k++) {
            double s = (k == 0) ? 1.0 : 0.0;
            for (int j = 0; j < k; j++) s -= r[(size_t)j * m + k] * w[j];
            if (r[(size_t)k * m + k] == 0.0) return false;
            w[k] = s / r[(size_t)k * m + k];
        }
        apply_q(w);
        c = w;
        return true;
    }
};

}  // namespace

extern "C" int mppi_savgol_filter(int T, int a, const float *U, int window, int polyorder, float *out)
{
    if (!U || !out || T <= 0 || a <= 0 || window <= 0 || window > T || polyorder < 0 || polyorder >= window || window > 256)
        return MPPI_ERR_BAD_ARG;
    const int n = window, m = polyorder + 1, half = n / 2;
    const double pos = (n & 1) ? (double)half : (double)half - 0.5;
    const int off = (n & 1) ? half : half - 1;         // the window of point t starts at t - off
    std::vector<double> xs(n);
    for (int i = 0; i < n; i++) xs[i] = (double)i - pos;
    VanderQR centre, edge;
    std::vector<double> c;
    if (!centre.build(xs, m) || !centre.weights_at_zero(c)) return MPPI_ERR_BAD_ARG;
    for (int i = 0; i < n; i++) xs[i] = (double)i;
    if (!edge.build(xs, m)) return MPPI_ERR_BAD_ARG;
    std::vector<double> y(n), coef;
    for (int j = 0; j < a; j++) {
        for (int t = 0; t < T; t++) {
            const int lo = t - off;
            if (lo < 0 || lo + n > T) continue;
            double s = 0.0;
            for (int i = 0; i < n; i++) s += c[i] * (double)U[(size_t)(lo + i) * a + j];
            out[(size_t)t * a + j] = (float)s;
        }
        for (int side = 0; side < 2; side++) {         // mode "interp": the first / last half points from one fit each
            const int w0 = side ? T - n : 0;
            for (int i = 0; i < n; i++) y[i] = (double)U[(size_t)(w0 + i) * a + j];
            if (!edge.solve(y, coef)) return MPPI_ERR_BAD_ARG;
            for (int q = 0; q < half; q++) {
                const int t = side ? T - half + q : q;
                const double x = (double)(t - w0);
                double s = 0.0;
                for (int k = m - 1; k >= 0; k--) s = s * x + coef[k];
                out[(size_t)t * a + j] = (float)s;
            }
        }
    }
    return MPPI_OK;
}
