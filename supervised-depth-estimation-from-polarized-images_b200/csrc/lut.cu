// Zenith-angle lookup tables: host construction (float64) and the polcue_lut handle.
//
// Reference arithmetic being replaced (file:line relative to the reference root):
//   manydepth/normals_vec.py:13-21   theta grid, diffuse DoLP table, interp1d(rho_d, theta_d, extrapolate)
//   manydepth/normals_vec.py:27-48   specular DoLP table, argmax split, two interp1d's
// scipy's interp1d(kind='linear', fill_value='extrapolate') stably sorts the knots by x, finds the
// segment with searchsorted(side='left') clipped to [1, N-1] and evaluates that segment's line; the end
// segments serve out-of-range queries.  Here the same piecewise-linear function is re-indexed so the
// device finds the segment with one multiply instead of a search:
//   * g(rho) = sqrt(rho) below 0.5 and sqrt2 - sqrt(1-rho) above spreads the knots (which crowd
//     quadratically at rho -> 0 and at the specular peak rho -> 1) almost evenly;
//   * a uniform grid over g in [0, g(last knot)] is chosen fine enough that a cell holds at most one knot
//     (queries beyond the last knot clamp to the last cell = the extrapolation segment);
//   * a cell stores that knot (or the next one to its right) and the slopes on either side of it.
// The grid size is found by verification against the exact interpolant, not assumed.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>

#include "polcue_host.h"

namespace {

constexpr int kKnots = 1000;       // normals_vec.py:13,27
constexpr double kSqrt2 = 1.41421356237309504880;
constexpr int kMaxCells = 12288;   // per table; 3 tables must also fit the shared-memory budget
constexpr size_t kMaxBlobBytes = 200 * 1024;
constexpr double kSteepSlope = 512.0;  // end segments steeper than this are evaluated in float64 (polcue_host.h)

double g_of(double rho) {
    return rho < 0.5 ? std::sqrt(std::max(rho, 0.0)) : kSqrt2 - std::sqrt(std::max(1.0 - rho, 0.0));
}
double rho_of(double g) {
    const double half = std::sqrt(0.5);
    return g < half ? g * g : 1.0 - (kSqrt2 - g) * (kSqrt2 - g);
}

struct Knots {
    std::vector<double> x, y;
    int n() const { return (int)x.size(); }
    double slope(int seg) const {  // segment [seg, seg+1]
        return (y[seg + 1] - y[seg]) / (x[seg + 1] - x[seg]);
    }
    // scipy _call_linear on sorted knots
    double exact(double q) const {
        int hi = (int)(std::lower_bound(x.begin(), x.end(), q) - x.begin());
        hi = std::min(std::max(hi, 1), n() - 1);
        const int lo = hi - 1;
        const double w = x[hi] - x[lo];
        return ((q - x[lo]) / w) * y[hi] + ((x[hi] - q) / w) * y[lo];
    }
};

void sort_knots(const double* x, const double* y, int count, Knots& out) {
    std::vector<int> order(count);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return x[a] < x[b]; });  // mergesort, as scipy
    out.x.resize(count);
    out.y.resize(count);
    for (int i = 0; i < count; ++i) {
        out.x[i] = x[order[i]];
        out.y[i] = y[order[i]];
    }
}

void fresnel_knots(double n, Knots (&tab)[3]) {
    std::vector<double> theta(kKnots), rd(kKnots), rs(kKnots);
    const double step = (M_PI / 2) / (kKnots - 1);
    for (int i = 0; i < kKnots; ++i) {
        theta[i] = i * step;  // np.linspace(0, pi/2, 1000)
    }
    theta[kKnots - 1] = M_PI / 2;
    for (int i = 0; i < kKnots; ++i) {
        const double s = std::sin(theta[i]), c = std::cos(theta[i]);
        const double s2 = s * s, root = std::sqrt(n * n - s2);
        rd[i] = ((n - 1 / n) * (n - 1 / n) * s2) / (2 + 2 * n * n - (n + 1 / n) * (n + 1 / n) * s2 + 4 * c * root);
        rs[i] = (2 * s2 * c * root) / (n * n - s2 - n * n * s2 + 2 * s2 * s2);
    }
    int imax = 0;  // np.argmax: first maximum
    for (int i = 1; i < kKnots; ++i)
        if (rs[i] > rs[imax]) imax = i;
    sort_knots(rd.data(), theta.data(), kKnots, tab[0]);
    sort_knots(rs.data(), theta.data(), imax, tab[1]);
    sort_knots(rs.data() + imax, theta.data() + imax, kKnots - imax, tab[2]);
}

struct CellD {
    double x, y, sl, sr;
};

// Span of the cell grid in g: just past the last knot.  Queries beyond it clamp to the last cell, which holds the
// last knot and the extrapolation slope, so no cells are spent on the (often long) extrapolation range.
double grid_span(const Knots& k) { return std::min(kSqrt2, g_of(k.x.back()) * (1.0 + 1e-6) + 1e-9); }

void build_cells(const Knots& k, int cells, std::vector<CellD>& out) {
    const int n = k.n();
    std::vector<double> gk(n);
    for (int i = 0; i < n; ++i) gk[i] = g_of(k.x[i]);
    out.resize(cells);
    int j = 0;
    for (int c = 0; c < cells; ++c) {
        const double g_start = grid_span(k) * c / cells;
        while (j < n && gk[j] < g_start) ++j;      // first knot at or right of the cell start
        const int kn = std::min(j, n - 1);
        const int left = std::min(std::max(kn - 1, 0), n - 2);
        const int right = std::min(kn, n - 2);
        out[c] = {k.x[kn], k.y[kn], k.slope(left), (j >= n) ? k.slope(n - 2) : k.slope(right)};
    }
}

double eval_cell(const CellD& e, double q) {
    const double d = q - e.x;
    return e.y + d * (d <= 0 ? e.sl : e.sr);
}

// Worst deviation of the cell form from the exact interpolant, probing knots, segment midpoints, both
// sides of every cell boundary (with BOTH neighbouring cells, since the device may round g across it)
// and the extrapolation ranges.
double verify(const Knots& k, const std::vector<CellD>& cells) {
    const int m = (int)cells.size();
    const double span = grid_span(k);
    auto cell_of = [&](double q) { return std::min((int)(g_of(q) * m / span), m - 1); };
    double worst = 0;
    auto probe = [&](double q, int c) {
        c = std::min(std::max(c, 0), m - 1);
        const double ref = k.exact(q);
        worst = std::max(worst, std::fabs(eval_cell(cells[c], q) - ref) / (1.0 + std::fabs(ref)));
    };
    for (int i = 0; i < k.n(); ++i) {
        probe(k.x[i], cell_of(k.x[i]));
        if (i + 1 < k.n()) {
            const double mid = 0.5 * (k.x[i] + k.x[i + 1]);
            probe(mid, cell_of(mid));
        }
    }
    for (int c = 1; c < m; ++c) {
        const double b = rho_of(span * c / m);
        for (double rel : {-4e-7, -1e-9, 1e-9, 4e-7}) {
            const double q = b + rel * (b < 0.5 ? b : 1.0 - b);  // relative to what the device rounds
            probe(q, c - 1 + (rel > 0));
            if (std::fabs(rel) < 1e-8) probe(q, c - (rel > 0));  // the neighbour the device might pick
        }
    }
    for (double q : {-1.0, -1e-3, -1e-12, 0.0, 1e-12, 1.0, 1.0 + 1e-7, 1.2, 2.0, 10.0}) probe(q, cell_of(q));
    return worst;
}

int build_host(double n, polcue_lut** out) {
    if (!out) return POLCUE_EINVAL;
    *out = nullptr;
    if (!(n > 1.0) || !std::isfinite(n)) return POLCUE_ERANGE;  // n <= 1: tables degenerate (rho_d == 0)
    Knots tab[3];
    fresnel_knots(n, tab);
    auto* lut = new polcue_lut();
    lut->n = n;
    int total = 0;
    for (int t = 0; t < 3; ++t) {
        if (tab[t].n() < 2) {
            delete lut;
            return POLCUE_ERANGE;
        }
        for (int i = 0; i + 1 < tab[t].n(); ++i)
            if (!(tab[t].x[i + 1] > tab[t].x[i])) {  // duplicate abscissa: scipy would divide by zero
                delete lut;
                return POLCUE_ERANGE;
            }
        std::vector<CellD> cells;
        int chosen = 0;
        for (int m = 192; m <= kMaxCells; m = (int)(m * 1.08) + 1) {
            build_cells(tab[t], m, cells);
            if (verify(tab[t], cells) <= 1e-9) {
                chosen = m;
                break;
            }
        }
        if (!chosen) {
            if (getenv("POLCUE_DEBUG")) {
                build_cells(tab[t], kMaxCells, cells);
                fprintf(stderr, "polcue: n=%g table %d: no cell grid up to %d cells (residual %.3e)\n", n, t, kMaxCells,
                        verify(tab[t], cells));
            }
            delete lut;
            return POLCUE_ERANGE;
        }
        lut->cells[t] = chosen;
        lut->offset[t] = total;
        lut->scale[t] = (float)(chosen / grid_span(tab[t]));
        total += chosen + 1;   // + one spare cell (copy of the last) so a 16-byte read one past the end stays in bounds
        lut->blob.resize(total);
        for (int c = 0; c <= chosen; ++c) {
            const CellD& e = cells[c < chosen ? c : chosen - 1];
            // device form: theta = y + d * sl + max(d, 0) * (sr - sl),  d = rho - x
            lut->blob[lut->offset[t] + c] = make_float4((float)e.x, (float)e.y, (float)e.sl, (float)(e.sr - e.sl));
        }
        lut->kx[t] = tab[t].x;
        lut->ky[t] = tab[t].y;
        const int last_seg = tab[t].n() - 2;
        if (std::fabs(tab[t].slope(last_seg)) > kSteepSlope) {
            lut->steep_mask |= 1 << t;
            lut->steep_x[t] = tab[t].x[last_seg];
            lut->steep_y[t] = tab[t].y[last_seg];
            lut->steep_slope[t] = tab[t].slope(last_seg);
        }
    }
    if (lut->bytes() > kMaxBlobBytes) {
        delete lut;
        return POLCUE_ERANGE;
    }
    *out = lut;
    return POLCUE_OK;
}

}  // namespace

extern "C" {

int polcue_lut_host_build(double n, polcue_lut** out) { return build_host(n, out); }

int polcue_lut_create(double n, polcue_lut** out) {
    polcue_lut* lut = nullptr;
    int rc = build_host(n, &lut);
    if (rc != POLCUE_OK) return rc;
    cudaError_t e = cudaGetDevice(&lut->device);
    if (e == cudaSuccess) e = cudaMalloc(&lut->d_blob, lut->bytes());
    if (e == cudaSuccess) e = cudaMemcpy(lut->d_blob, lut->blob.data(), lut->bytes(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (lut->d_blob) cudaFree(lut->d_blob);
        delete lut;
        *out = nullptr;
        return (int)e;
    }
    *out = lut;
    return POLCUE_OK;
}

void polcue_lut_destroy(polcue_lut* lut) {
    if (!lut) return;
    if (lut->d_blob) cudaFree(lut->d_blob);
    delete lut;
}

int polcue_lut_set_trig(polcue_lut* lut, int mufu) {
    if (!lut) return POLCUE_EINVAL;
    lut->trig_mufu = mufu ? 1 : 0;
    return POLCUE_OK;
}

int polcue_lut_cells(const polcue_lut* lut, int table) {
    if (!lut || table < 0 || table > 2) return POLCUE_EINVAL;
    return lut->cells[table];
}

int polcue_lut_knots(const polcue_lut* lut, int table, double* x, double* y, int capacity) {
    if (!lut || table < 0 || table > 2) return POLCUE_EINVAL;
    const int n = (int)lut->kx[table].size();
    if (x && y) {
        if (capacity < n) return POLCUE_EINVAL;
        std::memcpy(x, lut->kx[table].data(), n * sizeof(double));
        std::memcpy(y, lut->ky[table].data(), n * sizeof(double));
    }
    return n;
}

int polcue_lut_steep(const polcue_lut* lut, int table, double* xys) {
    if (!lut || table < 0 || table > 2) return POLCUE_EINVAL;
    if (!((lut->steep_mask >> table) & 1)) return 0;
    if (xys) {
        xys[0] = lut->steep_x[table];
        xys[1] = lut->steep_y[table];
        xys[2] = lut->steep_slope[table];
    }
    return 1;
}

int polcue_lut_eval_host(const polcue_lut* lut, int table, const float* rho, size_t count, float* theta) {
    if (!lut || table < 0 || table > 2 || (!rho && count) || (!theta && count)) return POLCUE_EINVAL;
    const float4* cells = lut->blob.data() + lut->offset[table];
    const float scale = lut->scale[table];
    for (size_t i = 0; i < count; ++i) {  // float32 arithmetic mirroring polcue::lut_coord / lut_eval
        const float r = rho[i];
        const bool low = r < 0.5f;
        const float t = fmaxf(low ? r : 1.0f - r, 0.0f);
        float g = sqrtf(t);
        g = low ? g : 1.41421356237309504880f - g;
        const float4 e = cells[std::min((int)(g * scale), lut->cells[table] - 1)];
        const float d = r - e.x;
        theta[i] = fmaf(fmaxf(d, 0.0f), e.w, fmaf(d, e.z, e.y));
        if (((lut->steep_mask >> table) & 1) && (double)r > lut->steep_x[table])   // as polcue::steep_theta
            theta[i] = (float)(lut->steep_slope[table] * ((double)r - lut->steep_x[table]) + lut->steep_y[table]);
    }
    return POLCUE_OK;
}

}  // extern "C"
