// ridge_link.cu -- host-side (serial) half of the FTLE ridge extraction: linking the per-pixel ridge
// points into ordered curves and joining curves whose end points line up.
//
// Replaces /root/reference/src/numbacs/extraction/ridges.py:
//   _link_points_stepper 321-415, _linked_ridge_pts 418-603   -> link_ridge_points
//   _endpoint_distances 606-642, _connect_endpoints 645-717, ftle_ordered_ridges 720-1054
//                                                              -> order_ridges
// The per-pixel data (r_pts, r_vec, sdd) come from the device kernel behind
// _ftle_ridge_pts_connect (tensor_kernels.cu).  Everything here is inherently sequential -- a
// greedy walk that marks points as it goes, then a greedy matching of curve end points -- so it
// runs on the host, in native code (the reference runs it under numba's nopython mode).
//
// The walk is restated as a state machine over pixels, not transliterated; what has to survive
// for identical results is spelled out where it matters:
//   * seeds are visited in ascending order of the second directional derivative (most negative
//     first); a walk looks at three of the eight neighbours, picked by the octant of the ridge
//     tangent, and moves to the one minimising  a*distance + c*angle  (a = 1/h);
//   * the neighbour test inside the walk is `sdd < 0`, NOT `sdd < -sdd_thresh` (the reference does
//     not forward sdd_thresh to its stepper), while seeds need `sdd < -sdd_thresh`;
//   * a walk stops when its best neighbour already belongs to a curve (it does not try the second
//     best);
//   * single-point curves are dropped but stay marked;
//   * end-point labels are +k for the first point of curve k and -(k + 0.1) for the last, and the
//     joined curve is traversed forwards or backwards according to the sign of the label it was
//     entered through.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace b200cs {

namespace {

constexpr double kTwoPi = 6.283185307179586;
constexpr double kPiD = 3.141592653589793;

// Python's float modulo for a positive modulus
inline double pymod(double a, double m) {
    double r = std::fmod(a, m);
    if (r < 0.0) r += m;
    return r;
}

// index of the smallest element; a NaN wins (numpy / numba argmin semantics), first one on ties
inline int argmin3(const double (&v)[3]) {
    int best = 0;
    for (int i = 0; i < 3; ++i) {
        if (std::isnan(v[i])) return i;
        if (v[i] < v[best]) best = i;
    }
    return best;
}

struct Walker {
    double *rpts;            // [npix, 3]: x, y, curve number (-1 = free)
    const double *rvec;      // [npix, 2]
    const double *sdd;       // [npix]
    long long offs[10];      // raveled offsets of the eight neighbours, counter-clockwise from (+1, -1), wrapped
    double a, c;

    // one step from pixel `ind` (point p, oriented normal n).  Returns the pixel moved to, or -1.
    long long step(long long ind, double (&p)[2], double (&n)[2], double curve) const {
        const double ang = pymod(std::atan2(n[0], -n[1]), kTwoPi);
        const long long oct = ((long long)std::floor(ang * 4.0 / kPiD + 0.5)) % 8;
        double metric[3], cand[3][2], vec[3][2];
        long long at[3];
        bool any = false;
        for (int i = 0; i < 3; ++i) {
            const long long k = ind + offs[oct + i];
            at[i] = k;
            cand[i][0] = rpts[3 * k];
            cand[i][1] = rpts[3 * k + 1];
            vec[i][0] = vec[i][1] = 0.0;
            if (sdd[k] + 0.0 < 0.0) {
                const double ddx = cand[i][0] - p[0], ddy = cand[i][1] - p[1];
                const double d = std::sqrt(ddx * ddx + ddy * ddy);
                double v0 = rvec[2 * k], v1 = rvec[2 * k + 1];
                double dot = n[0] * v0 + n[1] * v1;
                if (dot < 0.0) {
                    v0 = -v0;
                    v1 = -v1;
                    dot = -dot;
                }
                vec[i][0] = v0;
                vec[i][1] = v1;
                metric[i] = a * d + c * std::acos(dot);
                any = true;
            } else {
                metric[i] = 10000.0;
            }
        }
        if (!any) return -1;
        const int b = argmin3(metric);
        const long long k = at[b];
        if (!(rpts[3 * k + 2] + 0.5 < 0.0)) return -1;   // already on a curve: the walk ends here
        rpts[3 * k + 2] = curve;
        p[0] = cand[b][0];
        p[1] = cand[b][1];
        n[0] = vec[b][0];
        n[1] = vec[b][1];
        return k;
    }
};

}  // namespace

// -> number of curves; linked[n_pts, 2], ridge_len[n_curves, 2] = (end index, length),
//    endpoints[2 n_curves, 3], tanvecs[2 n_curves, 2]
void link_ridge_points(const double *r_pts_in, const double *r_vec, const double *sdd, long long nx, long long ny,
                       double h, double c, double sdd_thresh, std::vector<double> &linked,
                       std::vector<int32_t> &ridge_len, std::vector<double> &endpoints,
                       std::vector<double> &tanvecs) {
    const long long npix = nx * ny;
    std::vector<double> rpts(r_pts_in, r_pts_in + 3 * npix);
    // seeds: ascending sdd.  Only negative values can seed or be walked to, so only those are sorted
    // (the reference argsorts the whole array and stops at the first value that fails the seed test).
    std::vector<long long> order;
    for (long long k = 0; k < npix; ++k)
        if (sdd[k] < -sdd_thresh) order.push_back(k);
    std::stable_sort(order.begin(), order.end(), [&](long long u, long long v) { return sdd[u] < sdd[v]; });

    Walker w;
    w.rpts = rpts.data();
    w.rvec = r_vec;
    w.sdd = sdd;
    w.a = 1.0 / h;
    w.c = c;
    const int di[10] = {1, 1, 1, 0, -1, -1, -1, 0, 1, 1}, dj[10] = {-1, 0, 1, 1, 1, 0, -1, -1, -1, 0};
    for (int i = 0; i < 10; ++i) w.offs[i] = (long long)di[i] * ny + dj[i];

    linked.clear();
    ridge_len.clear();
    endpoints.clear();
    tanvecs.clear();
    long long n_curves = 0;
    for (long long seed : order) {
        if (!(rpts[3 * seed + 2] + 0.5 < 0.0)) continue;       // already on a curve
        const double label = (double)n_curves;
        rpts[3 * seed + 2] = label;
        const size_t start = linked.size() / 2;
        double p[2] = {rpts[3 * seed], rpts[3 * seed + 1]};
        double n[2] = {r_vec[2 * seed], r_vec[2 * seed + 1]};
        linked.push_back(p[0]);
        linked.push_back(p[1]);
        double e0[2] = {p[0], p[1]};
        long long at = seed;
        for (int it = 0; it < 10000; ++it) {       // one way along the tangent ...
            at = w.step(at, p, n, label);
            if (at < 0) break;
            e0[0] = p[0];
            e0[1] = p[1];
            linked.push_back(p[0]);
            linked.push_back(p[1]);
        }
        // ... that half is stored back to front, so that the curve reads end-to-end
        for (size_t lo = start, hi = linked.size() / 2; lo + 1 < hi; ++lo) {
            --hi;
            std::swap(linked[2 * lo], linked[2 * hi]);
            std::swap(linked[2 * lo + 1], linked[2 * hi + 1]);
        }
        p[0] = rpts[3 * seed];
        p[1] = rpts[3 * seed + 1];
        n[0] = -r_vec[2 * seed];
        n[1] = -r_vec[2 * seed + 1];
        double e1[2] = {p[0], p[1]};
        at = seed;
        for (int it = 0; it < 10000; ++it) {       // ... and the other way
            at = w.step(at, p, n, label);
            if (at < 0) break;
            e1[0] = p[0];
            e1[1] = p[1];
            linked.push_back(p[0]);
            linked.push_back(p[1]);
        }
        const size_t end = linked.size() / 2;
        if (end - start == 1) {                    // a lone point is not a curve (it stays marked)
            linked.resize(2 * start);
            continue;
        }
        double t0[2] = {linked[2 * (start + 1)] - e0[0], linked[2 * (start + 1) + 1] - e0[1]};
        double t1[2] = {e1[0] - linked[2 * (end - 2)], e1[1] - linked[2 * (end - 2) + 1]};
        const double n0 = std::sqrt(t0[0] * t0[0] + t0[1] * t0[1]), n1 = std::sqrt(t1[0] * t1[0] + t1[1] * t1[1]);
        ridge_len.push_back((int32_t)end);
        ridge_len.push_back((int32_t)(end - start));
        const double ep[6] = {e0[0], e0[1], label, e1[0], e1[1], -(label + 0.1)};
        endpoints.insert(endpoints.end(), ep, ep + 6);
        const double tv[4] = {t0[0] / n0, t0[1] / n0, t1[0] / n1, t1[1] / n1};
        tanvecs.insert(tanvecs.end(), tv, tv + 4);
        ++n_curves;
    }
}

namespace {

struct EndPoint {
    double x, y, label, tx, ty;
};

inline double sign_of(double v) { return std::signbit(v) ? -1.0 : 1.0; }   // copysign(1, v)

// The greedy end-point matcher.  `rem` holds the end points not yet used, always in pairs
// (first point, last point) of the same curve.
struct Joiner {
    std::vector<EndPoint> rem;
    double tol, ang;

    // From end point `e` (tangent tx, ty) look for the closest remaining end point within `tol`
    // such that both curve tangents are within `ang` of the connecting segment.  On success the
    // matched curve's two end points leave `rem`, `e` becomes its far end, and the matched label
    // is returned through `label`.
    bool hop(EndPoint &e, double &label) {
        const size_t n = rem.size();
        std::vector<double> dist(n), ux(n), uy(n);
        size_t within = 0;
        for (size_t k = 0; k < n; ++k) {
            ux[k] = rem[k].x - e.x;
            uy[k] = rem[k].y - e.y;
            dist[k] = std::sqrt(ux[k] * ux[k] + uy[k] * uy[k]);
            if (dist[k] < tol) {
                ++within;
                const double nn = std::sqrt(ux[k] * ux[k] + uy[k] * uy[k]);
                ux[k] /= nn;
                uy[k] /= nn;
            }
        }
        const double s_here = sign_of(e.label);
        for (size_t trial = 0; trial < within; ++trial) {
            const size_t k = std::min_element(dist.begin(), dist.end()) - dist.begin();
            const double s_there = sign_of(rem[k].label);
            const double a0 = std::acos(-s_here * (ux[k] * e.tx + uy[k] * e.ty));
            const double a1 = std::acos(s_there * (rem[k].tx * ux[k] + rem[k].ty * uy[k]));
            if (a0 < ang && a1 < ang) {
                label = rem[k].label;
                const long long other = (long long)k + (long long)s_there;   // its partner: +1 from a first point, -1 from a last
                e = rem[other];
                const size_t lo = std::min<size_t>(k, other);
                rem.erase(rem.begin() + lo, rem.begin() + lo + 2);
                return true;
            }
            dist[k] = 10.0 * tol;
        }
        return false;
    }

    // follow hops from `e` until none is found; labels[0] is the entry label of the starting curve
    std::vector<double> chase(EndPoint e, size_t max_hops) {
        std::vector<double> labels{-(e.label + 0.1)};
        for (size_t it = 0; it < max_hops; ++it) {
            double lab;
            if (!hop(e, lab)) break;
            labels.push_back(lab);
        }
        return labels;
    }

    bool any_within(const EndPoint &e, double &dmin) const {
        bool any = false;
        dmin = HUGE_VAL;
        for (const EndPoint &r : rem) {
            const double d = std::sqrt((r.x - e.x) * (r.x - e.x) + (r.y - e.y) * (r.y - e.y));
            if (d < tol) any = true;
            if (d < dmin) dmin = d;
        }
        return any;
    }
};

}  // namespace

// -> out_pts (concatenated curves) and offsets[n_out + 1]
void order_ridges(const std::vector<double> &linked, const std::vector<int32_t> &ridge_len,
                  const std::vector<double> &endpoints, const std::vector<double> &tanvecs, double dist_tol,
                  double ep_tan_ang, long long min_ridge_pts, std::vector<double> &out_pts,
                  std::vector<long long> &offsets) {
    const size_t n_curves = ridge_len.size() / 2;
    Joiner J;
    J.tol = dist_tol;
    J.ang = ep_tan_ang;
    for (size_t k = 0; k < 2 * n_curves; ++k)
        J.rem.push_back({endpoints[3 * k], endpoints[3 * k + 1], endpoints[3 * k + 2], tanvecs[2 * k], tanvecs[2 * k + 1]});
    out_pts.clear();
    offsets.assign(1, 0);

    auto emit = [&](const std::vector<double> &labels, bool enforce_min) {
        long long total = 0;
        for (double lab : labels) total += ridge_len[2 * (size_t)std::llround(std::fabs(lab)) + 1];
        if (enforce_min && total < min_ridge_pts) return;
        for (double lab : labels) {
            const size_t r = (size_t)std::llround(std::fabs(lab));
            const long long end = ridge_len[2 * r], len = ridge_len[2 * r + 1];
            const bool backwards = labels.size() > 1 && lab + 0.01 < 0.0;   // a single curve is emitted as stored
            for (long long q = 0; q < len; ++q) {
                const long long src = backwards ? end - 1 - q : end - len + q;
                out_pts.push_back(linked[2 * src]);
                out_pts.push_back(linked[2 * src + 1]);
            }
        }
        offsets.push_back((long long)out_pts.size() / 2);
    };

    while (!J.rem.empty()) {
        const EndPoint first = J.rem[0], last = J.rem[1];
        J.rem.erase(J.rem.begin(), J.rem.begin() + 2);
        double d0, d1;
        const bool near0 = J.any_within(first, d0), near1 = J.any_within(last, d1);
        if (near0 && near1) {
            // start from the end point with the closer candidate, then do the other end, and splice:
            // the first chain reversed (each curve traversed the other way) + the second chain
            const bool first_is_closer = !(d1 < d0);
            const std::vector<double> c0 = J.chase(first_is_closer ? first : last, n_curves);
            const std::vector<double> c1 = J.chase(first_is_closer ? last : first, n_curves);
            std::vector<double> labels;
            if (c0.size() > 1 && c1.size() > 1) {
                for (size_t k = c0.size(); k-- > 0;) labels.push_back(-(c0[k] + 0.1));
                labels.insert(labels.end(), c1.begin() + 1, c1.end());
            } else if (c0.size() > 1) {
                labels = c0;
            } else {
                labels = c1;
            }
            emit(labels, true);
        } else if (near1) {
            emit(J.chase(last, n_curves), true);
        } else if (near0) {
            const std::vector<double> labels = J.chase(first, n_curves);
            emit(labels, labels.size() > 1);        // the reference keeps a lone curve of any length here
        } else {
            emit(std::vector<double>{first.label}, true);
        }
    }
}

}  // namespace b200cs
