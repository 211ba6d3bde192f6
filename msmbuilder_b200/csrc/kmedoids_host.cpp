// kmedoids_host.cpp -- host-side k-medoids step of MiniBatchKMedoids.
//
// Stands in for msmbuilder.cluster._kmedoids.{kmedoids, contigify_ids}
// (cluster/_kmedoids.pyx:23-117 -> cluster/src/kmedoids.cc:74-309,386-401) for
// the only way the hot path calls it: npass == 0, i.e. ONE run that starts
// from the caller's assignment (minibatchkmedoids.py:116-118).
//
// Why host: m = n_clusters + batch_size points (a few hundred), O(m^2) per
// sweep, strictly sequential and data-dependent (early exits, first-wins ties).
// It sits between two GPU steps (gathered pdist -> k-medoids -> relabel); the
// condensed matrix arrives by one D2H copy of m(m-1)/2 doubles.
#include <cfloat>
#include <cstdint>
#include <vector>
#include "../../include/msmb200.h"

namespace {

// condensed (scipy pdist) offset of the unordered pair {a, b}, a != b
inline int64_t tri(int64_t a, int64_t b, int64_t n)
{
    const int64_t lo = a < b ? a : b, hi = a < b ? b : a;
    return n * lo - lo * (lo + 1) / 2 + (hi - lo - 1);
}

struct Problem {
    int64_t k, n;
    const double *dm;
    double at(int64_t a, int64_t b) const { return dm[tri(a, b, n)]; }
};

// For every cluster pick the member whose summed distance to the other members
// is smallest; earlier element wins ties; partial sums are abandoned as soon as
// they exceed the incumbent (this changes nothing but speed).
void pick_medoids(const Problem &P, const std::vector<int64_t> &member_of,
                  std::vector<int64_t> &medoid, std::vector<double> &spread)
{
    for (int64_t c = 0; c < P.k; ++c) spread[c] = DBL_MAX;
    for (int64_t e = 0; e < P.n; ++e) {
        const int64_t c = member_of[e];
        double sum = 0.0;
        for (int64_t o = 0; o < P.n; ++o) {
            if (o == e || member_of[o] != c) continue;
            sum += P.at(e, o);
            if (sum > spread[c]) break;
        }
        if (sum < spread[c]) {
            spread[c] = sum;
            medoid[c] = e;
        }
    }
}

// Reassign every element to its nearest medoid (a medoid belongs to itself);
// returns the total within-cluster distance.
double reassign(const Problem &P, const std::vector<int64_t> &medoid,
                std::vector<int64_t> &member_of)
{
    double total = 0.0;
    for (int64_t e = 0; e < P.n; ++e) {
        double nearest = DBL_MAX;
        for (int64_t c = 0; c < P.k; ++c) {
            const int64_t m = medoid[c];
            if (m == e) {
                nearest = 0.0;
                member_of[e] = c;
                break;
            }
            const double d = P.at(e, m);
            if (d < nearest) {
                nearest = d;
                member_of[e] = c;
            }
        }
        total += nearest;
    }
    return total;
}

// One descent from the assignment in member_of: iterate until the objective stops
// improving or a snapshot (taken whenever the sweep number is a multiple of a
// period that doubles after each snapshot: sweeps 0, 20, 40, 80, ...) comes back.
double descend(const Problem &P, std::vector<int64_t> &member_of, std::vector<int64_t> &medoid)
{
    std::vector<int64_t> snapshot(P.n);
    std::vector<double> spread(P.k);
    double total = DBL_MAX;
    int64_t sweep = 0, period = 10;
    for (;;) {
        const double before = total;
        if (sweep % period == 0) {
            snapshot = member_of;
            if (period < INT64_MAX / 2) period *= 2;
        }
        ++sweep;
        pick_medoids(P, member_of, medoid, spread);
        total = reassign(P, medoid, member_of);
        if (total >= before) break;
        if (snapshot == member_of) break;
    }
    return total;
}

}  // namespace

extern "C" int msmb200_kmedoids(int64_t n_clusters, int64_t n_elements,
                                const double *distmatrix, int64_t *clusterid,
                                double *error, int64_t *ifound)
{
    if (!distmatrix || !clusterid || !error || !ifound || n_clusters <= 0 || n_elements <= 0)
        return MSMB200_E_INVALID;
    if (n_elements < n_clusters) {   // more clusters than points: kmedoids.cc:141-144
        *ifound = 0;
        return MSMB200_OK;
    }
    for (int64_t e = 0; e < n_elements; ++e)
        if (clusterid[e] < 0 || clusterid[e] >= n_clusters) return MSMB200_E_INVALID;

    const Problem P{n_clusters, n_elements, distmatrix};
    std::vector<int64_t> member_of(clusterid, clusterid + n_elements);
    std::vector<int64_t> medoid(n_clusters, 0);
    const double total = descend(P, member_of, medoid);

    // Output convention: the label of a cluster is the element number of its
    // medoid.  If the solution is literally the input (cannot happen for label
    // inputs 0..k-1 unless element c is the medoid of cluster c for all
    // members), the reference reports ifound = 0 and leaves clusterid alone.
    *error = DBL_MAX;
    bool differs = false;
    for (int64_t e = 0; e < n_elements && !differs; ++e)
        differs = (member_of[e] != medoid[member_of[e]]);
    if (differs) {
        *ifound = 1;
        *error = total;
        for (int64_t e = 0; e < n_elements; ++e) clusterid[e] = medoid[member_of[e]];
    } else {
        *ifound = 0;
        for (int64_t e = 0; e < n_elements; ++e) clusterid[e] = member_of[e];
    }
    return MSMB200_OK;
}

// Random restarts (the npass > 1 branch of kmedoids.cc:160-250, used by
// cluster/kmedoids.py:92-94).  The reference draws each start from a Python
// RandomState inside C (kmedoids.cc:314-383); here the caller draws all n_pass
// starts with the same RandomState calls in the same order and hands them over
// as an (n_pass, n_elements) table, so the host loop stays free of Python.
extern "C" int msmb200_kmedoids_restarts(int64_t n_clusters, int64_t n_elements,
                                         const double *distmatrix, int64_t n_pass,
                                         const int64_t *starts, int64_t *clusterid,
                                         double *error, int64_t *ifound)
{
    if (!distmatrix || !clusterid || !error || !ifound || !starts || n_clusters <= 0 ||
        n_elements <= 0 || n_pass < 2)
        return MSMB200_E_INVALID;
    *ifound = 0;
    *error = DBL_MAX;
    if (n_elements < n_clusters) return MSMB200_OK;
    *ifound = -1;                      // kmedoids.cc:147: counts up from -1
    for (int64_t i = 0; i < n_pass * n_elements; ++i)
        if (starts[i] < 0 || starts[i] >= n_clusters) return MSMB200_E_INVALID;

    const Problem P{n_clusters, n_elements, distmatrix};
    std::vector<int64_t> member_of(n_elements), medoid(n_clusters, 0);
    for (int64_t pass = 0; pass < n_pass; ++pass) {
        member_of.assign(starts + pass * n_elements, starts + (pass + 1) * n_elements);
        const double total = descend(P, member_of, medoid);
        // keep the best solution; count how often the kept one was found again
        int64_t e = 0;
        while (e < n_elements && clusterid[e] == medoid[member_of[e]]) ++e;
        if (e == n_elements) {
            ++*ifound;
        } else if (total < *error) {
            *ifound = 1;
            *error = total;
            for (int64_t j = 0; j < n_elements; ++j) clusterid[j] = medoid[member_of[j]];
        }
    }
    return MSMB200_OK;
}

extern "C" int msmb200_contigify_ids(int64_t *ids, int64_t length, int64_t *keys,
                                     int64_t *n_keys)
{
    if (!ids || !keys || !n_keys || length < 0) return MSMB200_E_INVALID;
    // first-appearance order; the id universe is tiny (<= n_clusters distinct)
    std::vector<int64_t> seen;
    for (int64_t i = 0; i < length; ++i) {
        int64_t r = 0;
        const int64_t n = (int64_t)seen.size();
        while (r < n && seen[r] != ids[i]) ++r;
        if (r == n) seen.push_back(ids[i]);
        ids[i] = r;
    }
    for (size_t r = 0; r < seen.size(); ++r) keys[r] = seen[r];
    *n_keys = (int64_t)seen.size();
    return MSMB200_OK;
}
