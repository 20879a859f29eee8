// Strain limiting: reduceSuperelast, dcollid.cpp:485-596.
//
// The reference runs up to 10 sequential Gauss-Seidel sweeps over the edges of every non-rigid
// element in hseList order (3 per triangle, so an interior edge is visited twice per sweep; 1 per
// bond).  A visit reads avgVel of its two end points and, when the edge is stretched or stretching
// by more than 10 %, overwrites both with their mean -- so the result depends on the visit order.
//
// The order is kept exactly, the work is still parallel: two visits only conflict when they share a
// point, and which visits share a point is pure topology.  The host therefore builds, once per
// topology, the wavefront schedule of the whole 10-sweep sequence: visit i gets level
// 1 + max(level of the latest earlier visit of either end point).  Visits of one level touch
// disjoint points and every conflicting pair keeps its sequential order, so executing the levels
// one after the other reproduces the sequential result bit for bit.  Sweeps overlap in the
// schedule (sweep k+1 starts on the part of the mesh sweep k has left): a 1 M-triangle, 8-layer
// mesh needs ~2.7 K levels of ~11 K visits for all ten sweeps, where the sweeps run one by one
// would need ~25 K.  One persistent cooperative kernel walks the levels with a grid barrier in
// between.  The reference's early exit (`while (has_superelas && niter++ < 10)`) is kept: a sweep
// that averaged nothing leaves the state unchanged, hence all later sweeps are no-ops and the
// kernel stops at the level that completes such a sweep.
#pragma once
#include <cooperative_groups.h>
#include <vector>
#include "lbvh.cuh"
#include "narrow.cuh"

namespace clsn {

constexpr int STRAIN_MAX_SWEEPS = 10;   // max_iter, dcollid.cpp:590
constexpr double STRAIN_TOL = 0.10;     // superelasTol, dcollid.cpp:488

struct StrainResult {
    int viol[STRAIN_MAX_SWEEPS];  // edges averaged per sweep
    int sweeps;                   // sweeps the reference would have run
    int any;                      // pre-check: some edge is over the limit at all
};

__device__ __forceinline__ void load_cg3(const Vec4* p, double v[3])  // avgVel changes between levels: bypass L1
{
    const double2 a = __ldcg(reinterpret_cast<const double2*>(p));
    v[0] = a.x; v[1] = a.y;
    v[2] = __ldcg(reinterpret_cast<const double*>(p) + 2);
}

__device__ __forceinline__ double dist3(const double* p, const double* q)  // distance_between_positions
{
    double s = 0.0;
    for (int i = 0; i < 3; ++i) s += (p[i] - q[i]) * (p[i] - q[i]);
    return sqrt(s);
}

// one edge visit of reduceSuperelastOnce (dcollid.cpp:495-556); returns whether the edge was averaged.
// Everything but avgVel is constant during the sweeps and may have been fetched ahead of time.
template <bool WRITE>
__device__ __forceinline__ bool strain_visit(int2 e, double len0, const Vec4& X0, const Vec4& X1, Vec4* av, double dt)
{
    const double x0[3] = {X0.x, X0.y, X0.z}, x1[3] = {X1.x, X1.y, X1.z};
    double a0[3], a1[3], c0[3], c1[3];
    load_cg3(av + e.x, a0);
    load_cg3(av + e.y, a1);
    for (int k = 0; k < 3; ++k) { c0[k] = x0[k] + dt * a0[k]; c1[k] = x1[k] + dt * a1[k]; }
    const double len_new = dist3(c0, c1), len_old = dist3(x0, x1);
    bool fix;
    if (len_old > CLSN_ROUND_EPS && len_new > CLSN_ROUND_EPS) {
        const double strain_rate = (len_new - len_old) / len_old;
        const double strain = (len_new - len0) / len0;
        fix = fabs(strain) > STRAIN_TOL || fabs(strain_rate) > STRAIN_TOL;
    } else {
        fix = true;
    }
    if (WRITE && fix) {
        double* o0 = reinterpret_cast<double*>(av + e.x);
        double* o1 = reinterpret_cast<double*>(av + e.y);
        for (int k = 0; k < 3; ++k) {
            const double v = 0.5 * (a0[k] + a1[k]);
            __stcg(o0 + k, v);
            __stcg(o1 + k, v);
        }
    }
    return fix;
}

// every visit of one sweep against the current state, read-only: is anything over the limit?
__global__ void k_strain_check(int M, const int2* __restrict__ visits, const double* __restrict__ len0,
                               const Vec4* __restrict__ xo, Vec4* av, double dt, StrainResult* res)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    bool fix = false;
    if (v < M) {
        const int2 e = visits[v];
        fix = strain_visit<false>(e, len0[v], ldg_vec4(xo + e.x), ldg_vec4(xo + e.y), av, dt);
    }
    if (__any_sync(0xffffffffu, fix) && (threadIdx.x & 31) == 0) res->any = 1;
}

__global__ void __launch_bounds__(256) k_strain_wavefront(int nlev, const int* __restrict__ lev_off,
                                                          const unsigned* __restrict__ sched, int M,
                                                          const int2* __restrict__ visits, const double* __restrict__ len0,
                                                          const int* __restrict__ sweep_end_level,
                                                          const Vec4* __restrict__ xo, Vec4* av, double dt, StrainResult* res)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    if (res->any == 0) {  // the first sweep would average nothing: one sweep, no change (uniform exit)
        if (grid.thread_rank() == 0) res->sweeps = 1;
        return;
    }
    const int gtid = (int)grid.thread_rank(), gsize = (int)grid.size();
    int sweep_done = 0;
    // The critical path is one level = barrier + dependent loads.  Only avgVel depends on the previous
    // level, so a thread fetches the schedule entry, the edge, its rest length and x_old of its first visit
    // of the NEXT level before it enters the barrier; after the barrier one round trip (avgVel) is left.
    int lo = lev_off[0], hi = lev_off[1];
    int2 pe = make_int2(0, 0);
    double plen = 0.0;
    Vec4 pX0 = {0, 0, 0, 0}, pX1 = {0, 0, 0, 0};
    int ps = 0;
    auto prefetch = [&](int i) {
        const unsigned g = sched[i];
        ps = (int)(g / (unsigned)M);
        const int v = (int)(g - (unsigned)ps * (unsigned)M);
        pe = visits[v];
        plen = len0[v];
        pX0 = ldg_vec4(xo + pe.x);
        pX1 = ldg_vec4(xo + pe.y);
    };
    if (lo + gtid < hi) prefetch(lo + gtid);
    for (int lev = 0; lev < nlev; ++lev) {
        if (lo + gtid < hi && strain_visit<true>(pe, plen, pX0, pX1, av, dt)) atomicAdd(&res->viol[ps], 1);
        for (int i = lo + gtid + gsize; i < hi; i += gsize) {  // levels wider than the grid (rare)
            const unsigned g = sched[i];
            const int s = (int)(g / (unsigned)M), v = (int)(g - (unsigned)s * (unsigned)M);
            const int2 e = visits[v];
            if (strain_visit<true>(e, len0[v], ldg_vec4(xo + e.x), ldg_vec4(xo + e.y), av, dt)) atomicAdd(&res->viol[s], 1);
        }
        if (lev + 1 < nlev) {
            lo = hi;
            hi = lev_off[lev + 2];
            if (lo + gtid < hi) prefetch(lo + gtid);
        }
        grid.sync();
        bool stop = false;
        while (sweep_done < STRAIN_MAX_SWEEPS && sweep_end_level[sweep_done] == lev) {
            const int n = *reinterpret_cast<volatile int*>(&res->viol[sweep_done]);
            ++sweep_done;
            if (n == 0) { stop = true; break; }
        }
        if (stop || sweep_done == STRAIN_MAX_SWEEPS) break;
    }
    if (gtid == 0) res->sweeps = sweep_done;
}

struct StrainTopo {
    int M = 0, nlev = 0;
    bool built = false, have_len0 = false;
    std::vector<double> h_tri_len0, h_bond_len0;
    int2* d_visits = nullptr;
    double* d_len0 = nullptr;
    unsigned* d_sched = nullptr;
    int* d_lev_off = nullptr;
    int* d_sweep_end = nullptr;
    StrainResult* d_res = nullptr;
    StrainResult* h_res = nullptr;  // pinned

    void release_schedule()
    {
        if (d_visits) cudaFree(d_visits);
        if (d_len0) cudaFree(d_len0);
        if (d_sched) cudaFree(d_sched);
        if (d_lev_off) cudaFree(d_lev_off);
        if (d_sweep_end) cudaFree(d_sweep_end);
        d_visits = nullptr; d_len0 = nullptr; d_sched = nullptr; d_lev_off = nullptr; d_sweep_end = nullptr;
        built = false;
        M = nlev = 0;
    }
    void release()
    {
        release_schedule();
        if (d_res) cudaFree(d_res);
        if (h_res) cudaFreeHost(h_res);
        d_res = nullptr; h_res = nullptr;
    }

    // visits of one sweep in the reference's order + the wavefront schedule of all ten sweeps
    int build(int V, int T, int B, const int* tri, const int* bond, const uint8_t* vflags)
    {
        release_schedule();
        if (!d_res) {
            if (cudaMalloc((void**)&d_res, sizeof(StrainResult)) != cudaSuccess) return -1;
            if (cudaMallocHost((void**)&h_res, sizeof(StrainResult)) != cudaSuccess) return -1;
        }
        std::vector<int2> visits;
        std::vector<double> len0;
        visits.reserve(3 * (size_t)T + B);
        len0.reserve(3 * (size_t)T + B);
        for (int t = 0; t < T; ++t) {
            const int* p = tri + 3 * (size_t)t;
            if ((vflags[p[0]] | vflags[p[1]] | vflags[p[2]]) & 3) continue;  // isRigidBody(hse), dcollid.cpp:494
            for (int j = 0; j < 3; ++j) {
                visits.push_back(make_int2(p[j], p[(j + 1) % 3]));
                len0.push_back(h_tri_len0[3 * (size_t)t + j]);
            }
        }
        for (int b = 0; b < B; ++b) {
            const int* p = bond + 2 * (size_t)b;
            if ((vflags[p[0]] | vflags[p[1]]) & 3) continue;
            visits.push_back(make_int2(p[0], p[1]));
            len0.push_back(h_bond_len0[b]);
        }
        M = (int)visits.size();
        built = true;
        if (M == 0) return 0;
        if ((long long)M * STRAIN_MAX_SWEEPS >= (1ll << 32)) return -2;
        const size_t total = (size_t)M * STRAIN_MAX_SWEEPS;
        std::vector<int> last((size_t)V, 0), level(total);
        int sweep_end[STRAIN_MAX_SWEEPS];
        int maxlev = 0;
        for (int s = 0; s < STRAIN_MAX_SWEEPS; ++s) {
            int end = 0;
            for (int v = 0; v < M; ++v) {
                const int2 e = visits[v];
                const int lv = (last[e.x] > last[e.y] ? last[e.x] : last[e.y]) + 1;
                last[e.x] = last[e.y] = lv;
                level[(size_t)s * M + v] = lv - 1;
                if (lv > end) end = lv;
            }
            // levels are monotone along any chain but not along the visit index: a sweep is complete
            // once the highest level holding one of its visits has run
            sweep_end[s] = end - 1;
            if (end > maxlev) maxlev = end;
        }
        nlev = maxlev;
        std::vector<int> off((size_t)nlev + 1, 0);
        for (size_t g = 0; g < total; ++g) off[level[g] + 1]++;
        for (int l = 0; l < nlev; ++l) off[l + 1] += off[l];
        std::vector<unsigned> sched(total);
        {
            std::vector<int> fill(off.begin(), off.end() - 1);
            for (size_t g = 0; g < total; ++g) sched[fill[level[g]]++] = (unsigned)g;
        }
        // a later sweep can end on the same level as an earlier one only if it is empty: keep the ends non-decreasing
        for (int s = 1; s < STRAIN_MAX_SWEEPS; ++s)
            if (sweep_end[s] < sweep_end[s - 1]) sweep_end[s] = sweep_end[s - 1];
        if (cudaMalloc((void**)&d_visits, (size_t)M * sizeof(int2)) != cudaSuccess) return -1;
        if (cudaMalloc((void**)&d_len0, (size_t)M * sizeof(double)) != cudaSuccess) return -1;
        if (cudaMalloc((void**)&d_sched, total * sizeof(unsigned)) != cudaSuccess) return -1;
        if (cudaMalloc((void**)&d_lev_off, ((size_t)nlev + 1) * sizeof(int)) != cudaSuccess) return -1;
        if (cudaMalloc((void**)&d_sweep_end, STRAIN_MAX_SWEEPS * sizeof(int)) != cudaSuccess) return -1;
        cudaMemcpy(d_visits, visits.data(), (size_t)M * sizeof(int2), cudaMemcpyHostToDevice);
        cudaMemcpy(d_len0, len0.data(), (size_t)M * sizeof(double), cudaMemcpyHostToDevice);
        cudaMemcpy(d_sched, sched.data(), total * sizeof(unsigned), cudaMemcpyHostToDevice);
        cudaMemcpy(d_lev_off, off.data(), ((size_t)nlev + 1) * sizeof(int), cudaMemcpyHostToDevice);
        cudaMemcpy(d_sweep_end, sweep_end, sizeof(sweep_end), cudaMemcpyHostToDevice);
        return 0;
    }

    // enqueue the pre-check and the wavefront kernel; the result lands in h_res after the next stream sync
    int run(const Vec4* xo, Vec4* av, double dt, int sm_count, cudaStream_t st, long long* launches)
    {
        cudaMemsetAsync(d_res, 0, sizeof(StrainResult), st);
        if (M > 0) {
            k_strain_check<<<(M + 255) / 256, 256, 0, st>>>(M, d_visits, d_len0, xo, av, dt, d_res);
            int occ = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_strain_wavefront, 256, 0);
            if (occ < 1) return -1;
            int grid = sm_count * (occ > 2 ? 2 : occ);
            const int need = (M + 255) / 256;
            if (grid > need) grid = need;
            if (grid < 1) grid = 1;
            void* args[] = {&nlev, &d_lev_off, &d_sched, &M, &d_visits, &d_len0, &d_sweep_end, (void*)&xo, &av, &dt, &d_res};
            if (cudaLaunchCooperativeKernel((void*)k_strain_wavefront, dim3(grid), dim3(256), args, 0, st) != cudaSuccess) return -1;
            *launches += 2;
        }
        cudaMemcpyAsync(h_res, d_res, sizeof(StrainResult), cudaMemcpyDeviceToHost, st);
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
};

} // namespace clsn
