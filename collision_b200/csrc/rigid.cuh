// Movable rigid bodies: re-rigidification after every pass.
//
// Reference: updateImpactZoneVelocityForRG (dcollid.cpp:267-288) calls updateImpactListVelocity
// (dcollid3d.cpp:70-200) on the union-find list of every movable body, built once per assembly by
// createImpZoneForRG (dcollid3d.cpp:54-68) with the weighted union of dcollid.cpp:1039-1059.
// The list order is pure topology, so the host restates the union-find once per topology and
// uploads each list as a flat index array; the sums over a list (centre of mass, momentum,
// inertia) are then taken sequentially in that order -- one thread per body -- because their
// floating-point summation order is part of the result.  The per-point update is one thread
// per point.  sin/cos are the correctly rounded ones (crmath.cuh), like the coplanarity cubic.
#pragma once
#include <vector>
#include "lbvh.cuh"
#include "narrow.cuh"

namespace clsn {

struct RigidBodyState {
    double x_cm[3], v_cm[3], w[3];
    double mag_w, sin_over, cos_t;  // sin(dt*|w|)/|w|, cos(dt*|w|)
    int num;
    int pad;
};

__device__ __forceinline__ double det3(const double a[3][3])  // myDet3d, dcollid.cpp:977-981
{
    return a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2]) - a[0][1] * (a[1][0] * a[2][2] - a[2][0] * a[1][2]) +
           a[0][2] * (a[1][0] * a[2][1] - a[2][0] * a[1][1]);
}

// dcollid3d.cpp:80-138: sequential sums over one list, in list order
__global__ void k_rigid_sums(int nlists, const int* __restrict__ list_offs, const int* __restrict__ list_pts,
                             const Vec4* __restrict__ xo, const Vec4* __restrict__ av, double m, double dt,
                             RigidBodyState* __restrict__ out)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nlists) return;
    const int beg = list_offs[b], end = list_offs[b + 1];
    const int num = end - beg;
    double x_cm[3] = {0, 0, 0}, v_cm[3] = {0, 0, 0};
    for (int t = beg; t < end; ++t) {
        const int p = list_pts[t];
        const Vec4 x = ldg_vec4(xo + p), v = av[p];
        x_cm[0] += x.x; x_cm[1] += x.y; x_cm[2] += x.z;
        v_cm[0] += v.x; v_cm[1] += v.y; v_cm[2] += v.z;
    }
    for (int i = 0; i < 3; ++i) { x_cm[i] /= num; v_cm[i] /= num; }
    double L[3] = {0, 0, 0}, I[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int t = beg; t < end; ++t) {
        const int p = list_pts[t];
        const Vec4 x = ldg_vec4(xo + p), v = av[p];
        const double dx[3] = {x.x - x_cm[0], x.y - x_cm[1], x.z - x_cm[2]};
        const double dv[3] = {v.x - v_cm[0], v.y - v_cm[1], v.z - v_cm[2]};
        double Li[3];
        cross3(dx, dv, Li);
        for (int i = 0; i < 3; ++i) L[i] = m * Li[i] + L[i];
        const double mag_dx = mag3(dx);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double tmp = -dx[i] * dx[j];
                if (i == j) tmp += mag_dx * mag_dx;
                I[i][j] += tmp * m;
            }
    }
    double w[3];
    const double detI = det3(I);
    for (int i = 0; i < 3; ++i) {
        double tmp[3][3];
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) tmp[r][s] = I[r][s];
        for (int j = 0; j < 3; ++j) tmp[j][i] = L[j];
        if (detI < CLSN_ROUND_EPS) w[i] = 0.0;
        else w[i] = det3(tmp) / detI;
    }
    RigidBodyState s;
    for (int i = 0; i < 3; ++i) { s.x_cm[i] = x_cm[i]; s.v_cm[i] = v_cm[i]; s.w[i] = w[i]; }
    s.mag_w = mag3(w);
    s.sin_over = 0.0;
    s.cos_t = 1.0;
    if (!(s.mag_w < CLSN_ROUND_EPS)) {
        s.sin_over = crm::sin_cr(dt * s.mag_w) / s.mag_w;
        s.cos_t = crm::cos_cr(dt * s.mag_w);
    }
    s.num = num;
    s.pad = 0;
    out[b] = s;
}

// dcollid3d.cpp:140-198: avgVel of every non-static point of the list from the rigid motion
__global__ void k_rigid_apply(int npts, const int* __restrict__ list_pts, const int* __restrict__ pt_list,
                              const RigidBodyState* __restrict__ st, const Vec4* __restrict__ xo, Vec4* av,
                              const uint8_t* __restrict__ vflags, double dt, unsigned long long* counters,
                              uint8_t* dirty, const unsigned long long* __restrict__ gate)
{
    if (gate && *gate == 0ull) return;   // the pass this belongs to did not run (see enqueue_detect in clsn.cu)
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npts) return;
    const int p = list_pts[t];
    if (vflags[p] & 1) return;
    if (dirty) dirty[p] = 1;  // impact zones: avgVel changes outside the impulse reduction
    const RigidBodyState s = st[pt_list[t]];
    const Vec4 x = ldg_vec4(xo + p);
    const double xs[3] = {x.x, x.y, x.z};
    double dx[3], xF[3], xR[3], wxR[3];
    for (int i = 0; i < 3; ++i) dx[i] = xs[i] - s.x_cm[i];
    if (s.mag_w < CLSN_ROUND_EPS) {
        for (int i = 0; i < 3; ++i) { xF[i] = dx[i]; wxR[i] = 0.0; }
        sub3(dx, xF, xR);
    } else {
        const double f = dot3(dx, s.w) / dot3(s.w, s.w);
        for (int i = 0; i < 3; ++i) xF[i] = f * s.w[i];
        sub3(dx, xF, xR);
        double tmpV[3];
        for (int i = 0; i < 3; ++i) tmpV[i] = s.sin_over * s.w[i];
        cross3(tmpV, xR, wxR);
    }
    double v[3];
    for (int i = 0; i < 3; ++i) {
        const double x_new = s.x_cm[i] + dt * s.v_cm[i] + xF[i] + s.cos_t * xR[i] + wxR[i];
        v[i] = (x_new - xs[i]) / dt;
        if (isnan(v[i])) atomicAdd(&counters[CTR_ERROR], 1ull);
    }
    Vec4 o = av[p];
    o.x = v[0]; o.y = v[1]; o.z = v[2];
    av[p] = o;
}

// The reference's union-find with per-set point lists (dcollid.cpp:995-1059): weighted union, the
// lighter list is appended to the heavier one (ties: X's list goes behind Y's).  The list order
// decides the summation order of updateImpactListVelocity, so it is restated exactly.
struct HostUF {
    std::vector<int> root, next, tail, weight;

    void reset(int V)  // makeSet, dcollid.cpp:1015-1031
    {
        root.resize(V); tail.resize(V);
        next.assign(V, -1); weight.assign(V, 1);
        for (int v = 0; v < V; ++v) root[v] = tail[v] = v;
    }
    int find(int p)  // findSet :1033-1037 (iterative; compression does not change any list)
    {
        int r = p;
        while (root[r] != r) r = root[r];
        while (root[p] != r) { int n = root[p]; root[p] = r; p = n; }
        return r;
    }
    bool merge(int X, int Y)  // mergePoint :1039-1059
    {
        int PX = find(X), PY = find(Y);
        if (PX == PY) return false;
        if (weight[PX] > weight[PY]) {
            weight[PX] += weight[PY]; root[PY] = PX; next[tail[PX]] = PY; tail[PX] = tail[PY];
        } else {
            weight[PY] += weight[PX]; root[PX] = PY; next[tail[PY]] = PX; tail[PY] = tail[PX];
        }
        return true;
    }
    // createImpZoneForRG, dcollid3d.cpp:54-68: every triangle of a movable body, first = YES
    bool merge_movable_bodies(int T, const int* tri, const int* tri_surf, const uint8_t* vflags)
    {
        bool any = false;
        int t = 0;
        while (t < T) {
            const int s = tri_surf[t], t0 = t;
            while (t < T && tri_surf[t] == s) ++t;
            if (!(vflags[tri[3 * t0]] & 2)) continue;  // first_tri's point 0 decides (dcollid3d.cpp:62)
            any = true;
            for (int q = t0; q < t; ++q)
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < i; ++j) merge(tri[3 * q + i], tri[3 * q + j]);
        }
        return any;
    }
};

struct RigidTopo {
    int nlists = 0, npts = 0;
    int* d_offs = nullptr;
    int* d_pts = nullptr;
    int* d_pt_list = nullptr;
    RigidBodyState* d_state = nullptr;

    void release()
    {
        if (d_offs) cudaFree(d_offs);
        if (d_pts) cudaFree(d_pts);
        if (d_pt_list) cudaFree(d_pt_list);
        if (d_state) cudaFree(d_state);
        d_offs = d_pts = d_pt_list = nullptr;
        d_state = nullptr;
        nlists = npts = 0;
    }

    // makeSet + createImpZoneForRG (dcollid.cpp:1015-1059, dcollid3d.cpp:54-68): topology only.
    int build(int V, int T, const int* tri, const int* tri_surf, const uint8_t* vflags)
    {
        release();
        HostUF uf;
        uf.reset(V);
        const bool any = uf.merge_movable_bodies(T, tri, tri_surf, vflags);
        const std::vector<int>&root = uf.root, &next = uf.next, &weight = uf.weight;
        if (!any) return 0;
        std::vector<int> offs(1, 0), pts, pt_list;
        for (int v = 0; v < V; ++v) {
            // updateImpactZoneVelocityForRG only treats lists of weight > 1 that hold movable points
            if (!(vflags[v] & 2) || root[v] != v || weight[v] <= 1) continue;
            for (int p = v; p >= 0; p = next[p]) {
                pts.push_back(p);
                pt_list.push_back((int)offs.size() - 1);
            }
            offs.push_back((int)pts.size());
        }
        return upload(offs, pts, pt_list);
    }

    // flat lists -> device (also used for the impact zones, which change from iteration to iteration)
    int upload(const std::vector<int>& offs, const std::vector<int>& pts, const std::vector<int>& pt_list)
    {
        release();
        nlists = (int)offs.size() - 1;
        npts = (int)pts.size();
        if (nlists <= 0) { nlists = npts = 0; return 0; }
        if (cudaMalloc((void**)&d_offs, offs.size() * sizeof(int)) != cudaSuccess) return -1;
        if (cudaMalloc((void**)&d_pts, pts.size() * sizeof(int)) != cudaSuccess) return -1;
        if (cudaMalloc((void**)&d_pt_list, pts.size() * sizeof(int)) != cudaSuccess) return -1;
        if (cudaMalloc((void**)&d_state, nlists * sizeof(RigidBodyState)) != cudaSuccess) return -1;
        cudaMemcpy(d_offs, offs.data(), offs.size() * sizeof(int), cudaMemcpyHostToDevice);
        cudaMemcpy(d_pts, pts.data(), pts.size() * sizeof(int), cudaMemcpyHostToDevice);
        cudaMemcpy(d_pt_list, pt_list.data(), pts.size() * sizeof(int), cudaMemcpyHostToDevice);
        return 0;
    }

    int rigidify(const Vec4* xo, Vec4* av, const uint8_t* vflags, double m, double dt, unsigned long long* counters, cudaStream_t st,
                 uint8_t* dirty = nullptr, const unsigned long long* gate = nullptr)
    {
        if (nlists == 0) return 0;
        k_rigid_sums<<<(nlists + 31) / 32, 32, 0, st>>>(nlists, d_offs, d_pts, xo, av, m, dt, d_state);
        k_rigid_apply<<<(npts + 255) / 256, 256, 0, st>>>(npts, d_pts, d_pt_list, d_state, xo, av, vflags, dt, counters, dirty, gate);
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
};

} // namespace clsn
