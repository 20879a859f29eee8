// Drives the C++ host mirror the way the reference's test.cpp drives CollisionSolver3d (test.cpp:96-107):
// reads a flat scene file written by tests/test_gpu_host_cpp.py, builds the POINT/TRI/BOND mesh, runs
// `steps` x { spring solver stand-in; assembleFromInterface; setFrictionConstant; resolveCollision },
// writes final coords + vel.  Usage: host_check scene.bin out.bin steps [ngpus [nozones|pairs]]
// pairs (one GPU): afterwards also exercise isProximity / isCollision on a few triangle pairs (check_single_pairs).
// ngpus > 1: one host thread per GPU, each with its own copy of the mesh and its own CollisionSolver3d(device), joined by
// CollisionSolver::enableMultiGPU -- the multi-GPU step of the library driven through the reference-shaped C++ API; rank r
// writes out.bin.r (all of them must equal the single-GPU out.bin bit for bit).  nozones: impact-zone fail-safe off (it is
// not available in the multi-GPU step).
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "collid_b200.h"
using namespace clsn_host;

template <class T>
static std::vector<T> rd(FILE* f, size_t n)
{
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return v;
}

// what the application stores in TRI::side_length0 / BOND::length0 from the unstretched mesh
static double rest_length(const double* p, const double* q)
{
    double s = 0.0;
    for (int i = 0; i < 3; ++i) s += (p[i] - q[i]) * (p[i] - q[i]);
    return std::sqrt(s);
}

// The single-pair entry points (collid.h:199-200) on a few triangle pairs of the stepped mesh, each evaluated twice: the
// first call of a solver creates its pair context, later calls reuse it -- verdict and accumulators must not depend on
// what the context did before.  Returns the number of mismatches; *fired = evaluations that returned true.
static int check_single_pairs(CollisionSolver3d* solver, std::vector<TRI>& tris, int* fired)
{
    const int T = (int)tris.size(), half = T / 2;
    int bad = 0;
    *fired = 0;
    for (int t = 0; t < half && t < 24; ++t) {
        CD_TRI a(&tris[t]), b(&tris[t + half]);
        for (int mode = 0; mode < 2; ++mode) {
            STATE got[2][6];
            bool verdict[2];
            for (int rep = 0; rep < 2; ++rep) {
                for (int i = 0; i < 3; ++i)
                    for (TRI* tr : {&tris[t], &tris[t + half]}) {
                        STATE* sl = tr->pts[i]->state;
                        for (int j = 0; j < 3; ++j) sl->collsnImpulse[j] = sl->friction[j] = sl->collsnImpulse_RG[j] = 0.0;
                        sl->collsn_num = sl->collsn_num_RG = 0;
                    }
                verdict[rep] = mode ? solver->isCollision(&a, &b) : solver->isProximity(&a, &b);
                for (int i = 0; i < 3; ++i) {
                    got[rep][i] = *tris[t].pts[i]->state;
                    got[rep][3 + i] = *tris[t + half].pts[i]->state;
                }
            }
            if (verdict[0]) ++*fired;
            if (verdict[0] != verdict[1]) ++bad;
            for (int i = 0; i < 6; ++i)
                if (std::memcmp(got[0][i].collsnImpulse, got[1][i].collsnImpulse, sizeof(double) * 3) ||
                    std::memcmp(got[0][i].friction, got[1][i].friction, sizeof(double) * 3) ||
                    std::memcmp(got[0][i].collsnImpulse_RG, got[1][i].collsnImpulse_RG, sizeof(double) * 3) ||
                    got[0][i].collsn_num != got[1][i].collsn_num || got[0][i].collsn_num_RG != got[1][i].collsn_num_RG)
                    ++bad;
        }
    }
    return bad;
}

static int run(const char* scene_path, const std::string& out_path, int steps, int device, int rank, int nranks,
               const unsigned char* id, bool zones, bool pairs = false)
{
    FILE* f = fopen(scene_path, "rb");
    if (!f) return 2;
    int hdr[6];  // V T B n_surf n_curve nhs
    if (fread(hdr, sizeof(int), 6, f) != 6) return 2;
    const int V = hdr[0], T = hdr[1], B = hdr[2], NS = hdr[3], NC = hdr[4], NH = hdr[5];
    auto par = rd<double>(f, 13);  // eps thickness k m lambda cr dt lo[3] hi[3]
    auto x = rd<double>(f, 3 * (size_t)V), vel = rd<double>(f, 3 * (size_t)V);
    auto tri = rd<int>(f, 3 * (size_t)T), tsurf = rd<int>(f, T), bond = rd<int>(f, 2 * (size_t)B), bcur = rd<int>(f, B);
    auto kind = rd<int>(f, NH);
    auto mass = rd<double>(f, NH);
    auto flags = rd<unsigned char>(f, V);
    auto vhs = rd<int>(f, V);
    fclose(f);

    std::vector<HYPER_SURF> hs(NH);
    for (int i = 0; i < NH; ++i) { hs[i] = HYPER_SURF(); hs[i].wave_type = kind[i]; hs[i].body_index = i; hs[i].total_mass = mass[i]; }
    std::vector<STATE> st(V);
    std::vector<POINT> pt(V);
    for (int v = 0; v < V; ++v) {
        st[v] = STATE();
        pt[v] = POINT();
        pt[v].global_index = v; pt[v].state = &st[v]; pt[v].hs = &hs[vhs[v]];
        st[v].is_fixed = flags[v] & 1; st[v].is_movableRG = (flags[v] & 2) != 0;
        for (int j = 0; j < 3; ++j) { pt[v].coords[j] = x[3 * v + j]; st[v].x_old[j] = x[3 * v + j]; st[v].vel[j] = pt[v].vel[j] = vel[3 * v + j]; }
    }
    std::vector<SURFACE> surf(NS);
    std::vector<TRI> tris(T);
    std::vector<TRI*> last(NS, nullptr);
    for (int s = 0; s < NS; ++s) { surf[s].hs = &hs[s]; surf[s].first_tri = nullptr; surf[s].is_bdry = false; }
    for (int t = 0; t < T; ++t) {
        for (int i = 0; i < 3; ++i) tris[t].pts[i] = &pt[tri[3 * t + i]];
        tris[t].next = nullptr; tris[t].surf = &surf[tsurf[t]];
        for (int i = 0; i < 3; ++i) tris[t].side_length0[i] = rest_length(&x[3 * tri[3 * t + i]], &x[3 * tri[3 * t + (i + 1) % 3]]);
        if (last[tsurf[t]]) last[tsurf[t]]->next = &tris[t]; else surf[tsurf[t]].first_tri = &tris[t];
        last[tsurf[t]] = &tris[t];
    }
    std::vector<CURVE> cur(NC);
    std::vector<BOND> bonds(B);
    std::vector<BOND*> blast(NC, nullptr);
    for (int c = 0; c < NC; ++c) { cur[c].first = nullptr; cur[c].is_string = true; cur[c].hs = &hs[NS + c]; }
    for (int b = 0; b < B; ++b) {
        bonds[b].start = &pt[bond[2 * b]]; bonds[b].end = &pt[bond[2 * b + 1]]; bonds[b].next = nullptr;
        bonds[b].length0 = rest_length(&x[3 * bond[2 * b]], &x[3 * bond[2 * b + 1]]);
        if (blast[bcur[b]]) blast[bcur[b]]->next = &bonds[b]; else cur[bcur[b]].first = &bonds[b];
        blast[bcur[b]] = &bonds[b];
    }
    INTERFACE intfc;
    for (auto& s : surf) intfc.surfaces.push_back(&s);
    for (auto& c : cur) intfc.curves.push_back(&c);
    for (int i = 0; i < 3; ++i) { intfc.L[i] = par[7 + i]; intfc.U[i] = par[10 + i]; }

    CollisionSolver3d* solver = new CollisionSolver3d(device);
    if (!zones) solver->setImpactZones(false);
    if (nranks > 1) solver->enableMultiGPU(rank, nranks, id);
    CollisionSolver::setRoundingTolerance(par[0]);
    CollisionSolver::setFabricThickness(par[1]);
    CollisionSolver::setSpringConstant(par[2]);
    CollisionSolver::setPointMass(par[3]);
    CollisionSolver::setRestitutionCoef(par[5]);
    const double dt = par[6];
    for (int step = 0; step < steps; ++step) {
        // FT_Propagate's point hook + dummySpringSolver (test.cpp:192-258)
        for (int v = 0; v < V; ++v)
            for (int j = 0; j < 3; ++j) {
                st[v].x_old[j] = pt[v].coords[j];
                pt[v].coords[j] = st[v].x_old[j] + dt * st[v].vel[j];
            }
        solver->assembleFromInterface(&intfc, dt);
        CollisionSolver::setFrictionConstant(par[4]);
        solver->resolveCollision();
    }
    FILE* o = fopen(out_path.c_str(), "wb");
    for (int v = 0; v < V; ++v) fwrite(pt[v].coords, sizeof(double), 3, o);
    for (int v = 0; v < V; ++v) fwrite(st[v].vel, sizeof(double), 3, o);
    fclose(o);
    printf("host_check[%d/%d]: %d steps, has_collision=%d, ccd passes last step=%d\n", rank, nranks, steps, (int)solver->hasCollision(),
           solver->lastStats().n_ccd_passes);
    int rc = 0;
    if (pairs) {
        int fired = 0;
        const int bad = check_single_pairs(solver, tris, &fired);
        printf("host_check: single-pair entry points: %d evaluations fired, %d mismatches between first and repeated call\n", fired, bad);
        if (bad) rc = 4;
    }
    delete solver;
    return rc;
}

int main(int argc, char** argv)
{
    if (argc < 4) return 2;
    const int steps = atoi(argv[3]);
    const int ngpus = argc > 4 ? atoi(argv[4]) : 1;
    const bool zones = !(argc > 5 && std::strcmp(argv[5], "nozones") == 0) && ngpus <= 1;
    const bool pairs = argc > 5 && std::strcmp(argv[5], "pairs") == 0;
    if (ngpus <= 1) {
        try {
            return run(argv[1], argv[2], steps, 0, 0, 1, nullptr, zones, pairs);
        } catch (const std::exception& e) {   // e.g. no CUDA device: the mirror throws, there is no CPU path behind it
            fprintf(stderr, "host_check: %s\n", e.what());
            return 3;
        }
    }
    unsigned char id[128];
    CollisionSolver::multiGPUUniqueId(id);
    std::vector<std::thread> th;
    std::vector<int> rc((size_t)ngpus, 0);
    for (int r = 0; r < ngpus; ++r)
        th.emplace_back([&, r]() {
            try {
                rc[r] = run(argv[1], std::string(argv[2]) + "." + std::to_string(r), steps, r, r, ngpus, id, false);
            } catch (const std::exception& e) {
                fprintf(stderr, "rank %d: %s\n", r, e.what());
                rc[r] = 3;
            }
        });
    for (auto& t : th) t.join();
    for (int r = 0; r < ngpus; ++r)
        if (rc[r]) return rc[r];
    return 0;
}
