/* FronTier stand-in for the parity oracle  (TEST INFRASTRUCTURE, not product code).
 *
 * The reference (antdvid/Collision) lives inside a FronTier checkout (Makefile:19-24,
 * collid.h:6-7) that is not available here.  Its three hot-path translation units
 * (AABB.cpp, dcollid.cpp, dcollid3d.cpp) touch only the handful of FronTier types,
 * accessor macros and helpers declared below, so they compile UNMODIFIED against
 * this header.  Nothing here is copied from FronTier; the arithmetic macros are the
 * textbook forms (left-to-right dot product, component cross product), which is the
 * adopted definition of "the reference's result" -- see DESIGN.md ("parity pinning").
 */
#ifndef CLSN_ORACLE_FRONTIER_STANDIN_H
#define CLSN_ORACLE_FRONTIER_STANDIN_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <fenv.h>
#include <vector>
#include <iostream>
#include <algorithm>
#include <stdexcept>

#define MAXD 3
#define YES 1
#define NO 0
#define ERROR (-1)
#ifdef HUGE
#undef HUGE
#endif
#define HUGE 1.0e+18
#define MACH_EPS DBL_EPSILON

typedef void* POINTER;

enum {
    UNKNOWN_WAVE_TYPE = 0,
    NEUMANN_BOUNDARY = 4,
    MOVABLE_BODY_BOUNDARY = 5,
    FIRST_PHYSICS_WAVE_TYPE = 12
};
enum { UNKNOWN_HSBDRY = 0, STRING_HSBDRY = 7 };

struct HYPER_SURF {
    int wave_type;
    int body_index;
    double total_mass;
    double center_of_mass[3];
    double center_of_mass_velo[3];
};

struct POINT {
    double _coords[3];
    long global_index;
    long indx;
    int _sorted;
    POINTER _left_state;
    POINTER _right_state;
    HYPER_SURF* hs;
    double vel[3];
};

struct SURFACE;
struct TRI {
    POINT* __pts[3];
    TRI* prev;
    TRI* next;
    SURFACE* surf;
    double side_length0[3];
};

struct SURFACE {
    HYPER_SURF* hyper_surf;
    TRI* _first_tri;
    int _is_bdry;
};

struct BOND {
    POINT* start;
    POINT* end;
    BOND* prev;
    BOND* next;
    double length0;
};

struct CURVE {
    BOND* first;
    BOND* last;
    int _hsbdry_type;
};

struct RECT_GRID {
    double L[3];
    double U[3];
};
struct Table {
    RECT_GRID rect_grid;
};
struct INTERFACE {
    SURFACE** surfaces; /* NULL-terminated */
    CURVE** curves;     /* NULL-terminated */
    Table* table;
};
struct Front {
    INTERFACE* interf;
    double dt;
};

#define Coords(p) ((p)->_coords)
#define left_state(p) ((p)->_left_state)
#define right_state(p) ((p)->_right_state)
#define sorted(p) ((p)->_sorted)
#define Point_of_tri(t) ((t)->__pts)
#define is_bdry(s) ((s)->_is_bdry)
#define hsbdry_type(c) ((c)->_hsbdry_type)
#define first_tri(s) ((s)->_first_tri)
#define at_end_of_tri_list(t, s) ((t) == NULL)
#define wave_type(hs) ((hs)->wave_type)
#define body_index(hs) ((hs)->body_index)
#define total_mass(hs) ((hs)->total_mass)
#define center_of_mass(hs) ((hs)->center_of_mass)
#define center_of_mass_velo(hs) ((hs)->center_of_mass_velo)

#define intfc_surface_loop(intfc, s) for ((s) = (intfc)->surfaces; (s) && *(s); ++(s))
#define surf_tri_loop(s, tri) \
    for ((tri) = first_tri(s); !at_end_of_tri_list(tri, s); (tri) = (tri)->next)
#define intfc_curve_loop(intfc, c) for ((c) = (intfc)->curves; (c) && *(c); ++(c))
#define curve_bond_loop(c, b) for ((b) = (c)->first; (b) != NULL; (b) = (b)->next)

#define sqr(x) ((x) * (x))
#define Dot3d(A, B) ((A)[0] * (B)[0] + (A)[1] * (B)[1] + (A)[2] * (B)[2])
#define Mag3d(A) sqrt(Dot3d(A, A))
#define Cross3d(B, C, ans)                                   \
    {                                                        \
        (ans)[0] = ((B)[1]) * ((C)[2]) - ((B)[2]) * ((C)[1]); \
        (ans)[1] = ((B)[2]) * ((C)[0]) - ((B)[0]) * ((C)[2]); \
        (ans)[2] = ((B)[0]) * ((C)[1]) - ((B)[1]) * ((C)[0]); \
    }

static inline double distance_between_positions(const double* p, const double* q, int dim)
{
    double s = 0.0;
    for (int i = 0; i < dim; ++i) s += sqr(p[i] - q[i]);
    return sqrt(s);
}

/* Runtime services.  Implemented in ref_wrapper.cpp. */
struct clsn_ref_abort : public std::runtime_error {
    clsn_ref_abort() : std::runtime_error("reference clean_up(ERROR)") {}
};
bool debugging(const char*);
void clean_up(int);
void start_clock(const char*);
void stop_clock(const char*);
double cpu_seconds();
bool create_directory(const char*, int);

#endif
