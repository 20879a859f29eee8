/* Stand-in for the one CGAL kernel the reference names in three unused typedefs
 * (dcollid.cpp:23-25).  No CGAL arithmetic is used anywhere on the path. */
#pragma once
#include <vector>
#include <iostream>
#include <algorithm>
namespace CGAL {
struct Exact_predicates_inexact_constructions_kernel {
    struct Point_3 {};
    struct Triangle_3 {};
};
}
