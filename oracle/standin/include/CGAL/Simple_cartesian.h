/* Empty stand-in: the reference includes this CGAL header but every CGAL call site is
 * commented out (dcollid.cpp:245-252,392-400,444-451).  The reference relies on it for
 * transitive standard headers, so provide those. */
#pragma once
#include <vector>
#include <iostream>
#include <algorithm>
#include <cstddef>
