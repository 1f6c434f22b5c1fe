// Stub shadowing the reference's include/cl2.hpp (OpenCL C++ bindings) -- TEST INFRASTRUCTURE.
// Provides only the scalar typedefs and the one cl::Platform query that src/utils.h:57 names,
// so the reference's host-side data producers (bvh.cpp, sbvh.cpp, bvhnode.cpp, envmap.cpp)
// compile without an OpenCL SDK.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>
typedef float cl_float;
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint8_t cl_uchar;
#define CL_PLATFORM_NAME 0x0902
namespace cl
{
struct Platform
{
    template <int N> std::string getInfo() { return std::string(); }
};
} // namespace cl
