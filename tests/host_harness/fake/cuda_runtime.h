#pragma once
#include <cmath>
#include <cfloat>
#include <algorithm>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
using std::min; using std::max;
