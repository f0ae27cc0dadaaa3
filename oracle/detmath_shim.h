// oracle/detmath_shim.h -- TEST INFRASTRUCTURE.  Force-included (-include) when building the
// "detmath" flavour of oracle/_ref: the reference's src/render.cpp and the inline headers it
// pulls in then call include/tb200_detmath.h instead of libm for the transcendentals on the
// hot path.  Every standard header the reference uses is included first so that the macros
// below cannot rewrite declarations inside the standard library.
#pragma once
#include <math.h>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <cfloat>
#include <climits>
#include <cassert>
#include <algorithm>
#include <limits>
#include <random>
#include <vector>
#include <map>
#include <string>
#include <thread>
#include <stdint.h>

#include "tb200_detmath.h"

static inline float tb_shim_sin(float x) { return tbm_sinf(x); }
static inline double tb_shim_sin(double x) { return ::sin(x); }
static inline float tb_shim_cos(float x) { return tbm_cosf(x); }
static inline double tb_shim_cos(double x) { return ::cos(x); }
static inline float tb_shim_atan2(float y, float x) { return tbm_atan2f(y, x); }
static inline double tb_shim_atan2(double y, double x) { return ::atan2(y, x); }

#define sinf tbm_sinf
#define cosf tbm_cosf
#define expf tbm_expf
#define acosf tbm_acosf
#define atan2f tbm_atan2f
#define powf tbm_powf
#define sin tb_shim_sin
#define cos tb_shim_cos
#define atan2 tb_shim_atan2
