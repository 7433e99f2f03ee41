// simt.cuh -- the per-lane value types the kernel body is written in.
//
// nvcc (product build):  VD = double, VI = int, VB = bool -- one value per thread, shuffles are
//                        __shfl_*_sync, reductions are xor butterflies.  Zero overhead.
// g++ -DACMPC_EMULATE :  VD / VI / VB hold the values of all 32 lanes of a warp and every operation
//                        is applied lane by lane, shuffles are permutations.  This lets the GPU-less
//                        container execute the *same* warp-parallel algorithm (lane mapping, shuffles,
//                        cyclic reductions) and compare it with the oracle (tests/_emul).  TEST ONLY:
//                        it is never loaded by the package and is not a fallback.
//
// Rules the body follows so both builds mean the same thing: control flow depends on warp-uniform
// values only (plain int/double/bool); per-lane conditions go through vsel(); memory is touched
// through the ld_/st_ helpers below.
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef ACMPC_EMULATE
// ------------------------------------------------------------------------------------------------
#define AC_DEV static inline
#define AC_HD
#define AC_MEM inline
#define AC_UNROLL
#define AC_NOUNROLL
#define AC_LANE0 if (true)
#define AC_FOR_LANES for (int i_ = 0; i_ < 32; ++i_)

namespace acmpc {

struct VB {
    bool v[32];
};
struct VI {
    int v[32];
};
struct VD {
    double v[32];
    VD() {}
    VD(double s) { AC_FOR_LANES v[i_] = s; }
};

#define AC_BIN(op)                                                         \
    AC_DEV VD operator op(const VD& a, const VD& b)                        \
    {                                                                      \
        VD r;                                                              \
        AC_FOR_LANES r.v[i_] = a.v[i_] op b.v[i_];                         \
        return r;                                                          \
    }
AC_BIN(+) AC_BIN(-) AC_BIN(*) AC_BIN(/)
#undef AC_BIN
AC_DEV VD operator-(const VD& a)
{
    VD r;
    AC_FOR_LANES r.v[i_] = -a.v[i_];
    return r;
}
AC_DEV VD& operator+=(VD& a, const VD& b) { return a = a + b; }
AC_DEV VD& operator-=(VD& a, const VD& b) { return a = a - b; }
AC_DEV VD& operator*=(VD& a, const VD& b) { return a = a * b; }

#define AC_CMP(op)                                                         \
    AC_DEV VB operator op(const VD& a, const VD& b)                        \
    {                                                                      \
        VB r;                                                              \
        AC_FOR_LANES r.v[i_] = a.v[i_] op b.v[i_];                         \
        return r;                                                          \
    }
AC_CMP(<) AC_CMP(>) AC_CMP(<=) AC_CMP(>=)
#undef AC_CMP
AC_DEV VB operator&(const VB& a, const VB& b)
{
    VB r;
    AC_FOR_LANES r.v[i_] = a.v[i_] && b.v[i_];
    return r;
}
AC_DEV VB operator|(const VB& a, const VB& b)
{
    VB r;
    AC_FOR_LANES r.v[i_] = a.v[i_] || b.v[i_];
    return r;
}
AC_DEV VB operator!(const VB& a)
{
    VB r;
    AC_FOR_LANES r.v[i_] = !a.v[i_];
    return r;
}
AC_DEV VB vb_all(bool s)
{
    VB r;
    AC_FOR_LANES r.v[i_] = s;
    return r;
}

AC_DEV VI lane_iota()
{
    VI r;
    AC_FOR_LANES r.v[i_] = i_;
    return r;
}
AC_DEV VI operator*(const VI& a, int b)
{
    VI r;
    AC_FOR_LANES r.v[i_] = a.v[i_] * b;
    return r;
}
AC_DEV VI operator+(const VI& a, int b)
{
    VI r;
    AC_FOR_LANES r.v[i_] = a.v[i_] + b;
    return r;
}
AC_DEV VI operator|(const VI& a, const VI& b)
{
    VI r;
    AC_FOR_LANES r.v[i_] = a.v[i_] | b.v[i_];
    return r;
}
AC_DEV VI operator<<(const VI& a, int b)
{
    VI r;
    AC_FOR_LANES r.v[i_] = a.v[i_] << b;
    return r;
}
AC_DEV VI vi_all(int s)
{
    VI r;
    AC_FOR_LANES r.v[i_] = s;
    return r;
}
// ((a >> sh) & 3) == which
AC_DEV VB vi_field_is(const VI& a, int sh, int which)
{
    VB r;
    AC_FOR_LANES r.v[i_] = ((a.v[i_] >> sh) & 3) == which;
    return r;
}
#define AC_ICMP(name, op)                                                  \
    AC_DEV VB name(const VI& a, int b)                                     \
    {                                                                      \
        VB r;                                                              \
        AC_FOR_LANES r.v[i_] = a.v[i_] op b;                               \
        return r;                                                          \
    }
AC_ICMP(vi_lt, <) AC_ICMP(vi_le, <=) AC_ICMP(vi_ge, >=) AC_ICMP(vi_eq, ==)
#undef AC_ICMP

AC_DEV VD vsel(const VB& m, const VD& a, const VD& b)
{
    VD r;
    AC_FOR_LANES r.v[i_] = m.v[i_] ? a.v[i_] : b.v[i_];
    return r;
}
AC_DEV VI vseli(const VB& m, const VI& a, const VI& b)
{
    VI r;
    AC_FOR_LANES r.v[i_] = m.v[i_] ? a.v[i_] : b.v[i_];
    return r;
}

#define AC_UN(name, expr)                                                  \
    AC_DEV VD name(const VD& a)                                            \
    {                                                                      \
        VD r;                                                              \
        AC_FOR_LANES r.v[i_] = expr(a.v[i_]);                              \
        return r;                                                          \
    }
AC_UN(vabs, fabs) AC_UN(vsqrt, sqrt) AC_UN(vsin, sin) AC_UN(vcos, cos) AC_UN(vatan, atan)
#undef AC_UN
AC_DEV VD vmax(const VD& a, const VD& b)
{
    VD r;
    AC_FOR_LANES r.v[i_] = fmax(a.v[i_], b.v[i_]);
    return r;
}
AC_DEV VD vmin(const VD& a, const VD& b)
{
    VD r;
    AC_FOR_LANES r.v[i_] = fmin(a.v[i_], b.v[i_]);
    return r;
}
AC_DEV VD vatan2(const VD& y, const VD& x)
{
    VD r;
    AC_FOR_LANES r.v[i_] = atan2(y.v[i_], x.v[i_]);
    return r;
}
AC_DEV VD vfmod(const VD& a, double b)
{
    VD r;
    AC_FOR_LANES r.v[i_] = fmod(a.v[i_], b);
    return r;
}

// shuffles: out-of-range sources deliver 0 (the "0" variants) or the lane's own value
AC_DEV VD shfl_up0(const VD& a, int d)
{
    VD r;
    AC_FOR_LANES r.v[i_] = (i_ >= d) ? a.v[i_ - d] : 0.0;
    return r;
}
AC_DEV VD shfl_down0(const VD& a, int d)
{
    VD r;
    AC_FOR_LANES r.v[i_] = (i_ + d < 32) ? a.v[i_ + d] : 0.0;
    return r;
}
// "raw" shuffles: an out-of-range source delivers the lane's OWN value (the hardware behaviour); callers
// multiply the result by a coefficient that is zero on those lanes
AC_DEV VD shfl_up_raw(const VD& a, int d)
{
    VD r;
    AC_FOR_LANES r.v[i_] = (i_ >= d) ? a.v[i_ - d] : a.v[i_];
    return r;
}
AC_DEV VD shfl_down_raw(const VD& a, int d)
{
    VD r;
    AC_FOR_LANES r.v[i_] = (i_ + d < 32) ? a.v[i_ + d] : a.v[i_];
    return r;
}
// value of lane-1, lane 0 takes lane 31's (rotation)
AC_DEV VD shfl_rot_up1(const VD& a)
{
    VD r;
    AC_FOR_LANES r.v[i_] = a.v[(i_ + 31) & 31];
    return r;
}
// warp-uniform predicate (all lanes hold the same value; the device build tells the compiler so)
AC_DEV bool uni(bool p) { return p; }
AC_DEV VD vrsqrt(const VD& a)
{
    VD r;
    AC_FOR_LANES r.v[i_] = 1.0 / sqrt(a.v[i_]);
    return r;
}
AC_DEV double urecip(double a) { return 1.0 / a; }
AC_DEV VD vrecip(const VD& a)
{
    VD r;
    AC_FOR_LANES r.v[i_] = 1.0 / a.v[i_];
    return r;
}
AC_DEV double lane_value(const VD& a, int src) { return a.v[src]; }

// warp reductions (same xor-butterfly order as the device build); result is warp-uniform
AC_DEV double wmax(const VD& a)
{
    VD t = a;
    for (int o = 16; o > 0; o >>= 1) {
        VD u;
        AC_FOR_LANES u.v[i_] = fmax(t.v[i_], t.v[i_ ^ o]);
        t = u;
    }
    return t.v[0];
}
AC_DEV double wsum(const VD& a)
{
    VD t = a;
    for (int o = 16; o > 0; o >>= 1) {
        VD u;
        AC_FOR_LANES u.v[i_] = t.v[i_] + t.v[i_ ^ o];
        t = u;
    }
    return t.v[0];
}
AC_DEV bool wany(const VB& a)
{
    bool r = false;
    AC_FOR_LANES r = r || a.v[i_];
    return r;
}

// memory: base[lane], base[idx], predicated stores
AC_DEV VD ld_lane(const double* base)
{
    VD r;
    AC_FOR_LANES r.v[i_] = base[i_];
    return r;
}
AC_DEV void st_lane(double* base, const VD& a) { AC_FOR_LANES base[i_] = a.v[i_]; }
AC_DEV VD ld_idx_if(const VB& m, const double* base, const VI& idx, double other)
{
    VD r;
    AC_FOR_LANES r.v[i_] = m.v[i_] ? base[idx.v[i_]] : other;
    return r;
}
AC_DEV void st_idx_if(const VB& m, double* base, const VI& idx, const VD& a)
{
    AC_FOR_LANES if (m.v[i_]) base[idx.v[i_]] = a.v[i_];
}
// base[min(lane + off, 31)]: the slot of the lane `off` places up, clamped (callers multiply the
// clamped reads by a shuffled-in zero)
AC_DEV VD ld_lane_at(const double* base, int off)
{
    VD r;
    AC_FOR_LANES r.v[i_] = base[(i_ + off < 32) ? i_ + off : 31];
    return r;
}

// ---- tensor memory (TMEM) as a lane-private scratchpad: the emulation models it as [double column][lane]
struct Tm {
    double* p;
};
template <int N>
AC_DEV void tm_ld(const Tm& t, int dcol, VD (&o)[N])
{
    for (int k = 0; k < N; ++k) AC_FOR_LANES o[k].v[i_] = t.p[(dcol + k) * 32 + i_];
}
template <int N>
AC_DEV void tm_st(const Tm& t, int dcol, const VD (&v)[N])
{
    for (int k = 0; k < N; ++k) AC_FOR_LANES t.p[(dcol + k) * 32 + i_] = v[k].v[i_];
}
AC_DEV VD tm_ld1(const Tm& t, int dcol)
{
    VD r;
    AC_FOR_LANES r.v[i_] = t.p[dcol * 32 + i_];
    return r;
}
// grouped loads (device: several tcgen05.ld under one wait)
template <int N0, int N1, int N2, int N3>
AC_DEV void tm_ld_group4(const Tm& t, int d0, VD (&o0)[N0], int d1, VD (&o1)[N1], int d2, VD (&o2)[N2], int d3, VD (&o3)[N3])
{
    tm_ld<N0>(t, d0, o0), tm_ld<N1>(t, d1, o1), tm_ld<N2>(t, d2, o2), tm_ld<N3>(t, d3, o3);
}
template <int N0, int N1>
AC_DEV void tm_ld_group2(const Tm& t, int d0, VD (&o0)[N0], int d1, VD (&o1)[N1])
{
    tm_ld<N0>(t, d0, o0), tm_ld<N1>(t, d1, o1);
}
AC_DEV void warp_sync() {}
// dst[0..count) = src[0..count), the warp's lanes taking consecutive elements (coalesced rows on the device)
AC_DEV void warp_copy(double* dst, const double* src, int count)
{
    for (int i = 0; i < count; ++i) dst[i] = src[i];
}

}  // namespace acmpc

#else
// ------------------------------------------------------------------------------------------------
#define AC_DEV __device__ __forceinline__
#define AC_HD __host__ __device__
#define AC_MEM __device__ __forceinline__
#define AC_UNROLL _Pragma("unroll")
#define AC_NOUNROLL _Pragma("unroll 1")
#define AC_LANE0 if ((threadIdx.x & 31) == 0)

namespace acmpc {

using VD = double;
using VI = int;
using VB = bool;

constexpr unsigned kFull = 0xffffffffu;

AC_DEV VB vb_all(bool s) { return s; }
AC_DEV VI lane_iota() { return (int)(threadIdx.x & 31); }
AC_DEV VI vi_all(int s) { return s; }
AC_DEV VB vi_field_is(VI a, int sh, int which) { return ((a >> sh) & 3) == which; }
AC_DEV VB vi_lt(VI a, int b) { return a < b; }
AC_DEV VB vi_le(VI a, int b) { return a <= b; }
AC_DEV VB vi_ge(VI a, int b) { return a >= b; }
AC_DEV VB vi_eq(VI a, int b) { return a == b; }
AC_DEV VD vsel(bool m, double a, double b) { return m ? a : b; }
AC_DEV VI vseli(bool m, int a, int b) { return m ? a : b; }
AC_DEV VD vabs(double a) { return fabs(a); }
AC_DEV VD vsqrt(double a) { return sqrt(a); }
AC_DEV VD vsin(double a) { return sin(a); }
AC_DEV VD vcos(double a) { return cos(a); }
AC_DEV VD vatan(double a) { return atan(a); }
// compare-select instead of fmax/fmin: no NaN-quieting sequence (the operands here are never NaN)
AC_DEV VD vmax(double a, double b) { return a > b ? a : b; }
AC_DEV VD vmin(double a, double b) { return a < b ? a : b; }
AC_DEV VD vatan2(double y, double x) { return atan2(y, x); }
AC_DEV VD vfmod(double a, double b) { return fmod(a, b); }

AC_DEV VD shfl_up0(double a, int d)
{
    double t = __shfl_up_sync(kFull, a, d);
    return ((int)(threadIdx.x & 31) >= d) ? t : 0.0;
}
AC_DEV VD shfl_down0(double a, int d)
{
    double t = __shfl_down_sync(kFull, a, d);
    return ((int)(threadIdx.x & 31) + d < 32) ? t : 0.0;
}
AC_DEV VD shfl_up_raw(double a, int d) { return __shfl_up_sync(kFull, a, d); }
AC_DEV VD shfl_down_raw(double a, int d) { return __shfl_down_sync(kFull, a, d); }
AC_DEV VD shfl_rot_up1(double a) { return __shfl_sync(kFull, a, ((int)(threadIdx.x & 31) + 31) & 31); }
AC_DEV bool uni(bool p) { return __any_sync(kFull, p) != 0; }
// 1/sqrt(a) and 1/a for a normal a (equilibration norms clamped to [1e-4, 1e4], pivots of SPD blocks, rho):
// hardware seed (MUFU.RSQ64H / RCP64H, 20 bits) + two Newton steps, fully inline.  tools/seed_probe.cu measured
// them against the correctly rounded results over [1e-6, 1e6]: 2.7e-16 and 0 (bit-exact) -- CUDA's rsqrt() calls
// a helper per value and its division carries range checks and a slow path.
AC_DEV double urecip(double a)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double e = fma(-a, y, 1.0);
        y = fma(y, e, y);
    }
    return y;
}
AC_DEV VD vrecip(double a) { return urecip(a); }
AC_DEV VD vrsqrt(double a)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double e = fma(-(h * y), y, 0.5);
        y = fma(y, e, y);
    }
    return y;
}
AC_DEV double lane_value(double a, int src) { return __shfl_sync(kFull, a, src); }

// max over the warp of NON-NEGATIVE values (norms, magnitudes): two 32-bit REDUX instead of a five-level
// 64-bit shuffle butterfly -- for non-negative doubles the (high word, low word) order is the numeric order
AC_DEV double wmax(double v)
{
    const int hi = __double2hiint(v);   // signed: a -0.0 loses against everything instead of winning
    const unsigned lo = (unsigned)__double2loint(v);
    const int hmax = __reduce_max_sync(kFull, hi);
    const unsigned lmax = __reduce_max_sync(kFull, hi == hmax ? lo : 0u);
    return __hiloint2double(hmax, (int)lmax);
}
AC_DEV double wsum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
AC_DEV bool wany(bool p) { return __any_sync(kFull, p) != 0; }

AC_DEV VD ld_lane(const double* base) { return base[threadIdx.x & 31]; }
AC_DEV void st_lane(double* base, double a) { base[threadIdx.x & 31] = a; }
AC_DEV VD ld_idx_if(bool m, const double* base, int idx, double other) { return m ? base[idx] : other; }
AC_DEV void st_idx_if(bool m, double* base, int idx, double a)
{
    if (m) base[idx] = a;
}
AC_DEV VD ld_lane_at(const double* base, int off)
{
    int i = (int)(threadIdx.x & 31) + off;
    return base[i < 32 ? i : 31];
}

// ---- tensor memory (TMEM) as a lane-private scratchpad.  One double = two 32-bit columns of the lane's
// row; `a` = TMEM address (lane quarter of the warp << 16 | first column of the instance).  Every load is
// issued together with its tcgen05.wait::ld so the destination registers are defined when the asm ends.
struct Tm {
    uint32_t a;
};
AC_DEV void tm_ld_raw1(uint32_t addr, double* o)
{
    uint32_t r[2];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1])
                 : "r"(addr)
                 : "memory");
#pragma unroll
    for (int k = 0; k < 1; ++k) o[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}
AC_DEV void tm_st_raw1(uint32_t addr, const double* v)
{
    uint32_t r[2];
#pragma unroll
    for (int k = 0; k < 1; ++k) r[2 * k] = (uint32_t)__double2loint(v[k]), r[2 * k + 1] = (uint32_t)__double2hiint(v[k]);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%2], {%0,%1};\n\ttcgen05.wait::st.sync.aligned;"
                 :
                 : "r"(r[0]), "r"(r[1]), "r"(addr)
                 : "memory");
}
AC_DEV void tm_ld_raw2(uint32_t addr, double* o)
{
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr)
                 : "memory");
#pragma unroll
    for (int k = 0; k < 2; ++k) o[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}
AC_DEV void tm_st_raw2(uint32_t addr, const double* v)
{
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 2; ++k) r[2 * k] = (uint32_t)__double2loint(v[k]), r[2 * k + 1] = (uint32_t)__double2hiint(v[k]);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%4], {%0,%1,%2,%3};\n\ttcgen05.wait::st.sync.aligned;"
                 :
                 : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(addr)
                 : "memory");
}
AC_DEV void tm_ld_raw4(uint32_t addr, double* o)
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(addr)
                 : "memory");
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}
AC_DEV void tm_st_raw4(uint32_t addr, const double* v)
{
    uint32_t r[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) r[2 * k] = (uint32_t)__double2loint(v[k]), r[2 * k + 1] = (uint32_t)__double2hiint(v[k]);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};\n\ttcgen05.wait::st.sync.aligned;"
                 :
                 : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(addr)
                 : "memory");
}
AC_DEV void tm_ld_raw8(uint32_t addr, double* o)
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr)
                 : "memory");
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}
AC_DEV void tm_st_raw8(uint32_t addr, const double* v)
{
    uint32_t r[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[2 * k] = (uint32_t)__double2loint(v[k]), r[2 * k + 1] = (uint32_t)__double2hiint(v[k]);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};\n\ttcgen05.wait::st.sync.aligned;"
                 :
                 : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(addr)
                 : "memory");
}
AC_DEV void tm_ld_raw16(uint32_t addr, double* o)
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(addr)
                 : "memory");
#pragma unroll
    for (int k = 0; k < 16; ++k) o[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}
AC_DEV void tm_st_raw16(uint32_t addr, const double* v)
{
    uint32_t r[32];
#pragma unroll
    for (int k = 0; k < 16; ++k) r[2 * k] = (uint32_t)__double2loint(v[k]), r[2 * k + 1] = (uint32_t)__double2hiint(v[k]);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};\n\ttcgen05.wait::st.sync.aligned;"
                 :
                 : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(addr)
                 : "memory");
}
// Grouped loads: several tcgen05.ld issued back to back under ONE tcgen05.wait::ld, so a phase that needs three or four
// chunks of a stage exposes the tensor-memory latency once instead of once per chunk.
AC_DEV void tm_ld_group_8_1_16_8(uint32_t a0, double (&o0)[8], uint32_t a1, double (&o1)[1], uint32_t a2, double (&o2)[16], uint32_t a3, double (&o3)[8])
{
    uint32_t r[66];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%66];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%16,%17}, [%67];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49}, [%68];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63,%64,%65}, [%69];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63]), "=r"(r[64]), "=r"(r[65])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3)
                 : "memory");
#pragma unroll
    for (int k = 0; k < 8; ++k) o0[k] = __hiloint2double((int)r[0 + 2 * k + 1], (int)r[0 + 2 * k]);
#pragma unroll
    for (int k = 0; k < 1; ++k) o1[k] = __hiloint2double((int)r[16 + 2 * k + 1], (int)r[16 + 2 * k]);
#pragma unroll
    for (int k = 0; k < 16; ++k) o2[k] = __hiloint2double((int)r[18 + 2 * k + 1], (int)r[18 + 2 * k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) o3[k] = __hiloint2double((int)r[50 + 2 * k + 1], (int)r[50 + 2 * k]);
}
AC_DEV void tm_ld_group_16_8_2_16(uint32_t a0, double (&o0)[16], uint32_t a1, double (&o1)[8], uint32_t a2, double (&o2)[2], uint32_t a3, double (&o3)[16])
{
    uint32_t r[84];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%84];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47}, [%85];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%48,%49,%50,%51}, [%86];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63,%64,%65,%66,%67,%68,%69,%70,%71,%72,%73,%74,%75,%76,%77,%78,%79,%80,%81,%82,%83}, [%87];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63]), "=r"(r[64]), "=r"(r[65]), "=r"(r[66]), "=r"(r[67]), "=r"(r[68]), "=r"(r[69]), "=r"(r[70]), "=r"(r[71]), "=r"(r[72]), "=r"(r[73]), "=r"(r[74]), "=r"(r[75]), "=r"(r[76]), "=r"(r[77]), "=r"(r[78]), "=r"(r[79]), "=r"(r[80]), "=r"(r[81]), "=r"(r[82]), "=r"(r[83])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3)
                 : "memory");
#pragma unroll
    for (int k = 0; k < 16; ++k) o0[k] = __hiloint2double((int)r[0 + 2 * k + 1], (int)r[0 + 2 * k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) o1[k] = __hiloint2double((int)r[32 + 2 * k + 1], (int)r[32 + 2 * k]);
#pragma unroll
    for (int k = 0; k < 2; ++k) o2[k] = __hiloint2double((int)r[48 + 2 * k + 1], (int)r[48 + 2 * k]);
#pragma unroll
    for (int k = 0; k < 16; ++k) o3[k] = __hiloint2double((int)r[52 + 2 * k + 1], (int)r[52 + 2 * k]);
}
AC_DEV void tm_ld_group_16_16(uint32_t a0, double (&o0)[16], uint32_t a1, double (&o1)[16])
{
    uint32_t r[64];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%64];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%65];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                 : "r"(a0), "r"(a1)
                 : "memory");
#pragma unroll
    for (int k = 0; k < 16; ++k) o0[k] = __hiloint2double((int)r[0 + 2 * k + 1], (int)r[0 + 2 * k]);
#pragma unroll
    for (int k = 0; k < 16; ++k) o1[k] = __hiloint2double((int)r[32 + 2 * k + 1], (int)r[32 + 2 * k]);
}
template <int N>
AC_DEV void tm_ld(const Tm& t, int dcol, double (&o)[N])
{
    static_assert(N == 1 || N == 2 || N == 4 || N == 8 || N == 16, "power-of-two chunk");
    const uint32_t addr = t.a + 2u * (uint32_t)dcol;
    if (N == 1) tm_ld_raw1(addr, o);
    if (N == 2) tm_ld_raw2(addr, o);
    if (N == 4) tm_ld_raw4(addr, o);
    if (N == 8) tm_ld_raw8(addr, o);
    if (N == 16) tm_ld_raw16(addr, o);
}
template <int N>
AC_DEV void tm_st(const Tm& t, int dcol, const double (&v)[N])
{
    static_assert(N == 1 || N == 2 || N == 4 || N == 8 || N == 16, "power-of-two chunk");
    const uint32_t addr = t.a + 2u * (uint32_t)dcol;
    if (N == 1) tm_st_raw1(addr, v);
    if (N == 2) tm_st_raw2(addr, v);
    if (N == 4) tm_st_raw4(addr, v);
    if (N == 8) tm_st_raw8(addr, v);
    if (N == 16) tm_st_raw16(addr, v);
}
AC_DEV double tm_ld1(const Tm& t, int dcol)
{
    double o[1];
    tm_ld_raw1(t.a + 2u * (uint32_t)dcol, o);
    return o[0];
}
// grouped loads: the shapes the ADMM iteration uses
template <int N0, int N1, int N2, int N3>
AC_DEV void tm_ld_group4(const Tm& t, int d0, double (&o0)[N0], int d1, double (&o1)[N1], int d2, double (&o2)[N2], int d3,
                         double (&o3)[N3]);
template <>
AC_DEV void tm_ld_group4<8, 1, 16, 8>(const Tm& t, int d0, double (&o0)[8], int d1, double (&o1)[1], int d2, double (&o2)[16],
                                      int d3, double (&o3)[8])
{
    tm_ld_group_8_1_16_8(t.a + 2u * (uint32_t)d0, o0, t.a + 2u * (uint32_t)d1, o1, t.a + 2u * (uint32_t)d2, o2,
                         t.a + 2u * (uint32_t)d3, o3);
}
template <>
AC_DEV void tm_ld_group4<16, 8, 2, 16>(const Tm& t, int d0, double (&o0)[16], int d1, double (&o1)[8], int d2, double (&o2)[2],
                                       int d3, double (&o3)[16])
{
    tm_ld_group_16_8_2_16(t.a + 2u * (uint32_t)d0, o0, t.a + 2u * (uint32_t)d1, o1, t.a + 2u * (uint32_t)d2, o2,
                          t.a + 2u * (uint32_t)d3, o3);
}
template <int N0, int N1>
AC_DEV void tm_ld_group2(const Tm& t, int d0, double (&o0)[N0], int d1, double (&o1)[N1]);
template <>
AC_DEV void tm_ld_group2<16, 16>(const Tm& t, int d0, double (&o0)[16], int d1, double (&o1)[16])
{
    tm_ld_group_16_16(t.a + 2u * (uint32_t)d0, o0, t.a + 2u * (uint32_t)d1, o1);
}
AC_DEV void warp_sync() { __syncwarp(); }
AC_DEV void warp_copy(double* dst, const double* src, int count)
{
    for (int i = (int)(threadIdx.x & 31); i < count; i += 32) dst[i] = src[i];
}

}  // namespace acmpc
#endif
