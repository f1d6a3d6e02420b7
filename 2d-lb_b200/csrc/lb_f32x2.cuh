// Packed single precision for sm_100a: two IEEE fp32 lanes per 64-bit register pair, one issue slot per
// operation (SASS FADD2 / FMUL2 / FFMA2; PTX add|sub|mul|fma.rn.f32x2).  Every lane is rounded exactly
// like the scalar instruction, so code written over `F2` produces the same bits as the same code written
// over `float` -- the STRICT contract (lb_device.cuh) holds lane by lane -- while the fp32 collision
// issues half as many floating-point instructions.  The per-node arithmetic templates of lb_device.cuh
// are instantiated with T = F2 for the two nodes a thread owns side by side.
//
// The host versions (plain C on the two halves) exist so that the `__host__ __device__` templates also
// compile for the CPU replay tools; they are never on a product path.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace lb {

#define LB_F2_HD __host__ __device__ __forceinline__

struct F2 {
    unsigned long long r;      // {lo, hi} = two floats
    F2() = default;
    LB_F2_HD F2(float lo, float hi)
    {
#ifdef __CUDA_ARCH__
        asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
#else
        uint32_t a, b;
        memcpy(&a, &lo, 4); memcpy(&b, &hi, 4);
        r = ((unsigned long long)b << 32) | a;
#endif
    }
    LB_F2_HD F2(float a) : F2(a, a) {}
    LB_F2_HD explicit F2(double a) : F2((float)a, (float)a) {}
    LB_F2_HD explicit F2(int a) : F2((float)a, (float)a) {}
    LB_F2_HD float lo() const
    {
#ifdef __CUDA_ARCH__
        float a, b;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
        return a;
#else
        const uint32_t a = (uint32_t)r;
        float f; memcpy(&f, &a, 4);
        return f;
#endif
    }
    LB_F2_HD float hi() const
    {
#ifdef __CUDA_ARCH__
        float a, b;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
        return b;
#else
        const uint32_t b = (uint32_t)(r >> 32);
        float f; memcpy(&f, &b, 4);
        return f;
#endif
    }
};

LB_F2_HD F2 operator+(F2 a, F2 b)
{
#ifdef __CUDA_ARCH__
    F2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.r) : "l"(a.r), "l"(b.r)); return d;
#else
    return F2(a.lo() + b.lo(), a.hi() + b.hi());
#endif
}
LB_F2_HD F2 operator-(F2 a, F2 b)
{
#ifdef __CUDA_ARCH__
    F2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.r) : "l"(a.r), "l"(b.r)); return d;
#else
    return F2(a.lo() - b.lo(), a.hi() - b.hi());
#endif
}
// Sums whose operands are PRODUCTS must not use the operators above: ptxas (12.9) contracts `mul.rn.f32x2`
// followed by `add.rn.f32x2` / `sub.rn.f32x2` into one FFMA2 even though both carry an explicit rounding mode
// and the library is built with -fmad=false (the scalar instructions are never contracted under those
// conditions), which makes the packed lanes differ from the scalar code by one ulp -- found by the parity tests
// of round 2, first on the relaxation `f*(1-omega) + omega*feq`.  f2_add_nf / f2_sub_nf issue a + b as a*1 + b
// and a - b as b*(-1) + a with the ones read from constant memory: the product with one is exact, so the FMA
// rounds the sum once, like an addition; an FMA cannot absorb another multiplication, and the constant is
// opaque to the assembler, so nothing can be contracted.  lb_device.cuh uses them (through lb_add / lb_sub)
// wherever an operand of a sum is the direct result of a multiplication.
#ifdef __CUDACC__
static __device__ __constant__ unsigned long long lb_f2_ones[2] = {0x3f8000003f800000ull, 0xbf800000bf800000ull};
#endif
LB_F2_HD F2 f2_add_nf(F2 a, F2 b)
{
#ifdef __CUDA_ARCH__
    F2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.r) : "l"(a.r), "l"(lb_f2_ones[0]), "l"(b.r)); return d;
#else
    return F2(a.lo() + b.lo(), a.hi() + b.hi());
#endif
}
LB_F2_HD F2 f2_sub_nf(F2 a, F2 b)
{
#ifdef __CUDA_ARCH__
    F2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.r) : "l"(b.r), "l"(lb_f2_ones[1]), "l"(a.r)); return d;
#else
    return F2(a.lo() - b.lo(), a.hi() - b.hi());
#endif
}
LB_F2_HD F2 operator*(F2 a, F2 b)
{
#ifdef __CUDA_ARCH__
    F2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.r) : "l"(a.r), "l"(b.r)); return d;
#else
    return F2(a.lo() * b.lo(), a.hi() * b.hi());
#endif
}
// exact negation of both lanes: a sign-bit flip on the integer pipe, which the collision leaves idle
LB_F2_HD F2 operator-(F2 a) { F2 d; d.r = a.r ^ 0x8000000080000000ull; return d; }
LB_F2_HD F2 f2_fma(F2 a, F2 b, F2 c)
{
#ifdef __CUDA_ARCH__
    F2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.r) : "l"(a.r), "l"(b.r), "l"(c.r)); return d;
#else
    return F2(fmaf(a.lo(), b.lo(), c.lo()), fmaf(a.hi(), b.hi(), c.hi()));
#endif
}
// lane-wise MUFU.RCP (there is no packed special-function unit); host: exact quotient, only used where
// the Newton step below it lands on the IEEE result anyway
LB_F2_HD F2 f2_rcp_approx(F2 x)
{
#ifdef __CUDA_ARCH__
    float a, b;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(x.lo()));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(x.hi()));
    return F2(a, b);
#else
    return F2(1.0f / x.lo(), 1.0f / x.hi());
#endif
}

}  // namespace lb
