/*
 * oracle.c -- CPU restatement of SkellySim's hydrodynamic pair-kernel hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product path (skellysim_b200/csrc) never links, loads or calls anything here and
 * fails loudly when its CUDA library is missing.
 *
 * Parity status: the reference holds NO golden vectors for this path (its only kernel
 * test, tests/core/kernel_test.cpp, compares backends against each other on unseeded
 * random input).  The oracle is therefore pinned two ways:
 *   (1) against the reference's own importable numba kernels
 *       (src/skelly_sim/kernels.py:271-321 Stokeslet, :660-692 stresslet) -- fixtures in
 *       tests/golden/ made by tests/golden/make_golden.py inside the build container;
 *   (2) on the GPU box against the reference's own CUDA direct kernels
 *       (src/core/kernels.cu compiled UNMODIFIED into oracle/_ref/, see oracle/Makefile).
 * The reference's CPU direct path (src/core/kernels.cpp:54-83) calls PVFMM, which is an
 * un-vendored third-party dependency (pvfmm v1.3.0, ci/Dockerfile:23-25) absent from
 * /root/reference, so it cannot be compiled here; its algorithm is restated below from
 * the in-tree statements of the same math.
 *
 * Reference statements followed (all paths relative to /root/reference):
 *   Stokeslet pair     src/core/kernels.cu:57-77      r = trg - src, r2==0 -> rinv = 0
 *   stresslet pair     src/core/kernels.cu:24-55  ==  src/core/kernels.cpp:11-40
 *   scale 1/(8 pi)     src/core/kernels.cu:26,59,121-122 (applied after the sum)
 *   overwrite output   src/core/kernels.cu:93-96
 *   /eta in wrapper    src/core/kernels.cpp:358,365 (gpu), :66,:82 (cpu)
 *   OpenMP chunking    src/core/kernels.cpp:42-51,58-65
 *   oseen_tensor       src/core/kernels.cpp:146-195 (regularised self block)
 *   rotlet             src/core/kernels.cpp:206-242
 *   stresslet*n*rho    src/core/kernels.cpp:311-334
 *
 * Layout everywhere: Eigen column-major 3 x n  ==  AoS xyz, r[3*i+k]; stresslet strength
 * 9 x n  ==  f[9*i + 3*a + b].
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

static const double SCALE_8PI = 1.0 / 8.0 / M_PI; /* kernels.cu:26,59 */

/* ------------------------------------------------------------------------------------------
 * Scalar restatements (the checker).  Summation order: sources ascending, one accumulator per
 * target component, exactly like tiled_driver (kernels.cu:106-113).
 * ---------------------------------------------------------------------------------------- */

/* kernels.cu:57-77 (StokesCuda::uKernel) + driver :79-123.  Result includes 1/(8 pi), excludes 1/eta. */
ORACLE_API void oracle_stokeslet_direct(const double *r_src, const double *f_src, int n_src, const double *r_trg,
                                        double *u_trg, int n_trg) {
    for (int it = 0; it < n_trg; ++it) {
        const double tx = r_trg[3 * it + 0], ty = r_trg[3 * it + 1], tz = r_trg[3 * it + 2];
        double u0 = 0.0, u1 = 0.0, u2 = 0.0;
        for (int is = 0; is < n_src; ++is) {
            const double dx = tx - r_src[3 * is + 0];
            const double dy = ty - r_src[3 * is + 1];
            const double dz = tz - r_src[3 * is + 2];
            const double fx = f_src[3 * is + 0], fy = f_src[3 * is + 1], fz = f_src[3 * is + 2];
            const double r2 = dx * dx + dy * dy + dz * dz;
            const double rinv = (r2 == 0.0) ? 0.0 : 1.0 / sqrt(r2);
            const double rinv2 = rinv * rinv;
            const double inner = (fx * dx + fy * dy + fz * dz) * rinv2;
            u0 += rinv * (fx + dx * inner);
            u1 += rinv * (fy + dy * inner);
            u2 += rinv * (fz + dz * inner);
        }
        u_trg[3 * it + 0] = u0 * SCALE_8PI;
        u_trg[3 * it + 1] = u1 * SCALE_8PI;
        u_trg[3 * it + 2] = u2 * SCALE_8PI;
    }
}

/* kernels.cu:24-55 (StokesDoubleLayerCuda::uKernel), same polynomial as kernels.cpp:18-39. */
ORACLE_API void oracle_stresslet_direct(const double *r_src, const double *f_src, int n_src, const double *r_trg,
                                        double *u_trg, int n_trg) {
    for (int it = 0; it < n_trg; ++it) {
        const double tx = r_trg[3 * it + 0], ty = r_trg[3 * it + 1], tz = r_trg[3 * it + 2];
        double u0 = 0.0, u1 = 0.0, u2 = 0.0;
        for (int is = 0; is < n_src; ++is) {
            const double *f = f_src + 9 * (size_t)is;
            const double dx = tx - r_src[3 * is + 0];
            const double dy = ty - r_src[3 * is + 1];
            const double dz = tz - r_src[3 * is + 2];
            const double r2 = dx * dx + dy * dy + dz * dz;
            const double rinv = (r2 == 0.0) ? 0.0 : 1.0 / sqrt(r2);
            const double rinv2 = rinv * rinv;
            const double rinv5 = rinv * rinv2 * rinv2;
            double coeff = f[0] * dx * dx + f[4] * dy * dy + f[8] * dz * dz;
            coeff += (f[1] + f[3]) * dx * dy;
            coeff += (f[2] + f[6]) * dx * dz;
            coeff += (f[5] + f[7]) * dy * dz;
            coeff *= -3.0 * rinv5;
            u0 += dx * coeff;
            u1 += dy * coeff;
            u2 += dz * coeff;
        }
        u_trg[3 * it + 0] = u0 * SCALE_8PI;
        u_trg[3 * it + 1] = u1 * SCALE_8PI;
        u_trg[3 * it + 2] = u2 * SCALE_8PI;
    }
}

/* long-double (x87 80-bit) variants: used only to measure the TRUE error of both the FP64 oracle
 * and the CUDA path. */
ORACLE_API void oracle_stokeslet_direct_ld(const double *r_src, const double *f_src, int n_src, const double *r_trg,
                                           double *u_trg, int n_trg) {
    const long double scale = 1.0L / (8.0L * 3.14159265358979323846264338327950288L);
    for (int it = 0; it < n_trg; ++it) {
        const long double tx = r_trg[3 * it + 0], ty = r_trg[3 * it + 1], tz = r_trg[3 * it + 2];
        long double u0 = 0, u1 = 0, u2 = 0;
        for (int is = 0; is < n_src; ++is) {
            const long double dx = tx - r_src[3 * is + 0];
            const long double dy = ty - r_src[3 * is + 1];
            const long double dz = tz - r_src[3 * is + 2];
            const long double fx = f_src[3 * is + 0], fy = f_src[3 * is + 1], fz = f_src[3 * is + 2];
            const long double r2 = dx * dx + dy * dy + dz * dz;
            const long double rinv = (r2 == 0.0L) ? 0.0L : 1.0L / sqrtl(r2);
            const long double inner = (fx * dx + fy * dy + fz * dz) * rinv * rinv;
            u0 += rinv * (fx + dx * inner);
            u1 += rinv * (fy + dy * inner);
            u2 += rinv * (fz + dz * inner);
        }
        u_trg[3 * it + 0] = (double)(u0 * scale);
        u_trg[3 * it + 1] = (double)(u1 * scale);
        u_trg[3 * it + 2] = (double)(u2 * scale);
    }
}

ORACLE_API void oracle_stresslet_direct_ld(const double *r_src, const double *f_src, int n_src, const double *r_trg,
                                           double *u_trg, int n_trg) {
    const long double scale = 1.0L / (8.0L * 3.14159265358979323846264338327950288L);
    for (int it = 0; it < n_trg; ++it) {
        const long double tx = r_trg[3 * it + 0], ty = r_trg[3 * it + 1], tz = r_trg[3 * it + 2];
        long double u0 = 0, u1 = 0, u2 = 0;
        for (int is = 0; is < n_src; ++is) {
            const double *f = f_src + 9 * (size_t)is;
            const long double dx = tx - r_src[3 * is + 0];
            const long double dy = ty - r_src[3 * is + 1];
            const long double dz = tz - r_src[3 * is + 2];
            const long double r2 = dx * dx + dy * dy + dz * dz;
            const long double rinv = (r2 == 0.0L) ? 0.0L : 1.0L / sqrtl(r2);
            const long double rinv2 = rinv * rinv;
            const long double rinv5 = rinv * rinv2 * rinv2;
            long double coeff = (long double)f[0] * dx * dx + (long double)f[4] * dy * dy + (long double)f[8] * dz * dz;
            coeff += ((long double)f[1] + f[3]) * dx * dy;
            coeff += ((long double)f[2] + f[6]) * dx * dz;
            coeff += ((long double)f[5] + f[7]) * dy * dz;
            coeff *= -3.0L * rinv5;
            u0 += dx * coeff;
            u1 += dy * coeff;
            u2 += dz * coeff;
        }
        u_trg[3 * it + 0] = (double)(u0 * scale);
        u_trg[3 * it + 1] = (double)(u1 * scale);
        u_trg[3 * it + 2] = (double)(u2 * scale);
    }
}

/* ------------------------------------------------------------------------------------------
 * Small dense helpers on the flow() path.
 * ---------------------------------------------------------------------------------------- */

/* kernels.cpp:146-195 oseen_tensor_direct, contracted with a density on the fly:
 *   out[3*i+k] (+)= sum_j G(x_i - x_j)_{kl} rho[3*j+l],  G regularised for 0 < r <= eps, zero for r == 0.
 * This is `fib.stokeslet_ * wf` of fiber_container_finite_difference.cpp:203-210 without forming
 * the 3n x 3n matrix.  sign = +1 adds, -1 subtracts (the self-term subtraction). */
ORACLE_API void oracle_oseen_contract(const double *r_src, int n_src, const double *r_trg, int n_trg,
                                      const double *density, double eta, double reg, double eps, double sign,
                                      double *out) {
    const double factor = 1.0 / (8.0 * M_PI * eta);
    const double reg2 = reg * reg;
    for (int it = 0; it < n_trg; ++it) {
        double a0 = 0, a1 = 0, a2 = 0;
        for (int is = 0; is < n_src; ++is) {
            /* reference sign convention is src - trg (kernels.cpp:161-163); G is even in dr */
            const double dx = r_src[3 * is + 0] - r_trg[3 * it + 0];
            const double dy = r_src[3 * is + 1] - r_trg[3 * it + 1];
            const double dz = r_src[3 * is + 2] - r_trg[3 * it + 2];
            const double dr2 = dx * dx + dy * dy + dz * dz;
            if (dr2 == 0.0)
                continue;
            const double dr = sqrt(dr2);
            double fr, gr;
            if (dr > eps) {
                fr = factor / dr;
                gr = factor / (dr * dr * dr);
            } else {
                const double denom_inv = 1.0 / sqrt(dr * dr + reg2);
                fr = factor * denom_inv;
                gr = factor * denom_inv * denom_inv * denom_inv;
            }
            const double d0 = density[3 * is + 0], d1 = density[3 * is + 1], d2 = density[3 * is + 2];
            const double dot = gr * (dx * d0 + dy * d1 + dz * d2);
            a0 += fr * d0 + dx * dot;
            a1 += fr * d1 + dy * dot;
            a2 += fr * d2 + dz * dot;
        }
        out[3 * it + 0] += sign * a0;
        out[3 * it + 1] += sign * a1;
        out[3 * it + 2] += sign * a2;
    }
}

/* kernels.cpp:206-242 rotlet: u_t += (1/(8 pi eta)) sum_s (L_s x d)/|d|^3, d = trg - src, regularised
 * when |d|^2 < eps^2.  Accumulates into u_trg. */
ORACLE_API void oracle_rotlet_add(const double *r_src, int n_src, const double *r_trg, int n_trg,
                                  const double *density, double eta, double reg, double eps, double *u_trg) {
    const double eps2 = eps * eps, reg2 = reg * reg;
    const double factor = 1.0 / (8.0 * M_PI * eta);
    for (int it = 0; it < n_trg; ++it) {
        double a0 = 0, a1 = 0, a2 = 0;
        for (int is = 0; is < n_src; ++is) {
            const double dx = r_trg[3 * it + 0] - r_src[3 * is + 0];
            const double dy = r_trg[3 * it + 1] - r_src[3 * is + 1];
            const double dz = r_trg[3 * it + 2] - r_src[3 * is + 2];
            const double dr2 = dx * dx + dy * dy + dz * dz;
            const double dr = dr2 < eps2 ? sqrt(reg2 + dr2) : sqrt(dr2);
            const double fr = 1.0 / (dr * dr * dr);
            const double l0 = density[3 * is + 0], l1 = density[3 * is + 1], l2 = density[3 * is + 2];
            a0 += fr * (dz * l1 - dy * l2);
            a1 += fr * (dx * l2 - dz * l0);
            a2 += fr * (dy * l0 - dx * l1);
        }
        u_trg[3 * it + 0] += factor * a0;
        u_trg[3 * it + 1] += factor * a1;
        u_trg[3 * it + 2] += factor * a2;
    }
}

/* periphery.cpp:68-71 / body_container.cpp:296-302: f_dl[9*i + 3*a + b] = 2 eta n_a rho_b */
ORACLE_API void oracle_form_double_layer(const double *normals, const double *density, int n, double eta,
                                         double *f_dl) {
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                f_dl[9 * (size_t)i + 3 * a + b] = 2.0 * eta * normals[3 * i + a] * density[3 * i + b];
}

/* ------------------------------------------------------------------------------------------
 * CPU baseline ("port" of the reference's OpenMP direct path, kernels.cpp:42-83):
 * static target chunks per thread; inside a chunk the PVFMM GenericKernel idiom -- a SIMD vector
 * of targets against broadcast sources, rsqrt estimate + Newton (sctl::approx_rsqrt), mask r2>0.
 * AVX-512 / AVX2 clones are chosen at run time so the same .so runs on the GPU box's host CPU.
 * ---------------------------------------------------------------------------------------- */

/* kernels.cpp:42-51 */
static void get_chunk_start_and_size(int i_thr, int n_thr, int prob_size, int *start, int *size) {
    const int chunk_small = prob_size / n_thr;
    const int chunk_big = chunk_small + 1;
    const int remainder = prob_size % n_thr;
    if (i_thr < remainder) {
        *start = chunk_big * i_thr;
        *size = chunk_big;
    } else {
        *start = remainder * chunk_big + (i_thr - remainder) * chunk_small;
        *size = chunk_small;
    }
}

typedef void (*chunk_fn)(const double *r_src, const double *f_src, int n_src, const double *r_trg, double *u_trg,
                         int n_trg);

#if defined(__x86_64__)
/* ---- AVX-512: 8 targets per vector, 2 vectors in flight ---- */
__attribute__((target("avx512f,avx512dq,fma"))) static inline __m512d rsqrt_masked_512(__m512d r2) {
    const __mmask8 nz = _mm512_cmp_pd_mask(r2, _mm512_setzero_pd(), _CMP_GT_OQ);
    __m512d y = _mm512_maskz_rsqrt14_pd(nz, r2);
    /* two Newton steps: 14 -> 28 -> 56 bits; y <- y*(1.5 - 0.5*r2*y*y) */
    const __m512d half = _mm512_set1_pd(0.5), three_half = _mm512_set1_pd(1.5);
    for (int k = 0; k < 2; ++k) {
        const __m512d t = _mm512_mul_pd(_mm512_mul_pd(r2, y), y);
        y = _mm512_mul_pd(y, _mm512_fnmadd_pd(half, t, three_half));
    }
    /* final correction step in residual form for full double accuracy */
    {
        const __m512d e = _mm512_fnmadd_pd(_mm512_mul_pd(r2, y), y, _mm512_set1_pd(1.0));
        y = _mm512_fmadd_pd(_mm512_mul_pd(y, e), half, y);
    }
    return _mm512_maskz_mov_pd(nz, y);
}

__attribute__((target("avx512f,avx512dq,fma"))) static void stokeslet_chunk_avx512(const double *r_src,
                                                                                   const double *f_src, int n_src,
                                                                                   const double *r_trg,
                                                                                   double *u_trg, int n_trg) {
    const __m512d scale = _mm512_set1_pd(SCALE_8PI);
    for (int t0 = 0; t0 < n_trg; t0 += 16) {
        double tb[3][16], ub[3][16];
        const int nt = (n_trg - t0) < 16 ? (n_trg - t0) : 16;
        for (int i = 0; i < 16; ++i) {
            const int ii = i < nt ? i : nt - 1;
            tb[0][i] = r_trg[3 * (t0 + ii) + 0];
            tb[1][i] = r_trg[3 * (t0 + ii) + 1];
            tb[2][i] = r_trg[3 * (t0 + ii) + 2];
        }
        __m512d tx0 = _mm512_loadu_pd(&tb[0][0]), tx1 = _mm512_loadu_pd(&tb[0][8]);
        __m512d ty0 = _mm512_loadu_pd(&tb[1][0]), ty1 = _mm512_loadu_pd(&tb[1][8]);
        __m512d tz0 = _mm512_loadu_pd(&tb[2][0]), tz1 = _mm512_loadu_pd(&tb[2][8]);
        __m512d ux0 = _mm512_setzero_pd(), uy0 = ux0, uz0 = ux0, ux1 = ux0, uy1 = ux0, uz1 = ux0;
        for (int is = 0; is < n_src; ++is) {
            const __m512d sx = _mm512_set1_pd(r_src[3 * is + 0]), sy = _mm512_set1_pd(r_src[3 * is + 1]),
                          sz = _mm512_set1_pd(r_src[3 * is + 2]);
            const __m512d fx = _mm512_set1_pd(f_src[3 * is + 0]), fy = _mm512_set1_pd(f_src[3 * is + 1]),
                          fz = _mm512_set1_pd(f_src[3 * is + 2]);
#define SL_BODY(TX, TY, TZ, UX, UY, UZ)                                                                             \
    {                                                                                                               \
        const __m512d dx = _mm512_sub_pd(TX, sx), dy = _mm512_sub_pd(TY, sy), dz = _mm512_sub_pd(TZ, sz);           \
        const __m512d r2 = _mm512_fmadd_pd(dz, dz, _mm512_fmadd_pd(dy, dy, _mm512_mul_pd(dx, dx)));                 \
        const __m512d rinv = rsqrt_masked_512(r2);                                                                  \
        const __m512d rinv2 = _mm512_mul_pd(rinv, rinv);                                                            \
        const __m512d fr = _mm512_fmadd_pd(fz, dz, _mm512_fmadd_pd(fy, dy, _mm512_mul_pd(fx, dx)));                 \
        const __m512d ip = _mm512_mul_pd(fr, rinv2);                                                                \
        UX = _mm512_fmadd_pd(rinv, _mm512_fmadd_pd(dx, ip, fx), UX);                                                \
        UY = _mm512_fmadd_pd(rinv, _mm512_fmadd_pd(dy, ip, fy), UY);                                                \
        UZ = _mm512_fmadd_pd(rinv, _mm512_fmadd_pd(dz, ip, fz), UZ);                                                \
    }
            SL_BODY(tx0, ty0, tz0, ux0, uy0, uz0)
            SL_BODY(tx1, ty1, tz1, ux1, uy1, uz1)
#undef SL_BODY
        }
        _mm512_storeu_pd(&ub[0][0], _mm512_mul_pd(ux0, scale));
        _mm512_storeu_pd(&ub[0][8], _mm512_mul_pd(ux1, scale));
        _mm512_storeu_pd(&ub[1][0], _mm512_mul_pd(uy0, scale));
        _mm512_storeu_pd(&ub[1][8], _mm512_mul_pd(uy1, scale));
        _mm512_storeu_pd(&ub[2][0], _mm512_mul_pd(uz0, scale));
        _mm512_storeu_pd(&ub[2][8], _mm512_mul_pd(uz1, scale));
        for (int i = 0; i < nt; ++i) {
            u_trg[3 * (t0 + i) + 0] = ub[0][i];
            u_trg[3 * (t0 + i) + 1] = ub[1][i];
            u_trg[3 * (t0 + i) + 2] = ub[2][i];
        }
    }
}

__attribute__((target("avx512f,avx512dq,fma"))) static void stresslet_chunk_avx512(const double *r_src,
                                                                                   const double *f_src, int n_src,
                                                                                   const double *r_trg,
                                                                                   double *u_trg, int n_trg) {
    const __m512d scale = _mm512_set1_pd(-3.0 * SCALE_8PI);
    for (int t0 = 0; t0 < n_trg; t0 += 16) {
        double tb[3][16], ub[3][16];
        const int nt = (n_trg - t0) < 16 ? (n_trg - t0) : 16;
        for (int i = 0; i < 16; ++i) {
            const int ii = i < nt ? i : nt - 1;
            tb[0][i] = r_trg[3 * (t0 + ii) + 0];
            tb[1][i] = r_trg[3 * (t0 + ii) + 1];
            tb[2][i] = r_trg[3 * (t0 + ii) + 2];
        }
        __m512d tx0 = _mm512_loadu_pd(&tb[0][0]), tx1 = _mm512_loadu_pd(&tb[0][8]);
        __m512d ty0 = _mm512_loadu_pd(&tb[1][0]), ty1 = _mm512_loadu_pd(&tb[1][8]);
        __m512d tz0 = _mm512_loadu_pd(&tb[2][0]), tz1 = _mm512_loadu_pd(&tb[2][8]);
        __m512d ux0 = _mm512_setzero_pd(), uy0 = ux0, uz0 = ux0, ux1 = ux0, uy1 = ux0, uz1 = ux0;
        for (int is = 0; is < n_src; ++is) {
            const double *f = f_src + 9 * (size_t)is;
            const __m512d sx = _mm512_set1_pd(r_src[3 * is + 0]), sy = _mm512_set1_pd(r_src[3 * is + 1]),
                          sz = _mm512_set1_pd(r_src[3 * is + 2]);
            const __m512d sxx = _mm512_set1_pd(f[0]), syy = _mm512_set1_pd(f[4]), szz = _mm512_set1_pd(f[8]);
            const __m512d pxy = _mm512_set1_pd(f[1] + f[3]), pxz = _mm512_set1_pd(f[2] + f[6]),
                          pyz = _mm512_set1_pd(f[5] + f[7]);
#define DL_BODY(TX, TY, TZ, UX, UY, UZ)                                                                             \
    {                                                                                                               \
        const __m512d dx = _mm512_sub_pd(TX, sx), dy = _mm512_sub_pd(TY, sy), dz = _mm512_sub_pd(TZ, sz);           \
        const __m512d r2 = _mm512_fmadd_pd(dz, dz, _mm512_fmadd_pd(dy, dy, _mm512_mul_pd(dx, dx)));                 \
        const __m512d rinv = rsqrt_masked_512(r2);                                                                  \
        const __m512d rinv2 = _mm512_mul_pd(rinv, rinv);                                                            \
        const __m512d rinv5 = _mm512_mul_pd(rinv, _mm512_mul_pd(rinv2, rinv2));                                     \
        __m512d v1 = _mm512_fmadd_pd(pxz, dz, _mm512_fmadd_pd(pxy, dy, _mm512_mul_pd(sxx, dx)));                    \
        __m512d v2 = _mm512_fmadd_pd(pyz, dz, _mm512_mul_pd(syy, dy));                                              \
        __m512d v3 = _mm512_mul_pd(szz, dz);                                                                        \
        __m512d co = _mm512_fmadd_pd(v3, dz, _mm512_fmadd_pd(v2, dy, _mm512_mul_pd(v1, dx)));                       \
        co = _mm512_mul_pd(co, rinv5);                                                                              \
        UX = _mm512_fmadd_pd(dx, co, UX);                                                                           \
        UY = _mm512_fmadd_pd(dy, co, UY);                                                                           \
        UZ = _mm512_fmadd_pd(dz, co, UZ);                                                                           \
    }
            DL_BODY(tx0, ty0, tz0, ux0, uy0, uz0)
            DL_BODY(tx1, ty1, tz1, ux1, uy1, uz1)
#undef DL_BODY
        }
        _mm512_storeu_pd(&ub[0][0], _mm512_mul_pd(ux0, scale));
        _mm512_storeu_pd(&ub[0][8], _mm512_mul_pd(ux1, scale));
        _mm512_storeu_pd(&ub[1][0], _mm512_mul_pd(uy0, scale));
        _mm512_storeu_pd(&ub[1][8], _mm512_mul_pd(uy1, scale));
        _mm512_storeu_pd(&ub[2][0], _mm512_mul_pd(uz0, scale));
        _mm512_storeu_pd(&ub[2][8], _mm512_mul_pd(uz1, scale));
        for (int i = 0; i < nt; ++i) {
            u_trg[3 * (t0 + i) + 0] = ub[0][i];
            u_trg[3 * (t0 + i) + 1] = ub[1][i];
            u_trg[3 * (t0 + i) + 2] = ub[2][i];
        }
    }
}

/* ---- AVX2+FMA: 4 targets per vector, 2 vectors in flight ---- */
__attribute__((target("avx2,fma"))) static inline __m256d rsqrt_masked_256(__m256d r2) {
    const __m256d nz = _mm256_cmp_pd(r2, _mm256_setzero_pd(), _CMP_GT_OQ);
    /* single-precision estimate (12 bits) widened, then Newton: 12 -> 24 -> 48 -> residual step */
    __m256d y = _mm256_cvtps_pd(_mm_rsqrt_ps(_mm256_cvtpd_ps(r2)));
    y = _mm256_and_pd(y, nz); /* r2 == 0 -> estimate inf -> masked to 0 before it can make a NaN */
    const __m256d half = _mm256_set1_pd(0.5), three_half = _mm256_set1_pd(1.5);
    for (int k = 0; k < 2; ++k) {
        const __m256d t = _mm256_mul_pd(_mm256_mul_pd(r2, y), y);
        y = _mm256_mul_pd(y, _mm256_fnmadd_pd(half, t, three_half));
    }
    for (int k = 0; k < 2; ++k) {
        const __m256d e = _mm256_fnmadd_pd(_mm256_mul_pd(r2, y), y, _mm256_set1_pd(1.0));
        y = _mm256_fmadd_pd(_mm256_mul_pd(y, e), half, y);
    }
    return _mm256_and_pd(y, nz);
}

__attribute__((target("avx2,fma"))) static void stokeslet_chunk_avx2(const double *r_src, const double *f_src,
                                                                     int n_src, const double *r_trg, double *u_trg,
                                                                     int n_trg) {
    const __m256d scale = _mm256_set1_pd(SCALE_8PI);
    for (int t0 = 0; t0 < n_trg; t0 += 8) {
        double tb[3][8], ub[3][8];
        const int nt = (n_trg - t0) < 8 ? (n_trg - t0) : 8;
        for (int i = 0; i < 8; ++i) {
            const int ii = i < nt ? i : nt - 1;
            tb[0][i] = r_trg[3 * (t0 + ii) + 0];
            tb[1][i] = r_trg[3 * (t0 + ii) + 1];
            tb[2][i] = r_trg[3 * (t0 + ii) + 2];
        }
        __m256d tx0 = _mm256_loadu_pd(&tb[0][0]), tx1 = _mm256_loadu_pd(&tb[0][4]);
        __m256d ty0 = _mm256_loadu_pd(&tb[1][0]), ty1 = _mm256_loadu_pd(&tb[1][4]);
        __m256d tz0 = _mm256_loadu_pd(&tb[2][0]), tz1 = _mm256_loadu_pd(&tb[2][4]);
        __m256d ux0 = _mm256_setzero_pd(), uy0 = ux0, uz0 = ux0, ux1 = ux0, uy1 = ux0, uz1 = ux0;
        for (int is = 0; is < n_src; ++is) {
            const __m256d sx = _mm256_set1_pd(r_src[3 * is + 0]), sy = _mm256_set1_pd(r_src[3 * is + 1]),
                          sz = _mm256_set1_pd(r_src[3 * is + 2]);
            const __m256d fx = _mm256_set1_pd(f_src[3 * is + 0]), fy = _mm256_set1_pd(f_src[3 * is + 1]),
                          fz = _mm256_set1_pd(f_src[3 * is + 2]);
#define SL_BODY(TX, TY, TZ, UX, UY, UZ)                                                                             \
    {                                                                                                               \
        const __m256d dx = _mm256_sub_pd(TX, sx), dy = _mm256_sub_pd(TY, sy), dz = _mm256_sub_pd(TZ, sz);           \
        const __m256d r2 = _mm256_fmadd_pd(dz, dz, _mm256_fmadd_pd(dy, dy, _mm256_mul_pd(dx, dx)));                 \
        const __m256d rinv = rsqrt_masked_256(r2);                                                                  \
        const __m256d rinv2 = _mm256_mul_pd(rinv, rinv);                                                            \
        const __m256d fr = _mm256_fmadd_pd(fz, dz, _mm256_fmadd_pd(fy, dy, _mm256_mul_pd(fx, dx)));                 \
        const __m256d ip = _mm256_mul_pd(fr, rinv2);                                                                \
        UX = _mm256_fmadd_pd(rinv, _mm256_fmadd_pd(dx, ip, fx), UX);                                                \
        UY = _mm256_fmadd_pd(rinv, _mm256_fmadd_pd(dy, ip, fy), UY);                                                \
        UZ = _mm256_fmadd_pd(rinv, _mm256_fmadd_pd(dz, ip, fz), UZ);                                                \
    }
            SL_BODY(tx0, ty0, tz0, ux0, uy0, uz0)
            SL_BODY(tx1, ty1, tz1, ux1, uy1, uz1)
#undef SL_BODY
        }
        _mm256_storeu_pd(&ub[0][0], _mm256_mul_pd(ux0, scale));
        _mm256_storeu_pd(&ub[0][4], _mm256_mul_pd(ux1, scale));
        _mm256_storeu_pd(&ub[1][0], _mm256_mul_pd(uy0, scale));
        _mm256_storeu_pd(&ub[1][4], _mm256_mul_pd(uy1, scale));
        _mm256_storeu_pd(&ub[2][0], _mm256_mul_pd(uz0, scale));
        _mm256_storeu_pd(&ub[2][4], _mm256_mul_pd(uz1, scale));
        for (int i = 0; i < nt; ++i) {
            u_trg[3 * (t0 + i) + 0] = ub[0][i];
            u_trg[3 * (t0 + i) + 1] = ub[1][i];
            u_trg[3 * (t0 + i) + 2] = ub[2][i];
        }
    }
}

__attribute__((target("avx2,fma"))) static void stresslet_chunk_avx2(const double *r_src, const double *f_src,
                                                                     int n_src, const double *r_trg, double *u_trg,
                                                                     int n_trg) {
    const __m256d scale = _mm256_set1_pd(-3.0 * SCALE_8PI);
    for (int t0 = 0; t0 < n_trg; t0 += 8) {
        double tb[3][8], ub[3][8];
        const int nt = (n_trg - t0) < 8 ? (n_trg - t0) : 8;
        for (int i = 0; i < 8; ++i) {
            const int ii = i < nt ? i : nt - 1;
            tb[0][i] = r_trg[3 * (t0 + ii) + 0];
            tb[1][i] = r_trg[3 * (t0 + ii) + 1];
            tb[2][i] = r_trg[3 * (t0 + ii) + 2];
        }
        __m256d tx0 = _mm256_loadu_pd(&tb[0][0]), tx1 = _mm256_loadu_pd(&tb[0][4]);
        __m256d ty0 = _mm256_loadu_pd(&tb[1][0]), ty1 = _mm256_loadu_pd(&tb[1][4]);
        __m256d tz0 = _mm256_loadu_pd(&tb[2][0]), tz1 = _mm256_loadu_pd(&tb[2][4]);
        __m256d ux0 = _mm256_setzero_pd(), uy0 = ux0, uz0 = ux0, ux1 = ux0, uy1 = ux0, uz1 = ux0;
        for (int is = 0; is < n_src; ++is) {
            const double *f = f_src + 9 * (size_t)is;
            const __m256d sx = _mm256_set1_pd(r_src[3 * is + 0]), sy = _mm256_set1_pd(r_src[3 * is + 1]),
                          sz = _mm256_set1_pd(r_src[3 * is + 2]);
            const __m256d sxx = _mm256_set1_pd(f[0]), syy = _mm256_set1_pd(f[4]), szz = _mm256_set1_pd(f[8]);
            const __m256d pxy = _mm256_set1_pd(f[1] + f[3]), pxz = _mm256_set1_pd(f[2] + f[6]),
                          pyz = _mm256_set1_pd(f[5] + f[7]);
#define DL_BODY(TX, TY, TZ, UX, UY, UZ)                                                                             \
    {                                                                                                               \
        const __m256d dx = _mm256_sub_pd(TX, sx), dy = _mm256_sub_pd(TY, sy), dz = _mm256_sub_pd(TZ, sz);           \
        const __m256d r2 = _mm256_fmadd_pd(dz, dz, _mm256_fmadd_pd(dy, dy, _mm256_mul_pd(dx, dx)));                 \
        const __m256d rinv = rsqrt_masked_256(r2);                                                                  \
        const __m256d rinv2 = _mm256_mul_pd(rinv, rinv);                                                            \
        const __m256d rinv5 = _mm256_mul_pd(rinv, _mm256_mul_pd(rinv2, rinv2));                                     \
        __m256d v1 = _mm256_fmadd_pd(pxz, dz, _mm256_fmadd_pd(pxy, dy, _mm256_mul_pd(sxx, dx)));                    \
        __m256d v2 = _mm256_fmadd_pd(pyz, dz, _mm256_mul_pd(syy, dy));                                              \
        __m256d v3 = _mm256_mul_pd(szz, dz);                                                                        \
        __m256d co = _mm256_fmadd_pd(v3, dz, _mm256_fmadd_pd(v2, dy, _mm256_mul_pd(v1, dx)));                       \
        co = _mm256_mul_pd(co, rinv5);                                                                              \
        UX = _mm256_fmadd_pd(dx, co, UX);                                                                           \
        UY = _mm256_fmadd_pd(dy, co, UY);                                                                           \
        UZ = _mm256_fmadd_pd(dz, co, UZ);                                                                           \
    }
            DL_BODY(tx0, ty0, tz0, ux0, uy0, uz0)
            DL_BODY(tx1, ty1, tz1, ux1, uy1, uz1)
#undef DL_BODY
        }
        _mm256_storeu_pd(&ub[0][0], _mm256_mul_pd(ux0, scale));
        _mm256_storeu_pd(&ub[0][4], _mm256_mul_pd(ux1, scale));
        _mm256_storeu_pd(&ub[1][0], _mm256_mul_pd(uy0, scale));
        _mm256_storeu_pd(&ub[1][4], _mm256_mul_pd(uy1, scale));
        _mm256_storeu_pd(&ub[2][0], _mm256_mul_pd(uz0, scale));
        _mm256_storeu_pd(&ub[2][4], _mm256_mul_pd(uz1, scale));
        for (int i = 0; i < nt; ++i) {
            u_trg[3 * (t0 + i) + 0] = ub[0][i];
            u_trg[3 * (t0 + i) + 1] = ub[1][i];
            u_trg[3 * (t0 + i) + 2] = ub[2][i];
        }
    }
}
#endif /* __x86_64__ */

/* 0 = scalar, 1 = avx2+fma, 2 = avx512 */
ORACLE_API int oracle_simd_level(void) {
#if defined(__x86_64__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq"))
        return 2;
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma"))
        return 1;
#endif
    return 0;
}

ORACLE_API int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static void direct_cpu_driver(chunk_fn fn, const double *r_src, const double *f_src, int n_src, const double *r_trg,
                              double *u_trg, int n_trg, double eta, int n_threads) {
#ifdef _OPENMP
    if (n_threads <= 0)
        n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    /* kernels.cpp:58-65: one static chunk of targets per thread */
#pragma omp parallel for schedule(static) num_threads(n_threads)
    for (int i = 0; i < n_threads; ++i) {
        int start, size;
        get_chunk_start_and_size(i, n_threads, n_trg, &start, &size);
        if (size > 0)
            fn(r_src, f_src, n_src, r_trg + 3 * (size_t)start, u_trg + 3 * (size_t)start, size);
    }
    /* kernels.cpp:66 `return u_trg / eta` */
    const double inv_eta = 1.0 / eta;
    for (size_t i = 0; i < 3 * (size_t)n_trg; ++i)
        u_trg[i] *= inv_eta;
}

/* mirrors kernels::stokeslet_direct_cpu(r_sl, r_dl, r_trg, f_sl, f_dl, eta), kernels.cpp:54-67.
 * simd: -1 = best available, 0 = scalar oracle loop, 1 = avx2, 2 = avx512. */
ORACLE_API int oracle_stokeslet_direct_cpu(const double *r_src, const double *f_src, int n_src, const double *r_trg,
                                           double *u_trg, int n_trg, double eta, int n_threads, int simd) {
    chunk_fn fn = oracle_stokeslet_direct;
    if (simd < 0)
        simd = oracle_simd_level();
#if defined(__x86_64__)
    if (simd > oracle_simd_level())
        return -1;
    if (simd == 2)
        fn = stokeslet_chunk_avx512;
    else if (simd == 1)
        fn = stokeslet_chunk_avx2;
#endif
    direct_cpu_driver(fn, r_src, f_src, n_src, r_trg, u_trg, n_trg, eta, n_threads);
    return simd;
}

/* mirrors kernels::stresslet_direct_cpu, kernels.cpp:69-83 */
ORACLE_API int oracle_stresslet_direct_cpu(const double *r_src, const double *f_src, int n_src, const double *r_trg,
                                           double *u_trg, int n_trg, double eta, int n_threads, int simd) {
    chunk_fn fn = oracle_stresslet_direct;
    if (simd < 0)
        simd = oracle_simd_level();
#if defined(__x86_64__)
    if (simd > oracle_simd_level())
        return -1;
    if (simd == 2)
        fn = stresslet_chunk_avx512;
    else if (simd == 1)
        fn = stresslet_chunk_avx2;
#endif
    direct_cpu_driver(fn, r_src, f_src, n_src, r_trg, u_trg, n_trg, eta, n_threads);
    return simd;
}
