// csrc/kernels.cu -- per-frame CUDA kernels of the feather / no-blend stitch path (sm_100a).
//
// Arithmetic contracts (bit-exact against the reference's CPU functions, see oracle/):
//   colour   : imgproc/src/color.cpp:6087-6169 (YUV->RGB), :6430-6481 (RGB->YUV 4:2:0)
//   bilinear : imgproc/src/imgwarp.cpp:4383-4442 + :3812-4020 -> (sum S_k a_k b_k + 512) >> 10 with
//              a,b in {32-f, f}; identical to ((sum S_k w_k) + 2^14) >> 15 with the 15-bit table
//   gain     : core/src/arithm.cpp multiply-by-scalar in f64 -> sat_u8(rint(v*g))
//   feather  : stitching/src/cuda/blender.cu:73-98 (short)(v*W) truncation, blenders.cpp:581 (x 1/N, rint)
#include "kernels.cuh"
#include <cstdlib>
#include "device_common.cuh"

namespace ob {

// ------------------------------------------------------------------------------------------------
// K_convert
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t vignette_rgbx(uint32_t p, float k)
{
    // cudaarithm mul_mat.cu:198-213 : saturate_cast<uchar>(u8 * f32), round-to-nearest-even
    const int r = clamp255(__float2int_rn((float)(p & 255u) * k));
    const int g = clamp255(__float2int_rn((float)((p >> 8) & 255u) * k));
    const int b = clamp255(__float2int_rn((float)((p >> 16) & 255u) * k));
    return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
}

// L2 residency hints: the RGBX planes (98 MB for the 6 x 2.7K rig) are written here and gathered by
// K_blend right after -- keep them in the 126 MB L2 (evict_last); the table stream that K_blend reads
// once per frame is marked evict_first so it does not push them out.
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_v4_hint(uint32_t* ptr, uint4 v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

__device__ __forceinline__ uint32_t yuv_px(uint32_t Y, int ruv, int guv, int buv)
{
    const int yy = max((int)Y - 16, 0) * 1220542;
    return (uint32_t)__vimin_s32_relu((yy + ruv) >> 20, 255) | ((uint32_t)__vimin_s32_relu((yy + guv) >> 20, 255) << 8) |
           ((uint32_t)__vimin_s32_relu((yy + buv) >> 20, 255) << 16);
}

// one converted source pixel straight from the input planes (same arithmetic as K_convert) -- used by the gain kernel
__device__ __forceinline__ uint32_t source_px(const CamSrc& c, int x, int y)
{
    if (c.rgb) {                                            // packed RGB24 / BGR24: no colour conversion
        const size_t o = (size_t)y * c.y_pitch + 3 * (size_t)x;
        uint32_t px = (uint32_t)__ldg(c.y + o) | ((uint32_t)__ldg(c.u + o) << 8) | ((uint32_t)__ldg(c.v + o) << 16);
        if (c.vignette) px = vignette_rgbx(px, __ldg(c.vignette + (size_t)y * c.w + x));
        return px;
    }
    const int Y = __ldg(c.y + (size_t)y * c.y_pitch + x);
    const int u = (int)__ldg(c.u + (size_t)(y >> 1) * c.u_pitch + (size_t)(x >> 1) * c.uv_step) - 128;
    const int v = (int)__ldg(c.v + (size_t)(y >> 1) * c.v_pitch + (size_t)(x >> 1) * c.uv_step) - 128;
    uint32_t px = yuv_px((uint32_t)Y, (1 << 19) + 1673527 * v, (1 << 19) - 852492 * v - 409993 * u, (1 << 19) + 2116026 * u);
    if (c.vignette) px = vignette_rgbx(px, __ldg(c.vignette + (size_t)y * c.w + x));
    return px;
}

// one thread: 8 px x 2 rows (four chroma samples).  CTA = 32 x 8 threads = 256 x 16 px.
// grid = (ceil(max_w/256), ceil(max_h/16), cameras); CTAs outside a smaller camera exit at once.
__device__ __forceinline__ void convert_body(const ConvertParams& p, int block)
{
    // block -> (column block, row block, camera) of the (grid_x, grid_y, n) convert grid
    const int per_cam = p.grid_x * p.grid_y;
    const int cam = block / per_cam, rem = block - cam * per_cam;
    const int by = rem / p.grid_x, bx = rem - by * p.grid_x;
    const CamSrc& c = p.cam[cam];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x0 = c.col0 + (bx << 8) + (tx << 3);
    const int y0 = c.row0 + (by << 4) + (ty << 1);
    if (x0 >= c.col1 || y0 >= c.row1) return;
    const int w = c.w;
    if (c.rgb) {                                            // packed RGB24 / BGR24 -> RGBX: a byte shuffle (+ vignette)
        #pragma unroll
        for (int r = 0; r < 2; r++) {
            uint32_t px[8];
            #pragma unroll
            for (int k = 0; k < 8; k++) {
                const size_t o = (size_t)(y0 + r) * c.y_pitch + 3 * (size_t)min(x0 + k, w - 1);
                px[k] = (uint32_t)__ldg(c.y + o) | ((uint32_t)__ldg(c.u + o) << 8) | ((uint32_t)__ldg(c.v + o) << 16);
                if (c.vignette) px[k] = vignette_rgbx(px[k], __ldg(c.vignette + (size_t)(y0 + r) * w + min(x0 + k, w - 1)));
            }
            uint32_t* o = c.rgbx + (size_t)(y0 + r) * w + x0;
            if (x0 + 8 <= w && w % 4 == 0) {
                const uint64_t pol = policy_evict_last();
                st_v4_hint(o, make_uint4(px[0], px[1], px[2], px[3]), pol);
                st_v4_hint(o + 4, make_uint4(px[4], px[5], px[6], px[7]), pol);
            } else
                for (int k = 0; k < 8 && x0 + k < w; k++) o[k] = px[k];
        }
        return;
    }
    const uint8_t* yr0 = c.y + (size_t)y0 * c.y_pitch + x0;
    const uint8_t* yr1 = yr0 + c.y_pitch;
    const uint8_t* ur = c.u + (size_t)(y0 >> 1) * c.u_pitch + (size_t)(x0 >> 1) * c.uv_step;
    const uint8_t* vr = c.v + (size_t)(y0 >> 1) * c.v_pitch + (size_t)(x0 >> 1) * c.uv_step;
    uint32_t* o0 = c.rgbx + (size_t)y0 * w + x0;
    uint32_t* o1 = o0 + w;
    const float* vg = c.vignette;

    if (c.aligned4 && x0 + 8 <= w) {
        const uint2 ya = __ldg(reinterpret_cast<const uint2*>(yr0));
        const uint2 yb = __ldg(reinterpret_cast<const uint2*>(yr1));
        uint32_t ub, vb;                                   // four U and four V samples
        if (c.uv_step == 1) {
            ub = __ldg(reinterpret_cast<const uint32_t*>(ur));
            vb = __ldg(reinterpret_cast<const uint32_t*>(vr));
        } else {                                           // NV12: U0 V0 U1 V1 U2 V2 U3 V3
            const uint2 uv = __ldg(reinterpret_cast<const uint2*>(ur));
            ub = __byte_perm(uv.x, uv.y, 0x6420);
            vb = __byte_perm(uv.x, uv.y, 0x7531);
        }
        uint32_t a[8], b[8];
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            const int u = (int)((ub >> (8 * k)) & 255u) - 128, v = (int)((vb >> (8 * k)) & 255u) - 128;
            const int ruv = (1 << 19) + 1673527 * v;
            const int guv = (1 << 19) - 852492 * v - 409993 * u;
            const int buv = (1 << 19) + 2116026 * u;
            const uint32_t wa = k < 2 ? ya.x : ya.y, wb = k < 2 ? yb.x : yb.y;
            const int sh = 16 * (k & 1);
            a[2 * k] = yuv_px((wa >> sh) & 255u, ruv, guv, buv);
            a[2 * k + 1] = yuv_px((wa >> (sh + 8)) & 255u, ruv, guv, buv);
            b[2 * k] = yuv_px((wb >> sh) & 255u, ruv, guv, buv);
            b[2 * k + 1] = yuv_px((wb >> (sh + 8)) & 255u, ruv, guv, buv);
        }
        if (vg) {
            #pragma unroll
            for (int k = 0; k < 8; k++) {
                a[k] = vignette_rgbx(a[k], __ldg(vg + (size_t)y0 * w + x0 + k));
                b[k] = vignette_rgbx(b[k], __ldg(vg + (size_t)(y0 + 1) * w + x0 + k));
            }
        }
        const uint64_t pol = policy_evict_last();
        st_v4_hint(o0, make_uint4(a[0], a[1], a[2], a[3]), pol);
        st_v4_hint(o0 + 4, make_uint4(a[4], a[5], a[6], a[7]), pol);
        st_v4_hint(o1, make_uint4(b[0], b[1], b[2], b[3]), pol);
        st_v4_hint(o1 + 4, make_uint4(b[4], b[5], b[6], b[7]), pol);
    } else {
        for (int k = 0; k < 8 && x0 + k < w; k++) {
            const int u = (int)__ldg(ur + (k >> 1) * c.uv_step) - 128, v = (int)__ldg(vr + (k >> 1) * c.uv_step) - 128;
            const int ruv = (1 << 19) + 1673527 * v;
            const int guv = (1 << 19) - 852492 * v - 409993 * u;
            const int buv = (1 << 19) + 2116026 * u;
            uint32_t a = yuv_px(__ldg(yr0 + k), ruv, guv, buv), b = yuv_px(__ldg(yr1 + k), ruv, guv, buv);
            if (vg) {
                a = vignette_rgbx(a, __ldg(vg + (size_t)y0 * w + x0 + k));
                b = vignette_rgbx(b, __ldg(vg + (size_t)(y0 + 1) * w + x0 + k));
            }
            o0[k] = a; o1[k] = b;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// K_gain
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// exposure_compensate.cpp:138-153 (normal equations) + cv::solve: closed forms for n <= 3 (core/src/lapack.cpp:
// 1080-1170), LU with partial pivoting otherwise (matrix_decomp.cpp:52-109).  Executed by ONE WARP: lane = row
// while building, lane = column during elimination.  Multiplies and adds are kept separate (__dmul_rn /
// __dadd_rn, no FMA contraction) so every element sees the same operation sequence as the x86 reference.
// Aug is the n x (n+1) augmented matrix [A | b] in shared memory.
__device__ void gain_solve_warp(int n, const double* Nm, const double* Im, double* Aug, double* g)
{
    const int lane = threadIdx.x & 31, ld = n + 1;
    const double alpha = 0.01, beta = 100;
    if (lane < n) {
        const int i = lane;
        double bi = 0, aii = 0;
        for (int j = 0; j < n; j++) {
            const double nij = Nm[i * n + j];
            bi = __dadd_rn(bi, __dmul_rn(beta, nij));
            aii = __dadd_rn(aii, __dmul_rn(beta, nij));
            double aij = 0;
            if (j != i) {
                const double iij = Im[i * n + j], iji = Im[j * n + i];
                aii = __dadd_rn(aii, __dmul_rn(__dmul_rn(__dmul_rn(2 * alpha, iij), iij), nij));
                aij = -__dmul_rn(__dmul_rn(__dmul_rn(2 * alpha, iij), iji), nij);
                Aug[i * ld + j] = aij;
            }
        }
        Aug[i * ld + i] = aii;
        Aug[i * ld + n] = bi;
    }
    __syncwarp();
    #define S(i, j) Aug[(i) * ld + (j)]
    #define Bv(i) Aug[(i) * ld + n]
    #define MUL __dmul_rn
    #define SUB(a, b) __dadd_rn((a), -(b))
    #define ADD __dadd_rn
    if (n == 1) { if (lane == 0) g[0] = Bv(0) / S(0, 0); return; }
    if (n == 2) {
        if (lane == 0) {
            double d = SUB(MUL(S(0,0), S(1,1)), MUL(S(0,1), S(1,0)));
            d = 1. / d;
            g[0] = MUL(SUB(MUL(Bv(0), S(1,1)), MUL(Bv(1), S(0,1))), d);
            g[1] = MUL(SUB(MUL(Bv(1), S(0,0)), MUL(Bv(0), S(1,0))), d);
        }
        return;
    }
    if (n == 3) {
        if (lane == 0) {
            double d = ADD(SUB(MUL(S(0,0), SUB(MUL(S(1,1), S(2,2)), MUL(S(1,2), S(2,1)))),
                               MUL(S(0,1), SUB(MUL(S(1,0), S(2,2)), MUL(S(1,2), S(2,0))))),
                           MUL(S(0,2), SUB(MUL(S(1,0), S(2,1)), MUL(S(1,1), S(2,0)))));
            d = 1. / d;
            g[0] = MUL(ADD(ADD(MUL(SUB(MUL(S(1,1), S(2,2)), MUL(S(1,2), S(2,1))), Bv(0)), MUL(SUB(MUL(S(0,2), S(2,1)), MUL(S(0,1), S(2,2))), Bv(1))),
                           MUL(SUB(MUL(S(0,1), S(1,2)), MUL(S(0,2), S(1,1))), Bv(2))), d);
            g[1] = MUL(ADD(ADD(MUL(SUB(MUL(S(1,2), S(2,0)), MUL(S(1,0), S(2,2))), Bv(0)), MUL(SUB(MUL(S(0,0), S(2,2)), MUL(S(0,2), S(2,0))), Bv(1))),
                           MUL(SUB(MUL(S(0,2), S(1,0)), MUL(S(0,0), S(1,2))), Bv(2))), d);
            g[2] = MUL(ADD(ADD(MUL(SUB(MUL(S(1,0), S(2,1)), MUL(S(1,1), S(2,0))), Bv(0)), MUL(SUB(MUL(S(0,1), S(2,0)), MUL(S(0,0), S(2,1))), Bv(1))),
                           MUL(SUB(MUL(S(0,0), S(1,1)), MUL(S(0,1), S(1,0))), Bv(2))), d);
        }
        return;
    }
    // LU with partial pivoting by ONE thread: every element sees the operations of matrix_decomp.cpp:52-109 in the same
    // order (row updates are independent across columns), and the matrix is small enough that the dependent chain of
    // f64 operations, not parallelism, sets the time -- no warp barriers or shared-memory round trips between steps
    if (lane == 0) {
        for (int i = 0; i < n; i++) {
            int k = i;
            for (int j = i + 1; j < n; j++) if (fabs(S(j, i)) > fabs(S(k, i))) k = j;
            if (k != i) for (int c = i; c <= n; c++) { const double t = S(i, c); S(i, c) = S(k, c); S(k, c) = t; }
            const double d = -1 / S(i, i);
            for (int j = i + 1; j < n; j++) {
                const double al = MUL(S(j, i), d);
                for (int c = i + 1; c <= n; c++) S(j, c) = ADD(S(j, c), MUL(al, S(i, c)));
            }
            S(i, i) = -d;
        }
        for (int i = n - 1; i >= 0; i--) {
            double sacc = Bv(i);
            for (int k = i + 1; k < n; k++) sacc = SUB(sacc, MUL(S(i, k), Bv(k)));
            Bv(i) = MUL(sacc, S(i, i));
        }
        for (int i = 0; i < n; i++) g[i] = Bv(i);
    }
    #undef S
    #undef Bv
    #undef MUL
    #undef SUB
    #undef ADD
}

// The same solve for n > 3 by ONE WARP with the matrix in shared memory (lane = column of [A | b]): no CTA barriers, the
// rows of an elimination step are independent so their loads overlap.  Every element sees exactly the operations of
// matrix_decomp.cpp:52-109 in the same order (alpha = A[j][i] * d, A[j][k] += alpha * A[i][k]; no FMA contraction).
__device__ void gain_solve_lu_warp(int n, const double* Nm, const double* Im, double* Aug, double* g)
{
    const int lane = threadIdx.x & 31, ld = n + 1;
    const double alpha = 0.01, beta = 100;
    if (lane < n) {                                         // normal equations: lane = row (exposure_compensate.cpp:138-153)
        const int i = lane;
        double bi = 0, aii = 0;
        #pragma unroll 1
        for (int j = 0; j < n; j++) {
            const double nij = Nm[i * n + j];
            bi = __dadd_rn(bi, __dmul_rn(beta, nij));
            aii = __dadd_rn(aii, __dmul_rn(beta, nij));
            if (j != i) {
                const double iij = Im[i * n + j], iji = Im[j * n + i];
                aii = __dadd_rn(aii, __dmul_rn(__dmul_rn(__dmul_rn(2 * alpha, iij), iij), nij));
                Aug[i * ld + j] = -__dmul_rn(__dmul_rn(__dmul_rn(2 * alpha, iij), iji), nij);
            }
        }
        Aug[i * ld + i] = aii; Aug[i * ld + n] = bi;
    }
    __syncwarp();
    #pragma unroll 1
    for (int i = 0; i < n; i++) {
        // pivot: the first row k >= i with the largest |A[k][i]| (the reference's strict '>' scan keeps the earliest)
        double pv = (lane >= i && lane < n) ? fabs(Aug[lane * ld + i]) : -1.0;
        int k = lane;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, pv, o);
            const int ok = __shfl_xor_sync(0xffffffffu, k, o);
            if (ov > pv || (ov == pv && ok < k)) { pv = ov; k = ok; }
        }
        if (k != i && lane >= i && lane <= n) {             // swap rows i and k from column i on (and b)
            const double t0 = Aug[i * ld + lane], t1 = Aug[k * ld + lane];
            Aug[i * ld + lane] = t1; Aug[k * ld + lane] = t0;
        }
        __syncwarp();
        const double d = -1 / Aug[i * ld + i];
        if (lane > i && lane <= n) {
            const double pivot_c = Aug[i * ld + lane];
            #pragma unroll 4
            for (int r = i + 1; r < n; r++) {
                const double al = __dmul_rn(Aug[r * ld + i], d);
                Aug[r * ld + lane] = __dadd_rn(Aug[r * ld + lane], __dmul_rn(al, pivot_c));
            }
        }
        __syncwarp();
        if (lane == 0) Aug[i * ld + i] = -d;
        __syncwarp();
    }
    if (lane == 0) {                                        // back substitution: a serial chain by definition
        #pragma unroll 1
        for (int i = n - 1; i >= 0; i--) {
            double sacc = Aug[i * ld + n];
            #pragma unroll 1
            for (int k = i + 1; k < n; k++) sacc = __dadd_rn(sacc, -__dmul_rn(Aug[i * ld + k], Aug[k * ld + n]));
            Aug[i * ld + n] = __dmul_rn(sacc, Aug[i * ld + i]);
        }
        #pragma unroll 1
        for (int i = 0; i < n; i++) g[i] = Aug[i * ld + n];
    }
}

// Per camera: the exact u8 gain LUT in f64, and an f32 multiplier for which the single-FMA formula of
// gain_apply_f32() reproduces that LUT for all 256 inputs (searched within +-2 ulp of (float)g).
// If none exists the camera is flagged and the blend kernel reads the LUT instead.  256 threads = 256 inputs.
// sg: the gains in shared memory (just solved by this CTA), or null: read them from p.gains (predefined / shared gains)
__device__ void gain_tables(const GainParams& p, const double* sg)
{
    __shared__ unsigned int s_ok[MAX_CAMS];
    __shared__ unsigned int s_bad;                           // cameras for which (float)g itself does not reproduce the f64 rule
    const int v = threadIdx.x;                               // called by threads 0..255 only
    if (v < MAX_CAMS) s_ok[v] = 0x1Fu;
    if (v == 0) s_bad = 0u;
    asm volatile("bar.sync 1, 256;");
    // candidate k of a camera: (float)g + {0,+1,-1,+2,-2} ulp; it is good for input v if all three kernel forms give the f64 result
    auto good = [&](float g0, int k, int exact) {
        const int step = (k == 0) ? 0 : (k & 1) ? (k + 1) / 2 : -(k / 2);
        const float gc = __int_as_float(__float_as_int(g0) + step);
        return (int)gain_apply_f32((float)v, gc) == exact && (int)gain_apply_biased(MAGIC_RD + (float)v, gc, gain_bias_f32(gc)) == exact &&
               (int)gain_apply_two23((float)v, gc) == exact;
    };
    // pass 1, no barrier inside (the loads of all cameras overlap): exact LUT + candidate 0.  Candidate 0 almost always works.
    unsigned int bad = 0u;
    #pragma unroll 4
    for (int c = 0; c < p.n; c++) {
        const double g = sg ? sg[c] : __ldcg(p.gains + c);
        const int exact = clamp255(__double2int_rn((double)v * g));
        p.gain_lut[c * 256 + v] = (uint8_t)exact;
        if (!(g > 0. && g < 4096.) || !good((float)g, 0, exact)) bad |= 1u << c;
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if ((v & 31) == 0 && bad) atomicOr(&s_bad, bad);
    asm volatile("bar.sync 1, 256;");
    bad = s_bad;
    if (bad) {                                               // rare: search the other candidates for those cameras (uniform branch)
        for (int c = 0; c < p.n; c++) {
            if (!((bad >> c) & 1u)) continue;
            const double g = sg ? sg[c] : __ldcg(p.gains + c);
            const int exact = clamp255(__double2int_rn((double)v * g));
            unsigned int ok = 0u;
            if (g > 0. && g < 4096.) {
                #pragma unroll 1
                for (int k = 0; k < 5; k++) if (good((float)g, k, exact)) ok |= 1u << k;
            }
            if (ok != 0x1Fu) atomicAnd(&s_ok[c], ok);
        }
        asm volatile("bar.sync 1, 256;");
    }
    if (v < p.n) {
        const unsigned int ok = s_ok[v];
        const int k = ok ? __ffs(ok) - 1 : 0;
        const int step = (k == 0) ? 0 : (k & 1) ? (k + 1) / 2 : -(k / 2);
        p.gain_f32[v] = __int_as_float(__float_as_int((float)(sg ? sg[v] : __ldcg(p.gains + v))) + step);
        p.gain_flag[v] = ok ? 0 : 1;
    }
}

// the four taps of a gain sample one by one (vignette maps, packed RGB input); kept out of line so that its registers do
// not weigh on the common path
__device__ __noinline__ uint4 gain_taps_slow(const CamSrc& sc, uint32_t ex, uint32_t ey)
{
    uint32_t t[4] = { 0u, 0u, 0u, 0u };
    const int ix = (int)(ex & 0xFFFFu) - 1, iy = (int)(ex >> 16) - 1;
    const uint32_t bt = (ey & C_BORDER) ? (ey >> C_TAP_SHIFT) : 15u;
    if (bt & 1u) t[0] = source_px(sc, ix, iy);
    if (bt & 2u) t[1] = source_px(sc, ix + 1, iy);
    if (bt & 4u) t[2] = source_px(sc, ix, iy + 1);
    if (bt & 8u) t[3] = source_px(sc, ix + 1, iy + 1);
    return make_uint4(t[0], t[1], t[2], t[3]);
}

// Working-scale statistics (mapper.cpp:94-99: a ~0.1 Mpix canvas) + gain solve in ONE launch, written for latency.
// A CTA takes one chunk of the canvas: up to 256 pixels and the (at most 512) samples (pixel, camera) whose
// working-scale mask is 255 there (CPU compensator's intersect rule, exposure_compensate.cpp:71-78,112), listed by the
// host.  Phase A: thread = two samples; it remaps each pixel (nearest-resized position, mapper.cpp:235-237) straight
// from the input planes and stores ||rgb||_2 (as exact 2^-52 fixed point) in shared memory -- every table and tap load of the chunk is in
// flight at once.  Phase B: warp = camera pair, lanes = pixels.  The sums are accumulated EXACTLY: a norm is sqrt of
// an integer >= 1 (or 0), i.e. an integer multiple of 2^-52 below 2^9, so its 61-bit fixed-point image is split into
// two 64-bit integer accumulators (high / low 32 bits) that cannot overflow over 2^17 samples.  Integer addition is
// associative: atomics in any order give the same totals, no per-CTA partial rows, no reduction pass, and the totals
// are the exact sums rounded to f64 once (the reference's sequential f64 sum differs from that by its own rounding).
// The last CTA to finish (ticket) solves for the gains (one warp) and builds the gain tables.  Deterministic.
constexpr int MAX_PAIRS = MAX_CAMS * (MAX_CAMS + 1) / 2;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
constexpr int GAIN_BLK = 256;                              // CTA size of the fused kernel
constexpr int GAIN_SPT = 2;                                // samples per thread: a chunk holds up to GAIN_BLK * GAIN_SPT samples ...
constexpr int GAIN_PX = 256;                               // ... of up to GAIN_PX canvas pixels (mapper.cpp builds the chunks with the same numbers)
constexpr unsigned long long GAIN_NONE = ~0ull;
// s_nrm: [n][GAIN_PX] fixed-point norms (2^52 x ||rgb||_2 < 2^61, exact), GAIN_NONE where a camera has no sample
__device__ __forceinline__ void gain_body(const GainParams& p, unsigned long long* s_nrm, const unsigned gain_blocks, const unsigned chunk)
{
    __shared__ double s_part[3 * MAX_PAIRS];
    __shared__ double Nm[MAX_CAMS * MAX_CAMS], Im[MAX_CAMS * MAX_CAMS], Aug[MAX_CAMS * (MAX_CAMS + 1)];
    __shared__ uint8_t s_pi[MAX_PAIRS], s_pj[MAX_PAIRS];
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n, np = p.n_pairs;
    const unsigned long long T0 = gtime();
    if (chunk == 0 && tid == 0 && p.dbg) p.dbg[5] = T0;      // diagnostics: when the first gain CTA started
    unsigned long long* const tr = (p.dbg_trace && tid == 0 && chunk < 1024u) ? p.dbg_trace + 8 + 6 * chunk : nullptr;   // phase stamps (OCTVR_GAIN_TRACE)
    if (tr) tr[0] = T0;
    uint4 sm[GAIN_SPT];                                      // entry.x, entry.y, camera | local pixel << 8 (all ones: empty slot)
    #pragma unroll
    for (int h = 0; h < GAIN_SPT; h++) sm[h] = __ldg(p.samples + ((size_t)chunk * GAIN_SPT + h) * GAIN_BLK + tid);
    for (int q = tid; q < n * GAIN_PX; q += GAIN_BLK) s_nrm[q] = GAIN_NONE;
    if (tid == 0) { int q = 0; for (int i = 0; i < n; i++) for (int j = i; j < n; j++, q++) { s_pi[q] = (uint8_t)i; s_pj[q] = (uint8_t)j; } }
    __syncthreads();
    if (tr) tr[1] = gtime() + (sm[0].x & 0u);              // (depends on the sample load)
    {
        uint32_t t[GAIN_SPT][4];
        // Every tap load of both samples is issued before the first one is used (24 byte loads in flight: one memory
        // round trip instead of one per tap -- a load inside a conditional block is followed by its use, which stalls
        // the warp before the next block's loads are issued).  Taps outside the source are loaded from a clamped
        // position and zeroed afterwards (BORDER_CONSTANT).
        uint8_t yv[GAIN_SPT][4], uv[GAIN_SPT][4], vv[GAIN_SPT][4];
        uint32_t bits[GAIN_SPT];
        bool slow[GAIN_SPT];
        #pragma unroll
        for (int h = 0; h < GAIN_SPT; h++) {
            const bool valid = sm[h].z != 0xFFFFFFFFu && (sm[h].y & C_VALID);
            const CamSrc& sc = p.src[valid ? (sm[h].z & 255u) : 0u];
            slow[h] = valid && (sc.vignette != nullptr || sc.rgb);   // vignette maps, packed RGB input: the per-tap path below
            bits[h] = !valid || slow[h] ? 0u : (sm[h].y & C_BORDER) ? (sm[h].y >> C_TAP_SHIFT) : 15u;
            const int ix = valid ? (int)(sm[h].x & 0xFFFFu) - 1 : sc.xlo, iy = valid ? (int)(sm[h].x >> 16) - 1 : 0;
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                const int x = min(max(ix + (k & 1), sc.xlo), sc.xhi), y = min(max(iy + (k >> 1), 0), sc.h - 1);
                yv[h][k] = __ldg(sc.y + (size_t)y * sc.y_pitch + x);
                uv[h][k] = __ldg(sc.u + (size_t)(y >> 1) * sc.u_pitch + (size_t)(x >> 1) * sc.uv_step);
                vv[h][k] = __ldg(sc.v + (size_t)(y >> 1) * sc.v_pitch + (size_t)(x >> 1) * sc.uv_step);
            }
        }
        #pragma unroll
        for (int h = 0; h < GAIN_SPT; h++) {
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                const int u = (int)uv[h][k] - 128, v = (int)vv[h][k] - 128;
                const uint32_t px = yuv_px((uint32_t)yv[h][k], (1 << 19) + 1673527 * v, (1 << 19) - 852492 * v - 409993 * u, (1 << 19) + 2116026 * u);
                t[h][k] = ((bits[h] >> k) & 1u) ? px : 0u;
            }
            if (slow[h]) { const uint4 q = gain_taps_slow(p.src[sm[h].z & 255u], sm[h].x, sm[h].y); t[h][0] = q.x; t[h][1] = q.y; t[h][2] = q.z; t[h][3] = q.w; }
        }
        #pragma unroll
        for (int h = 0; h < GAIN_SPT; h++)
            if (sm[h].z != 0xFFFFFFFFu) {
                int r, g, b;
                bilerp_rgbx(t[h][0], t[h][1], t[h][2], t[h][3], sm[h].y & 31u, (sm[h].y >> 5) & 31u, r, g, b);
                // x 2^52: exact (the norm is a multiple of 2^-52 below 2^9)
                s_nrm[(sm[h].z & 255u) * GAIN_PX + (sm[h].z >> 8)] = __double2ull_rz(sqrt((double)(r * r + g * g + b * b)) * 4503599627370496.0);
            }
    }
    __syncthreads();
    if (tr) tr[2] = gtime();
    #pragma unroll 1
    for (int q = warp; q < np; q += GAIN_BLK / 32) {
        const unsigned long long* na = s_nrm + s_pi[q] * GAIN_PX, *nb = s_nrm + s_pj[q] * GAIN_PX;
        unsigned long long cnt = 0, ha = 0, la = 0, hb = 0, lb = 0;
        #pragma unroll 4
        for (int l = lane; l < GAIN_PX; l += 32) {
            const unsigned long long qa = na[l], qb = nb[l];
            if (qa != GAIN_NONE && qb != GAIN_NONE) { cnt++; ha += qa >> 32; la += qa & 0xFFFFFFFFull; hb += qb >> 32; lb += qb & 0xFFFFFFFFull; }
        }
        #pragma unroll 1
        for (int o = 16; o > 0; o >>= 1) {                  // integer sums: any order gives the same result
            cnt += __shfl_down_sync(0xffffffffu, cnt, o); ha += __shfl_down_sync(0xffffffffu, ha, o); la += __shfl_down_sync(0xffffffffu, la, o);
            hb += __shfl_down_sync(0xffffffffu, hb, o); lb += __shfl_down_sync(0xffffffffu, lb, o);
        }
        if (lane == 0 && cnt) {
            unsigned long long* o = p.totals + 5 * q;
            atomicAdd(o, cnt); atomicAdd(o + 1, ha); atomicAdd(o + 2, la); atomicAdd(o + 3, hb); atomicAdd(o + 4, lb);
        }
    }
    if (tr) tr[3] = gtime();
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        if (tr) tr[4] = gtime();
        const unsigned int t = atomicInc(p.ticket, gain_blocks - 1);   // wraps to 0: self-resetting
        is_last = (t == gain_blocks - 1);
    }
    __syncthreads();
    if (!is_last) return;
    const unsigned long long T1 = gtime();
    __threadfence();
    for (int k = tid; k < n * n; k += GAIN_BLK) { Nm[k] = 0; Im[k] = 0; }
    // totals -> f64 (one rounding each), and reset for the next frame
    for (int q = tid; q < np; q += GAIN_BLK) {
        unsigned long long v[5];
        #pragma unroll
        for (int u = 0; u < 5; u++) { v[u] = __ldcg(p.totals + 5 * q + u); __stcg(p.totals + 5 * q + u, 0ull); }
        s_part[3 * q] = (double)v[0];
        s_part[3 * q + 1] = fma((double)v[1], 4294967296.0, (double)v[2]) * 2.220446049250313e-16;     // (hi 2^32 + lo) 2^-52
        s_part[3 * q + 2] = fma((double)v[3], 4294967296.0, (double)v[4]) * 2.220446049250313e-16;
    }
    __syncthreads();
    if (tid < np) {
        const int ii = s_pi[tid], jj = s_pj[tid];
        const GainCam ca = p.cam[ii], cb = p.cam[jj];
        const bool overlap = max(ca.sx, cb.sx) < min(ca.sx + ca.sw, cb.sx + cb.sw) &&
                             max(ca.sy, cb.sy) < min(ca.sy + ca.sh, cb.sy + cb.sh);
        if (overlap) {                                   // otherwise N = I = 0 (exposure_compensate.cpp:105)
            const double c = s_part[3 * tid], u = s_part[3 * tid + 1], w = s_part[3 * tid + 2];
            const double nn = c > 1 ? c : 1;             // N = max(1, countNonZero)
            Nm[ii * n + jj] = nn; Nm[jj * n + ii] = nn;
            Im[ii * n + jj] = u / nn;
            if (jj != ii) Im[jj * n + ii] = w / nn;
        }
    }
    __syncthreads();
    const unsigned long long T2 = gtime();
    __shared__ double s_g[MAX_CAMS];
    if (warp == 0) { if (n > 3) gain_solve_lu_warp(n, Nm, Im, Aug, s_g); else gain_solve_warp(n, Nm, Im, Aug, s_g); }
    __syncthreads();
    if (tid < n) p.gains[tid] = s_g[tid];                   // Mapper::gains(); the tables below take them from shared memory
    const unsigned long long T3 = gtime();
    gain_tables(p, s_g);                    // 256 threads = the 256 input values
    if (tid == 0 && p.dbg) { p.dbg[0] = T0; p.dbg[1] = T1; p.dbg[2] = T2; p.dbg[3] = T3; p.dbg[4] = gtime(); }
}

__global__ void __launch_bounds__(256) k_gain_finalize(const GainParams p) { gain_tables(p, nullptr); }

// Horizontally fused front end of a frame: `gain_blocks` CTAs compute the gain statistics and solve (they read the
// input planes directly, so they do not depend on the conversion), all other CTAs convert the inputs to RGBX.  One
// launch, both parts run concurrently, no cross-stream synchronisation.  The gain CTAs are latency-bound and light, the
// conversion CTAs DRAM-bound: at the head of the grid they are interleaved 1 : 2, so the whole gain chain starts within
// the first third of the conversion without ever holding more than a fraction of the resident CTA slots.
__global__ void __launch_bounds__(256, 5) k_convert_gain(const __grid_constant__ ConvertParams cp, const __grid_constant__ GainParams gp, const unsigned gain_blocks,
                                                         const unsigned interleave)
{
    extern __shared__ unsigned long long s_dyn_nrm[];
    const unsigned b = blockIdx.x;
    // the next kernel of the stream (K_blend_ring, when launched with programmatic stream serialisation) may start its prologue
    // once every CTA of this grid has got this far; it still waits for this grid to complete before reading anything we write
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    struct Stamp {                                          // diagnostics (OCTVR_GAIN_TRACE): first start / last end of the conversion CTAs
        unsigned long long* d; bool on;
        __device__ Stamp(unsigned long long* dbg) : d(dbg), on(dbg && threadIdx.x == 0) { if (on) atomicMin(d + 6, gtime()); }
        __device__ ~Stamp() { if (on) atomicMax(d + 7, gtime()); }
    };
    if (interleave) {
        if (b < 3u * gain_blocks) {
            if (b % 3u == 0u) gain_body(gp, s_dyn_nrm, gain_blocks, b / 3u);
            else { Stamp st(gp.dbg_trace); convert_body(cp, (int)(b - (b / 3u + 1u))); }
        } else { Stamp st(gp.dbg_trace); convert_body(cp, (int)(b - gain_blocks)); }
    } else if (b < gain_blocks) gain_body(gp, s_dyn_nrm, gain_blocks, b);
    else { Stamp st(gp.dbg_trace); convert_body(cp, (int)(b - gain_blocks)); }
}

void launch_convert_gain(const ConvertParams& cp, const GainParams* gp, cudaStream_t s)
{
    static const GainParams none = {};
    const unsigned gb = gp ? (unsigned)gp->grid : 0u, cb = (unsigned)(cp.grid_x * cp.grid_y * cp.n);
    const size_t smem = gp ? (size_t)gp->n * GAIN_PX * sizeof(unsigned long long) : 0;
    static const int mode = [] { const char* e = getenv("OCTVR_GAIN_INTERLEAVE"); return e ? atoi(e) : 1; }();   // diagnostic: 0 = gain CTAs first
    k_convert_gain<<<gb + cb, 256, smem, s>>>(cp, gp ? *gp : none, gb, (mode && gb > 0 && cb >= 2u * gb) ? 1u : 0u);
}
void launch_gain_finalize(const GainParams& p, cudaStream_t s) { k_gain_finalize<<<1, 256, 0, s>>>(p); }

// ------------------------------------------------------------------------------------------------
// K_blend : one CTA per 32x8 output tile, one thread per output pixel.  For each camera that covers
// the tile ("job") a thread reads its table entry (8 B coords + 4 B weight, coalesced, streamed past
// L2 with evict_first), gathers four RGBX taps, interpolates, applies the gain and accumulates
// trunc(v * W) in registers.  The pixel is normalised, converted to YUV 4:2:0 and stored once -- no
// intermediate image touches DRAM.  Per-job constants (source plane, pitch, gain) are staged in
// shared memory once per CTA.
// ------------------------------------------------------------------------------------------------
struct JobInfo { const uint32_t* src; int pitch; uint32_t gain; };   // gain: f32 bits, 0xFFFFFFFF -> use the LUT
static_assert(sizeof(JobInfo) == 16, "JobInfo is read with one 128-bit shared load");

__device__ __forceinline__ uint2 ld_stream_u2(const uint2* ptr, uint64_t pol)
{
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_stream_f(const float* ptr, uint64_t pol)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(ptr), "l"(pol));
    return v;
}

// accumulate one table entry into the pixel's sums
__device__ __forceinline__ void blend_pair(const JobInfo& ji, uint2 cc, float ww, const uint8_t* __restrict__ lut_base,
                                           uint32_t t00, uint32_t t01, uint32_t t10, uint32_t t11,
                                           uint32_t& ar, uint32_t& ag, uint32_t& ab)
{
    int r, g, b;
    bilerp_rgbx(t00, t01, t10, t11, cc.y & 31u, (cc.y >> 5) & 31u, r, g, b);
    float rf = (float)r, gf = (float)g, bf = (float)b;
    if (ji.gain != 0u) {
        if (ji.gain != 0xFFFFFFFFu) {
            const float g32 = __int_as_float((int)ji.gain);
            rf = gain_apply_f32(rf, g32); gf = gain_apply_f32(gf, g32); bf = gain_apply_f32(bf, g32);
        } else {
            const uint8_t* lut = lut_base + (ji.pitch >> 24) * 256;
            rf = (float)__ldg(lut + r); gf = (float)__ldg(lut + g); bf = (float)__ldg(lut + b);
        }
    }
    // (short)(v * W): f32 product, truncated (v*W >= 0 so floor == trunc); 2^23 bias removed in the same add.
    // Entries that do not contribute carry W = 0 and a harmless offset, so the loop needs no validity branch.
    ar += (uint32_t)__float_as_int(__fadd_rd(__fmul_rn(rf, ww), MAGIC_RD)) - 0x4B000000u;
    ag += (uint32_t)__float_as_int(__fadd_rd(__fmul_rn(gf, ww), MAGIC_RD)) - 0x4B000000u;
    ab += (uint32_t)__float_as_int(__fadd_rd(__fmul_rn(bf, ww), MAGIC_RD)) - 0x4B000000u;
}

__device__ __forceinline__ uint32_t luma(uint32_t px)
{
    return (269484u * (px & 255u) + 528482u * ((px >> 8) & 255u) + 102760u * ((px >> 16) & 255u) + (1u << 19) + (16u << 20)) >> 20;
}

// CTA = one 32 x 16 output tile, 256 threads, TWO pixels per thread (rows r and r+8) so every job
// iteration keeps two independent table->gather chains in flight; the next job's table entries are
// prefetched before the current job's arithmetic.
__global__ void __launch_bounds__(256) k_blend(const BlendParams p)
{
    __shared__ JobInfo s_job[MAX_CAMS];
    __shared__ uint32_t s_px[TILE_H][TILE_W + 1];
    const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;
    const int tile = (blockIdx.y + p.tile_y0) * p.tiles_x + blockIdx.x;
    const uint32_t j0 = __ldg(p.tile_job_start + tile);
    const int nj = (int)(__ldg(p.tile_job_start + tile + 1) - j0);
    if (tid < nj) {
        const int cam = __ldg(p.job_cam + j0 + tid);
        JobInfo ji;
        ji.src = p.rgbx[cam]; ji.pitch = p.src_pitch[cam];
        ji.gain = 0xFFFFFFFFu;
        if (!p.use_gain) ji.gain = 0u;
        else if (__ldg(p.gain_flag + cam) == 0) ji.gain = (uint32_t)__float_as_int(__ldg(p.gain_f32 + cam));
        if (ji.gain == 0xFFFFFFFFu) ji.pitch |= cam << 24;          // LUT path needs the camera id (pitch < 2^24)
        s_job[tid] = ji;
    }
    __syncthreads();

    const uint64_t pol = policy_evict_first();
    const uint2* cp = p.coords + (size_t)j0 * TILE_PX + tid;
    const float* wp = p.weights + (size_t)j0 * TILE_PX + tid;
    uint32_t ar0 = 0, ag0 = 0, ab0 = 0, ar1 = 0, ag1 = 0, ab1 = 0;      // sums of floor(v*W) for the two pixels
    uint2 c0 = make_uint2(0, 0), c1 = make_uint2(0, 0);
    float w0 = 0.f, w1 = 0.f;
    if (nj > 0) {
        c0 = ld_stream_u2(cp, pol); c1 = ld_stream_u2(cp + 256, pol);
        w0 = ld_stream_f(wp, pol); w1 = ld_stream_f(wp + 256, pol);
    }
    #pragma unroll 1
    for (int k = 0; k < nj; k++) {
        const JobInfo ji = s_job[k];
        const uint2 a0 = c0, a1 = c1;
        const float v0 = w0, v1 = w1;
        uint32_t t00, t01, t10, t11, u00, u01, u10, u11;
        fetch_taps(ji.src, ji.pitch & 0xFFFFFF, a0, t00, t01, t10, t11);
        fetch_taps(ji.src, ji.pitch & 0xFFFFFF, a1, u00, u01, u10, u11);
        if (k + 1 < nj) {
            cp += TILE_PX; wp += TILE_PX;
            c0 = ld_stream_u2(cp, pol); c1 = ld_stream_u2(cp + 256, pol);
            w0 = ld_stream_f(wp, pol); w1 = ld_stream_f(wp + 256, pol);
        }
        blend_pair(ji, a0, v0, p.gain_lut, t00, t01, t10, t11, ar0, ag0, ab0);
        blend_pair(ji, a1, v1, p.gain_lut, u00, u01, u10, u11, ar1, ag1, ab1);
    }
    const uint32_t px0 = normalise_px(ar0, ag0, ab0, p.inv_n), px1 = normalise_px(ar1, ag1, ab1, p.inv_n);
    s_px[ly][lx] = px0;
    s_px[ly + 8][lx] = px1;
    __syncthreads();

    // ---- store phase: threads re-mapped so each writes packed, coalesced words ----
    const int tx0 = blockIdx.x * TILE_W, ty0 = (blockIdx.y + p.tile_y0) * TILE_H;
    if (p.oy) {
        if (tid < 128) {                         // luma: 16 rows x 8 groups of 4 px -> one 32-bit store each
            const int row = tid >> 3, gx = (tid & 7) << 2;
            const int x = tx0 + gx, y = ty0 + row;
            if (y < p.out_h && x < p.out_w) {
                const uint32_t yv = luma(s_px[row][gx]) | (luma(s_px[row][gx + 1]) << 8) | (luma(s_px[row][gx + 2]) << 16) | (luma(s_px[row][gx + 3]) << 24);
                uint8_t* o = p.oy + (size_t)y * p.oy_pitch + x;
                if (x + 3 < p.out_w && (((uintptr_t)o) & 3) == 0) *reinterpret_cast<uint32_t*>(o) = yv;
                else for (int k = 0; k < 4 && x + k < p.out_w; k++) o[k] = (uint8_t)(yv >> (8 * k));
            }
        } else {                                 // chroma: 8 rows x 16 samples, top-left pixel of each 2x2
            const int t = tid - 128, row = t >> 4, cx = t & 15;
            const int x = tx0 + 2 * cx, y = ty0 + 2 * row;
            if (y < p.out_h && x < p.out_w) {
                const uint32_t px = s_px[2 * row][2 * cx];
                const int R = px & 255u, G = (px >> 8) & 255u, B = (px >> 16) & 255u;
                const size_t co = (size_t)(x >> 1) * p.uv_step;
                p.ou[(size_t)(y >> 1) * p.ou_pitch + co] = (uint8_t)((-155188 * R - 305135 * G + 460324 * B + (1 << 19) + (128 << 20)) >> 20);
                p.ov[(size_t)(y >> 1) * p.ov_pitch + co] = (uint8_t)((460324 * R - 385875 * G - 74448 * B + (1 << 19) + (128 << 20)) >> 20);
            }
        }
    }
    if (p.rgb_out) {
        #pragma unroll
        for (int h = 0; h < 2; h++) {
            const int x = tx0 + lx, y = ty0 + ly + 8 * h;
            if (x < p.out_w && y < p.out_h) {
                const uint32_t px = h ? px1 : px0;
                uint8_t* o = p.rgb_out + (size_t)y * p.rgb_pitch + 3 * x;
                o[0] = (uint8_t)px; o[1] = (uint8_t)(px >> 8); o[2] = (uint8_t)(px >> 16);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K_blend_staged
// ------------------------------------------------------------------------------------------------
// ---- TMA + mbarrier (sm_90+ PTX; SASS: UTMALDG / SYNCS) ----
__device__ __forceinline__ void mbar_init(uint64_t* mbar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the waiting warp sleeps in hardware instead of re-issuing the poll (the ncu source
// page showed the bare try_wait + branch loop at 14 % of all issued instructions of K_blend_staged)
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(phase), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int x, int y, uint64_t* mbar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"((uint64_t)tmap), "r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef OCTVR_STAGED_WAIT_ALL
#define OCTVR_STAGED_WAIT_ALL 0
#endif
template <int GAIN>   // 0: no gain, 1: verified f32 multiplier (LUT fallback per job if flagged)
__global__ void __launch_bounds__(256, 7) k_blend_staged(const __grid_constant__ StagedParams p)
{
    __shared__ __align__(128) uint32_t s_buf[STAGE_CAP];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ JobMeta s_job[MAX_CAMS];
    __shared__ uint32_t s_gain[MAX_CAMS];
    __shared__ uint32_t s_px[TILE_H][TILE_W + 1];
    const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;
    const int tile = (blockIdx.y + p.tile_y0) * p.tiles_x + blockIdx.x;
    const JobMeta* rec = p.jobs + (size_t)tile * MAX_CAMS;
    const int nj = __ldg(&rec->grp_nj) & 0xFF;
    const uint32_t j0 = (uint32_t)__ldg(&rec->j0);
    if (tid < nj) {
        const JobMeta jm = rec[tid];
        s_job[tid] = jm;
        uint32_t g = 0u;
        if (GAIN) g = __ldg(p.gain_flag + jm.cam) == 0 ? (uint32_t)__float_as_int(__ldg(p.gain_f32 + jm.cam)) : 0xFFFFFFFFu;
        s_gain[tid] = g;
    }
    if (tid == 0) { mbar_init(&s_mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();

    // sums of floor(v * W) <= 255 * MAX_CAMS < 2^16: R and G of a pixel share one accumulator, the two blue sums another
    uint32_t arg0 = 0, arg1 = 0, abb = 0;
    const uint2* ep = p.entries + (size_t)j0 * TILE_PX + tid;
    uint2 e0 = make_uint2(0, 0), e1 = make_uint2(0, 0);
    if (nj > 0) { e0 = __ldcs(ep); e1 = __ldcs(ep + 256); }
    int k = 0;
    uint32_t phase = 0;
    while (k < nj) {                                        // one pass per job group (almost always a single group)
        const int grp = s_job[k].grp_nj >> 8;
        int kend = k;
        while (kend < nj && (s_job[kend].grp_nj >> 8) == grp) kend++;
        if (tid == 0) {                                     // one thread issues the group's TMA box loads
            uint32_t bytes = 0;
            for (int q = k; q < kend; q++) bytes += (uint32_t)(s_job[q].bw * s_job[q].bh) * 4u;
            mbar_expect_tx(&s_mbar, bytes);
            for (int q = k; q < kend; q++)
                tma_load_2d(s_buf + s_job[q].soff, (const char*)p.tmaps + (size_t)s_job[q].tmap * 128, s_job[q].bx0, s_job[q].by0, &s_mbar);
        }
        // one warp polls the mbarrier, the others park at the CTA barrier (which costs no issue slots); the polling thread's
        // acquire plus the barrier make the TMA-written stage visible to every thread
        if (OCTVR_STAGED_WAIT_ALL || tid < 32) mbar_wait(&s_mbar, phase);
        if (!OCTVR_STAGED_WAIT_ALL) __syncthreads();
        phase ^= 1u;
        #pragma unroll 1
        for (; k < kend; k++) {
            const uint2 a0 = e0, a1 = e1;
            if (k + 1 < nj) { ep += TILE_PX; e0 = __ldcs(ep); e1 = __ldcs(ep + 256); }
            const uint8_t* s0 = reinterpret_cast<const uint8_t*>(s_buf);
            const uint8_t* s1 = s0 + s_job[k].bw * 4;
            const uint32_t g = s_gain[k];
            // the gather of device_common.cuh (fused_pair): floor / rint by magic-number adds on the FP32 pipe, IDP.2A bilinear
            if (!GAIN) {
                fused_pair<0, false, false>(a0.x, a0.y, s0, s1, 0.f, 0.f, nullptr, arg0, abb);
                fused_pair<0, false, true>(a1.x, a1.y, s0, s1, 0.f, 0.f, nullptr, arg1, abb);
            } else if (g != 0xFFFFFFFFu) {
                const float g32 = __int_as_float((int)g), gb = gain_bias_f32(g32);
                fused_pair<1, false, false>(a0.x, a0.y, s0, s1, g32, gb, nullptr, arg0, abb);
                fused_pair<1, false, true>(a1.x, a1.y, s0, s1, g32, gb, nullptr, arg1, abb);
            } else {
                const uint8_t* lut = p.gain_lut + s_job[k].cam * 256;
                fused_pair<1, true, false>(a0.x, a0.y, s0, s1, 0.f, 0.f, lut, arg0, abb);
                fused_pair<1, true, true>(a1.x, a1.y, s0, s1, 0.f, 0.f, lut, arg1, abb);
            }
        }
        if (k < nj) { __syncthreads(); fence_proxy_async(); }   // the stage is refilled (async proxy) by the next group
    }
    const uint32_t px0 = normalise_px(arg0 & 0xFFFFu, arg0 >> 16, abb & 0xFFFFu, p.inv_n), px1 = normalise_px(arg1 & 0xFFFFu, arg1 >> 16, abb >> 16, p.inv_n);
    s_px[ly][lx] = px0;
    s_px[ly + 8][lx] = px1;
    __syncthreads();

    const int tx0 = blockIdx.x * TILE_W, ty0 = (blockIdx.y + p.tile_y0) * TILE_H;
    if (p.oy) {
        if (tid < 128) {
            const int row = tid >> 3, gx = (tid & 7) << 2;
            const int x = tx0 + gx, y = ty0 + row;
            if (y < p.out_h && x < p.out_w) {
                const uint32_t yv = luma(s_px[row][gx]) | (luma(s_px[row][gx + 1]) << 8) | (luma(s_px[row][gx + 2]) << 16) | (luma(s_px[row][gx + 3]) << 24);
                uint8_t* o = p.oy + (size_t)y * p.oy_pitch + x;
                if (x + 3 < p.out_w && (((uintptr_t)o) & 3) == 0) *reinterpret_cast<uint32_t*>(o) = yv;
                else for (int q = 0; q < 4 && x + q < p.out_w; q++) o[q] = (uint8_t)(yv >> (8 * q));
            }
        } else {
            const int t = tid - 128, row = t >> 4, cx = t & 15;
            const int x = tx0 + 2 * cx, y = ty0 + 2 * row;
            if (y < p.out_h && x < p.out_w) {
                const uint32_t px = s_px[2 * row][2 * cx];
                const int R = px & 255u, G = (px >> 8) & 255u, B = (px >> 16) & 255u;
                const size_t co = (size_t)(x >> 1) * p.uv_step;
                p.ou[(size_t)(y >> 1) * p.ou_pitch + co] = (uint8_t)rgb_cb(R, G, B);
                p.ov[(size_t)(y >> 1) * p.ov_pitch + co] = (uint8_t)rgb_cr(R, G, B);
            }
        }
    }
    if (p.rgb_out) {
        #pragma unroll
        for (int h = 0; h < 2; h++) {
            const int x = tx0 + lx, y = ty0 + ly + 8 * h;
            if (x < p.out_w && y < p.out_h) {
                const uint32_t px = h ? px1 : px0;
                uint8_t* o = p.rgb_out + (size_t)y * p.rgb_pitch + 3 * x;
                o[0] = (uint8_t)px; o[1] = (uint8_t)(px >> 8); o[2] = (uint8_t)(px >> 16);
            }
        }
    }
}

void launch_blend_staged(const StagedParams& p, cudaStream_t s)
{
    if (p.use_gain) k_blend_staged<1><<<dim3(p.tiles_x, p.tiles_y_run), 256, 0, s>>>(p);
    else k_blend_staged<0><<<dim3(p.tiles_x, p.tiles_y_run), 256, 0, s>>>(p);
}

void launch_blend(const BlendParams& p, cudaStream_t s)
{
    k_blend<<<dim3(p.tiles_x, p.tiles_y_run), 256, 0, s>>>(p);
}

}  // namespace ob
