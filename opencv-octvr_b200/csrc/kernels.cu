// csrc/kernels.cu -- per-frame CUDA kernels of the feather / no-blend stitch path (sm_100a).
//
// Arithmetic contracts (bit-exact against the reference's CPU functions, see oracle/):
//   colour   : imgproc/src/color.cpp:6087-6169 (YUV->RGB), :6430-6481 (RGB->YUV 4:2:0)
//   bilinear : imgproc/src/imgwarp.cpp:4383-4442 + :3812-4020 -> (sum S_k a_k b_k + 512) >> 10 with
//              a,b in {32-f, f}; identical to ((sum S_k w_k) + 2^14) >> 15 with the 15-bit table
//   gain     : core/src/arithm.cpp multiply-by-scalar in f64 -> sat_u8(rint(v*g))
//   feather  : stitching/src/cuda/blender.cu:73-98 (short)(v*W) truncation, blenders.cpp:581 (x 1/N, rint)
#include "kernels.cuh"

namespace ob {

// ------------------------------------------------------------------------------------------------
// shared device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }

// fixed-point bilinear of three 8-bit channels from four RGBX taps.  fx, fy in [0, 32).
__device__ __forceinline__ void bilerp_rgbx(uint32_t t00, uint32_t t01, uint32_t t10, uint32_t t11,
                                            uint32_t fx, uint32_t fy, int& r, int& g, int& b)
{
    const uint32_t wx = fx * 65535u + 32u;                 // (32-fx) | fx << 16
    const uint32_t ay = 32u - fy, by = fy;
    const uint32_t rg0 = __byte_perm(t00, t01, 0x5140);    // R00 R01 G00 G01
    const uint32_t bb0 = __byte_perm(t00, t01, 0x6262);    // B00 B01 .. ..
    const uint32_t rg1 = __byte_perm(t10, t11, 0x5140);
    const uint32_t bb1 = __byte_perm(t10, t11, 0x6262);
    const uint32_t hr0 = __dp2a_lo(wx, rg0, 0u), hg0 = __dp2a_hi(wx, rg0, 0u), hb0 = __dp2a_lo(wx, bb0, 0u);
    const uint32_t hr1 = __dp2a_lo(wx, rg1, 0u), hg1 = __dp2a_hi(wx, rg1, 0u), hb1 = __dp2a_lo(wx, bb1, 0u);
    r = (int)((hr0 * ay + hr1 * by + 512u) >> 10);
    g = (int)((hg0 * ay + hg1 * by + 512u) >> 10);
    b = (int)((hb0 * ay + hb1 * by + 512u) >> 10);
}

// gather the four taps for a table entry; taps outside the source contribute 0 (BORDER_CONSTANT)
__device__ __forceinline__ void fetch_taps(const uint32_t* __restrict__ src, int pitch, uint2 c,
                                           uint32_t& t00, uint32_t& t01, uint32_t& t10, uint32_t& t11)
{
    const int off = (int)c.x;
    if (!(c.y & C_BORDER)) {
        t00 = __ldg(src + off); t01 = __ldg(src + off + 1);
        t10 = __ldg(src + off + pitch); t11 = __ldg(src + off + pitch + 1);
    } else {
        const uint32_t in = c.y >> C_TAP_SHIFT;
        t00 = (in & 1u) ? __ldg(src + off) : 0u;
        t01 = (in & 2u) ? __ldg(src + off + 1) : 0u;
        t10 = (in & 4u) ? __ldg(src + off + pitch) : 0u;
        t11 = (in & 8u) ? __ldg(src + off + pitch + 1) : 0u;
    }
}

constexpr float MAGIC_RN = 12582912.f;   // 1.5 * 2^23 : x + MAGIC rounds x to nearest-even integer
constexpr float MAGIC_RD = 8388608.f;    // 2^23 with round-down add : floor(x)

// sat_u8(rint(v * g)) for integer v in [0,255] as a float; g32 has been verified against the f64 rule
__device__ __forceinline__ float gain_apply_f32(float v, float g32)
{
    return fminf(__fadd_rn(__fmaf_rn(v, g32, MAGIC_RN), -MAGIC_RN), 255.f);
}

// ------------------------------------------------------------------------------------------------
// K_convert
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t yuv_to_rgbx(int Y, int ruv, int guv, int buv)
{
    const int yy = max(Y - 16, 0) * 1220542;
    return (uint32_t)clamp255((yy + ruv) >> 20) | ((uint32_t)clamp255((yy + guv) >> 20) << 8) |
           ((uint32_t)clamp255((yy + buv) >> 20) << 16);
}
__device__ __forceinline__ uint32_t vignette_rgbx(uint32_t p, float k)
{
    // cudaarithm mul_mat.cu:198-213 : saturate_cast<uchar>(u8 * f32), round-to-nearest-even
    const int r = clamp255(__float2int_rn((float)(p & 255u) * k));
    const int g = clamp255(__float2int_rn((float)((p >> 8) & 255u) * k));
    const int b = clamp255(__float2int_rn((float)((p >> 16) & 255u) * k));
    return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
}

// one thread: 4 px x 2 rows (two chroma samples).  CTA = 64 x 4 threads = 256 x 8 px.
__global__ void __launch_bounds__(256) k_convert(const ConvertParams p)
{
    int ci = 0;
    #pragma unroll 1
    while (ci + 1 < p.n && (int)blockIdx.x >= p.block_start[ci + 1]) ci++;
    const CamSrc& c = p.cam[ci];
    const int lb = blockIdx.x - p.block_start[ci];
    const int bx_n = (c.w + 255) >> 8;
    const int bx = lb % bx_n, by = lb / bx_n;
    const int x0 = (bx << 8) + (threadIdx.x << 2);
    const int y0 = (by << 3) + (threadIdx.y << 1);
    if (x0 >= c.w || y0 >= c.h) return;

    const uint8_t* yr0 = c.y + (size_t)y0 * c.y_pitch + x0;
    const uint8_t* yr1 = yr0 + c.y_pitch;
    const uint8_t* ur = c.u + (size_t)(y0 >> 1) * c.u_pitch + (size_t)(x0 >> 1) * c.uv_step;
    const uint8_t* vr = c.v + (size_t)(y0 >> 1) * c.v_pitch + (size_t)(x0 >> 1) * c.uv_step;
    uint32_t* o0 = c.rgbx + (size_t)y0 * c.w + x0;
    uint32_t* o1 = o0 + c.w;

    if (c.aligned4) {
        const uint32_t ya = __ldg(reinterpret_cast<const uint32_t*>(yr0));
        const uint32_t yb = __ldg(reinterpret_cast<const uint32_t*>(yr1));
        uint32_t a[4], b[4];
        #pragma unroll
        for (int k = 0; k < 2; k++) {
            const int u = (int)__ldg(ur + k * c.uv_step) - 128, v = (int)__ldg(vr + k * c.uv_step) - 128;
            const int ruv = (1 << 19) + 1673527 * v;
            const int guv = (1 << 19) - 852492 * v - 409993 * u;
            const int buv = (1 << 19) + 2116026 * u;
            a[2 * k] = yuv_to_rgbx((ya >> (16 * k)) & 255, ruv, guv, buv);
            a[2 * k + 1] = yuv_to_rgbx((ya >> (16 * k + 8)) & 255, ruv, guv, buv);
            b[2 * k] = yuv_to_rgbx((yb >> (16 * k)) & 255, ruv, guv, buv);
            b[2 * k + 1] = yuv_to_rgbx((yb >> (16 * k + 8)) & 255, ruv, guv, buv);
        }
        if (c.vignette) {
            const float4 k0 = __ldg(reinterpret_cast<const float4*>(c.vignette + (size_t)y0 * c.w + x0));
            const float4 k1 = __ldg(reinterpret_cast<const float4*>(c.vignette + (size_t)(y0 + 1) * c.w + x0));
            a[0] = vignette_rgbx(a[0], k0.x); a[1] = vignette_rgbx(a[1], k0.y); a[2] = vignette_rgbx(a[2], k0.z); a[3] = vignette_rgbx(a[3], k0.w);
            b[0] = vignette_rgbx(b[0], k1.x); b[1] = vignette_rgbx(b[1], k1.y); b[2] = vignette_rgbx(b[2], k1.z); b[3] = vignette_rgbx(b[3], k1.w);
        }
        *reinterpret_cast<uint4*>(o0) = make_uint4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<uint4*>(o1) = make_uint4(b[0], b[1], b[2], b[3]);
    } else {
        for (int k = 0; k < 4 && x0 + k < c.w; k++) {
            const int u = (int)__ldg(ur + (k >> 1) * c.uv_step) - 128, v = (int)__ldg(vr + (k >> 1) * c.uv_step) - 128;
            const int ruv = (1 << 19) + 1673527 * v;
            const int guv = (1 << 19) - 852492 * v - 409993 * u;
            const int buv = (1 << 19) + 2116026 * u;
            uint32_t a = yuv_to_rgbx(__ldg(yr0 + k), ruv, guv, buv), b = yuv_to_rgbx(__ldg(yr1 + k), ruv, guv, buv);
            if (c.vignette) {
                a = vignette_rgbx(a, __ldg(c.vignette + (size_t)y0 * c.w + x0 + k));
                b = vignette_rgbx(b, __ldg(c.vignette + (size_t)(y0 + 1) * c.w + x0 + k));
            }
            o0[k] = a; o1[k] = b;
        }
    }
}

void launch_convert(const ConvertParams& p, cudaStream_t s)
{
    k_convert<<<p.block_start[p.n], dim3(64, 4), 0, s>>>(p);
}

// ------------------------------------------------------------------------------------------------
// K_gain
// ------------------------------------------------------------------------------------------------
// one thread per working-scale pixel of every camera: remap that one pixel, store r^2+g^2+b^2
__global__ void __launch_bounds__(256) k_gain_norms(const GainParams p)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.total) return;
    int ci = 0;
    #pragma unroll 1
    while (ci + 1 < p.n && t >= p.cam[ci + 1].off) ci++;
    int out = -1;
    if (__ldg(p.smask + t) == 255) {               // CPU compensator's intersect rule (exposure_compensate.cpp:71-78,112)
        out = 0;
        const uint2 c = __ldg(p.gcoord + t);
        if (c.y & C_VALID) {
            uint32_t t00, t01, t10, t11;
            fetch_taps(p.rgbx[ci], p.src_pitch[ci], c, t00, t01, t10, t11);
            int r, g, b;
            bilerp_rgbx(t00, t01, t10, t11, c.y & 31u, (c.y >> 5) & 31u, r, g, b);
            out = r * r + g * g + b * b;
        }
    }
    p.sq[t] = out;
}

__device__ __forceinline__ double warp_sum(double v)
{
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// exposure_compensate.cpp:138-153 + core LU / closed forms (matrix_decomp.cpp:52-109, lapack.cpp:1080-1170)
__device__ void gain_solve(int n, const double* Nm, const double* Im, double* A, double* b, double* g)
{
    const double alpha = 0.01, beta = 100;
    for (int i = 0; i < n; i++) { b[i] = 0; for (int j = 0; j < n; j++) A[i * n + j] = 0; }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            b[i] += beta * Nm[i * n + j];
            A[i * n + i] += beta * Nm[i * n + j];
            if (j == i) continue;
            A[i * n + i] += 2 * alpha * Im[i * n + j] * Im[i * n + j] * Nm[i * n + j];
            A[i * n + j] -= 2 * alpha * Im[i * n + j] * Im[j * n + i] * Nm[i * n + j];
        }
    if (n == 2) {
        double d = 1. / (A[0] * A[3] - A[1] * A[2]);
        g[0] = (b[0] * A[3] - b[1] * A[1]) * d;
        g[1] = (b[1] * A[0] - b[0] * A[2]) * d;
        return;
    }
    if (n == 3) {
        #define S(i, j) A[(i) * 3 + (j)]
        double d = S(0,0) * (S(1,1) * S(2,2) - S(1,2) * S(2,1)) - S(0,1) * (S(1,0) * S(2,2) - S(1,2) * S(2,0)) +
                   S(0,2) * (S(1,0) * S(2,1) - S(1,1) * S(2,0));
        d = 1. / d;
        g[0] = ((S(1,1) * S(2,2) - S(1,2) * S(2,1)) * b[0] + (S(0,2) * S(2,1) - S(0,1) * S(2,2)) * b[1] + (S(0,1) * S(1,2) - S(0,2) * S(1,1)) * b[2]) * d;
        g[1] = ((S(1,2) * S(2,0) - S(1,0) * S(2,2)) * b[0] + (S(0,0) * S(2,2) - S(0,2) * S(2,0)) * b[1] + (S(0,2) * S(1,0) - S(0,0) * S(1,2)) * b[2]) * d;
        g[2] = ((S(1,0) * S(2,1) - S(1,1) * S(2,0)) * b[0] + (S(0,1) * S(2,0) - S(0,0) * S(2,1)) * b[1] + (S(0,0) * S(1,1) - S(0,1) * S(1,0)) * b[2]) * d;
        #undef S
        return;
    }
    for (int i = 0; i < n; i++) {                    // LU with partial pivoting
        int k = i;
        for (int j = i + 1; j < n; j++) if (fabs(A[j * n + i]) > fabs(A[k * n + i])) k = j;
        if (k != i) {
            for (int j = i; j < n; j++) { double t = A[i * n + j]; A[i * n + j] = A[k * n + j]; A[k * n + j] = t; }
            double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        double d = -1 / A[i * n + i];
        for (int j = i + 1; j < n; j++) {
            double al = A[j * n + i] * d;
            for (int kk = i + 1; kk < n; kk++) A[j * n + kk] += al * A[i * n + kk];
            b[j] += al * b[i];
        }
        A[i * n + i] = -d;
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < n; k++) s -= A[i * n + k] * b[k];
        b[i] = s * A[i * n + i];
    }
    for (int i = 0; i < n; i++) g[i] = b[i];
}

// Per camera: the exact u8 gain LUT in f64, and an f32 multiplier for which the single-FMA formula of
// gain_apply_f32() reproduces that LUT for all 256 inputs (searched within +-2 ulp of (float)g).
// If none exists the camera is flagged and the blend kernel reads the LUT instead.  256 threads.
__device__ void gain_tables(const GainParams& p)
{
    const int v = threadIdx.x;
    for (int c = 0; c < p.n; c++) {
        const double g = p.gains[c];
        const int exact = clamp255(__double2int_rn((double)v * g));
        p.gain_lut[c * 256 + v] = (uint8_t)exact;
        int chosen = -1;
        const float g0 = (float)g;
        if (g > 0. && g < 4096.) {
            #pragma unroll 1
            for (int k = 0; k < 5 && chosen < 0; k++) {
                const int step = (k == 0) ? 0 : (k & 1) ? (k + 1) / 2 : -(k / 2);
                const float gc = __int_as_float(__float_as_int(g0) + step);
                const int got = (int)gain_apply_f32((float)v, gc);
                if (__syncthreads_and(got == exact)) chosen = k;
            }
        }
        if (v == 0) {
            const int step = (chosen <= 0) ? 0 : (chosen & 1) ? (chosen + 1) / 2 : -(chosen / 2);
            p.gain_f32[c] = __int_as_float(__float_as_int(g0) + step);
            p.gain_flag[c] = chosen < 0 ? 1 : 0;
        }
    }
}

// CTA (pair, chunk): masked sums over the pair's overlap rectangle; the last CTA to finish reduces
// the partials in a fixed order, solves for the gains and builds the gain tables.  256 threads.
__global__ void __launch_bounds__(256) k_gain_reduce_solve(const GainParams p)
{
    __shared__ double red[3][8];
    __shared__ double Nm[MAX_CAMS * MAX_CAMS], Im[MAX_CAMS * MAX_CAMS], A[MAX_CAMS * MAX_CAMS], bb[MAX_CAMS];
    __shared__ bool is_last;
    const int pair = blockIdx.x / p.chunks, chunk = blockIdx.x % p.chunks;
    int i = 0, rem = pair;
    while (rem >= p.n - i) { rem -= p.n - i; i++; }
    const int j = i + rem;
    const GainCam a = p.cam[i], b = p.cam[j];
    const int x_tl = max(a.sx, b.sx), y_tl = max(a.sy, b.sy);
    const int x_br = min(a.sx + a.sw, b.sx + b.sw), y_br = min(a.sy + a.sh, b.sy + b.sh);
    double cnt = 0, s1 = 0, s2 = 0;
    if (x_tl < x_br && y_tl < y_br) {
        const int rw = x_br - x_tl, area = rw * (y_br - y_tl);
        for (int t = chunk * 256 + threadIdx.x; t < area; t += p.chunks * 256) {
            const int x = x_tl + t % rw, y = y_tl + t / rw;
            const int qa = p.sq[a.off + (y - a.sy) * a.sw + (x - a.sx)];
            const int qb = p.sq[b.off + (y - b.sy) * b.sw + (x - b.sx)];
            if (qa >= 0 && qb >= 0) { cnt += 1; s1 += sqrt((double)qa); s2 += sqrt((double)qb); }
        }
    }
    cnt = warp_sum(cnt); s1 = warp_sum(s1); s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = cnt; red[1][threadIdx.x >> 5] = s1; red[2][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0, u = 0, w = 0;
        for (int k = 0; k < 8; k++) { c += red[0][k]; u += red[1][k]; w += red[2][k]; }
        double* o = p.partial + (size_t)blockIdx.x * 3;
        o[0] = c; o[1] = u; o[2] = w;
        __threadfence();
        const unsigned int t = atomicInc(p.ticket, gridDim.x - 1);   // wraps to 0: self-resetting
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        const int n = p.n;
        for (int k = 0; k < n * n; k++) { Nm[k] = 0; Im[k] = 0; }
        int pr = 0;
        for (int ii = 0; ii < n; ii++)
            for (int jj = ii; jj < n; jj++, pr++) {
                const GainCam ca = p.cam[ii], cb = p.cam[jj];
                const bool overlap = max(ca.sx, cb.sx) < min(ca.sx + ca.sw, cb.sx + cb.sw) &&
                                     max(ca.sy, cb.sy) < min(ca.sy + ca.sh, cb.sy + cb.sh);
                if (!overlap) continue;
                double c = 0, u = 0, w = 0;
                for (int k = 0; k < p.chunks; k++) {
                    const volatile double* o = p.partial + ((size_t)pr * p.chunks + k) * 3;
                    c += o[0]; u += o[1]; w += o[2];
                }
                const double nn = c > 1 ? c : 1;      // N = max(1, countNonZero)
                Nm[ii * n + jj] = Nm[jj * n + ii] = nn;
                Im[ii * n + jj] = u / nn;
                Im[jj * n + ii] = w / nn;
            }
        gain_solve(n, Nm, Im, A, bb, p.gains);
        __threadfence();
    }
    __syncthreads();
    gain_tables(p);
}

__global__ void __launch_bounds__(256) k_gain_finalize(const GainParams p) { gain_tables(p); }

void launch_gain_norms(const GainParams& p, cudaStream_t s)
{
    k_gain_norms<<<(p.total + 255) / 256, 256, 0, s>>>(p);
}
void launch_gain_reduce_solve(const GainParams& p, cudaStream_t s)
{
    k_gain_reduce_solve<<<p.n_pairs * p.chunks, 256, 0, s>>>(p);
}
void launch_gain_finalize(const GainParams& p, cudaStream_t s) { k_gain_finalize<<<1, 256, 0, s>>>(p); }

// ------------------------------------------------------------------------------------------------
// K_blend : one CTA per 32x8 output tile, one thread per output pixel.  For each camera that covers
// the tile ("job") a thread reads its table entry (8 B coords + 4 B weight, coalesced), gathers four
// RGBX taps, interpolates, applies the gain and accumulates trunc(v * W) in registers.  The pixel is
// normalised, converted to YUV 4:2:0 and stored once -- no intermediate image touches DRAM.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_PX) k_blend(const BlendParams p)
{
    const int tile = blockIdx.x;
    const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    const int tid = threadIdx.x;
    const int x = tx * TILE_W + (tid & (TILE_W - 1)), y = ty * TILE_H + (tid >> 5);
    const uint32_t j0 = __ldg(p.tile_job_start + tile), j1 = __ldg(p.tile_job_start + tile + 1);

    // accumulators hold sum of (2^23-biased) floor(v*W) bit patterns; the bias is removed at the end
    uint32_t ar = 0, ag = 0, ab = 0, nacc = 0;
    uint2 c = make_uint2(0, 0);
    float w = 0.f;
    if (j0 < j1) { c = __ldg(p.coords + (size_t)j0 * TILE_PX + tid); w = __ldg(p.weights + (size_t)j0 * TILE_PX + tid); }
    for (uint32_t j = j0; j < j1; j++) {
        const uint2 cc = c;
        const float ww = w;
        if (j + 1 < j1) {                                  // prefetch the next job's entry
            c = __ldg(p.coords + (size_t)(j + 1) * TILE_PX + tid);
            w = __ldg(p.weights + (size_t)(j + 1) * TILE_PX + tid);
        }
        if (!(cc.y & C_VALID)) continue;
        const int cam = __ldg(p.job_cam + j);
        uint32_t t00, t01, t10, t11;
        fetch_taps(p.rgbx[cam], p.src_pitch[cam], cc, t00, t01, t10, t11);
        int r, g, b;
        bilerp_rgbx(t00, t01, t10, t11, cc.y & 31u, (cc.y >> 5) & 31u, r, g, b);
        float rf, gf, bf;
        if (p.use_gain) {
            if (__ldg(p.gain_flag + cam) == 0) {
                const float g32 = __ldg(p.gain_f32 + cam);
                rf = gain_apply_f32((float)r, g32); gf = gain_apply_f32((float)g, g32); bf = gain_apply_f32((float)b, g32);
            } else {
                const uint8_t* lut = p.gain_lut + cam * 256;
                rf = (float)__ldg(lut + r); gf = (float)__ldg(lut + g); bf = (float)__ldg(lut + b);
            }
        } else { rf = (float)r; gf = (float)g; bf = (float)b; }
        // (short)(v * W): f32 product, truncated (v*W >= 0 so floor == trunc)
        ar += (uint32_t)__float_as_int(__fadd_rd(__fmul_rn(rf, ww), MAGIC_RD));
        ag += (uint32_t)__float_as_int(__fadd_rd(__fmul_rn(gf, ww), MAGIC_RD));
        ab += (uint32_t)__float_as_int(__fadd_rd(__fmul_rn(bf, ww), MAGIC_RD));
        nacc++;
    }
    if (x >= p.out_w || y >= p.out_h) return;
    const uint32_t bias = nacc * 0x4B000000u;
    // dst_16s.convertTo(CV_8UC3, 1.0/N): sat_u8(rint((float)acc * (float)(1/N)))
    const int R = min(__float2int_rn(__fmul_rn((float)(int)(ar - bias), p.inv_n)), 255);
    const int G = min(__float2int_rn(__fmul_rn((float)(int)(ag - bias), p.inv_n)), 255);
    const int B = min(__float2int_rn(__fmul_rn((float)(int)(ab - bias), p.inv_n)), 255);
    if (p.rgb_out) {
        uint8_t* o = p.rgb_out + (size_t)y * p.rgb_pitch + 3 * x;
        o[0] = (uint8_t)R; o[1] = (uint8_t)G; o[2] = (uint8_t)B;
    }
    if (p.oy) {
        p.oy[(size_t)y * p.oy_pitch + x] = (uint8_t)((269484 * R + 528482 * G + 102760 * B + (1 << 19) + (16 << 20)) >> 20);
        if (((x | y) & 1) == 0) {
            const size_t co = (size_t)(x >> 1) * p.uv_step;
            p.ou[(size_t)(y >> 1) * p.ou_pitch + co] = (uint8_t)((-155188 * R - 305135 * G + 460324 * B + (1 << 19) + (128 << 20)) >> 20);
            p.ov[(size_t)(y >> 1) * p.ov_pitch + co] = (uint8_t)((460324 * R - 385875 * G - 74448 * B + (1 << 19) + (128 << 20)) >> 20);
        }
    }
}

void launch_blend(const BlendParams& p, cudaStream_t s)
{
    k_blend<<<p.tiles_x * p.tiles_y, TILE_PX, 0, s>>>(p);
}

}  // namespace ob
