// Pre-inference clean-up pipes: the step right before the UNet.
//   ImageNorm        reference pipeline.py:338-356   (x - mean) / std per image and channel
//   ImageOutliers    reference pipeline.py:266-296   hot-pixel removal against a size x size median
//   ImageBGSubtract  reference pipeline.py:360-405   least-squares quadratic background surface
// All three are HBM-bound: one or two streaming passes over a float32 (N,H,W,C) stack.
#include "sq_common.cuh"
#include <cuda_bf16.h>
#include <cmath>

namespace {

constexpr int PREP_BLOCKS = 512;          // partial-sum blocks per image (>= 3 per SM even for one frame)
constexpr int PREP_THREADS = 256;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sums of NV values per thread; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *sm)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = warp_sum(v[i]);
        if (lane == 0) sm[warp * NV + i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
            for (int w = 0; w < PREP_THREADS / 32; ++w) s += sm[w * NV + i];
            v[i] = s;
        }
    }
}

// four consecutive stored values (float32, or raw uint16 / uint8 camera counts) as floats, one load
template <typename T> __device__ __forceinline__ float4 load4(const T *p);
template <> __device__ __forceinline__ float4 load4<float>(const float *p)
{
    return __ldg(reinterpret_cast<const float4 *>(p));
}
template <> __device__ __forceinline__ float4 load4<uint16_t>(const uint16_t *p)
{
    const uint2 u = __ldg(reinterpret_cast<const uint2 *>(p));
    return make_float4((float)(u.x & 0xffffu), (float)(u.x >> 16), (float)(u.y & 0xffffu), (float)(u.y >> 16));
}
template <> __device__ __forceinline__ float4 load4<uint8_t>(const uint8_t *p)
{
    const unsigned u = __ldg(reinterpret_cast<const unsigned *>(p));
    return make_float4((float)(u & 0xffu), (float)((u >> 8) & 0xffu), (float)((u >> 16) & 0xffu), (float)(u >> 24));
}

// ------------------------------------------------------------------ ImageNorm
// pass 1: per (image, channel, block) partial sum and sum of squares in fp64
template <typename T>
__global__ void __launch_bounds__(PREP_THREADS)
norm_partial(const T *__restrict__ in, long long npix, int C, double *__restrict__ part)
{
    __shared__ double sm[PREP_THREADS / 32 * 2];
    const int n = blockIdx.y, c = blockIdx.z;
    const T *img = in + (size_t)n * npix * C + c;
    double v[2] = {0.0, 0.0};
    if (C == 1 && (npix & 3) == 0) {
        // single channel: 16-byte loads, four independent accumulators per moment
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0, q0 = 0, q1 = 0, q2 = 0, q3 = 0;
        // four 16-byte loads in flight per thread (a frame is only ~8 iterations per thread: without
        // the batching the pass is latency-bound), accumulated in the same order as a plain loop
        constexpr long long STRIDE = (long long)PREP_BLOCKS * PREP_THREADS;
        const long long nv = npix >> 2;
        long long p = (long long)blockIdx.x * PREP_THREADS + threadIdx.x;
        auto acc4 = [&](const float4 f) {
            const double a = f.x, b = f.y, cc = f.z, d = f.w;
            s0 += a; s1 += b; s2 += cc; s3 += d;
            q0 = fma(a, a, q0); q1 = fma(b, b, q1); q2 = fma(cc, cc, q2); q3 = fma(d, d, q3);
        };
        for (; p + 3 * STRIDE < nv; p += 4 * STRIDE) {
            const float4 f0 = load4<T>(img + 4 * p), f1 = load4<T>(img + 4 * (p + STRIDE));
            const float4 f2 = load4<T>(img + 4 * (p + 2 * STRIDE)), f3 = load4<T>(img + 4 * (p + 3 * STRIDE));
            acc4(f0); acc4(f1); acc4(f2); acc4(f3);
        }
        for (; p < nv; p += STRIDE) acc4(load4<T>(img + 4 * p));
        v[0] = (s0 + s1) + (s2 + s3);
        v[1] = (q0 + q1) + (q2 + q3);
    } else {
        for (long long p = (long long)blockIdx.x * PREP_THREADS + threadIdx.x; p < npix;
             p += (long long)PREP_BLOCKS * PREP_THREADS) {
            const double x = (double)img[p * C];
            v[0] += x;
            v[1] = fma(x, x, v[1]);
        }
    }
    block_sum<2>(v, sm);
    if (threadIdx.x == 0) {
        double *o = part + (((size_t)n * C + c) * PREP_BLOCKS + blockIdx.x) * 2;
        o[0] = v[0];
        o[1] = v[1];
    }
}

// pass 2: mean / std -> float32 (the dtype numpy's float32 reductions return, pipeline.py:353-355)
__global__ void norm_final(const double *__restrict__ part, long long npix, int nc, float2 *__restrict__ stats)
{
    // one warp per (image, channel); fixed summation order -> deterministic
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= nc) return;
    double s = 0.0, q = 0.0;
    double2 v[PREP_BLOCKS / 32];
#pragma unroll
    for (int k = 0; k < PREP_BLOCKS / 32; ++k)          // every load in flight before the first add
        v[k] = reinterpret_cast<const double2 *>(part)[(size_t)i * PREP_BLOCKS + lane + 32 * k];
#pragma unroll
    for (int k = 0; k < PREP_BLOCKS / 32; ++k) { s += v[k].x; q += v[k].y; }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane != 0) return;
    const double mean = s / (double)npix;
    double var = q / (double)npix - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[i] = make_float2((float)mean, (float)sqrt(var));
}

// pass 3: float32 arithmetic exactly as numpy does it on a float32 image (epsilon 1e-99 vanishes in
// float32, pipeline.py:350,354-355: a constant image divides by zero there too)
template <typename T>
__global__ void norm_apply(const T *__restrict__ in, float *__restrict__ out, long long per_image, int C,
                           const float2 *__restrict__ stats)
{
    // grid (chunks, n): per_image = pixels * C values of image blockIdx.y
    const int n = blockIdx.y;
    const T *src = in + (size_t)n * per_image;
    float *dst = out + (size_t)n * per_image;
    if (C == 1 && (per_image & 3) == 0) {
        const float2 st = stats[n];
        const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= (per_image >> 2)) return;
        float4 f = load4<T>(src + 4 * i);
        f.x = __fdiv_rn(__fsub_rn(f.x, st.x), st.y);
        f.y = __fdiv_rn(__fsub_rn(f.y, st.x), st.y);
        f.z = __fdiv_rn(__fsub_rn(f.z, st.x), st.y);
        f.w = __fdiv_rn(__fsub_rn(f.w, st.x), st.y);
        reinterpret_cast<float4 *>(dst)[i] = f;
    } else {
        for (int k = 0; k < 4; ++k) {
            const long long i = ((long long)blockIdx.x * 4 + k) * blockDim.x + threadIdx.x;
            if (i >= per_image) return;
            const float2 st = stats[n * C + (int)(i % C)];
            dst[i] = __fdiv_rn(__fsub_rn((float)src[i], st.x), st.y);
        }
    }
}

// -------------------------------------------------------------- ImageOutliers
// scipy.ndimage.median_filter(x, size): window = offsets -(size/2) .. -(size/2)+size-1 per axis, 'reflect'
// boundary (d c b a | a b c d | d c b a), value of rank (size*size)/2; pipeline.py:287-294.
__device__ __forceinline__ int reflect_idx(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? (-i - 1) : (2 * n - 1 - i);
    return i;
}

template <int K>
__global__ void __launch_bounds__(256)
outliers_kernel(const float *__restrict__ in, float *__restrict__ out, int H, int W, int C, float threshold)
{
    // grid (ceil(W*C/256), H, n): one thread per (pixel, channel) of a row
    const int xc = blockIdx.x * 256 + threadIdx.x;
    if (xc >= W * C) return;
    const int y = blockIdx.y, n = blockIdx.z;
    const int x = xc / C, c = xc - x * C;
    const float *img = in + (size_t)n * H * W * C + c;
    float v[K * K];
#pragma unroll
    for (int dy = 0; dy < K; ++dy) {
        const int yy = reflect_idx(y - K / 2 + dy, H);
#pragma unroll
        for (int dx = 0; dx < K; ++dx) {
            const int xx = reflect_idx(x - K / 2 + dx, W);
            v[dy * K + dx] = __ldg(img + ((size_t)yy * W + xx) * C);
        }
    }
    // partial selection sort up to the median rank
    constexpr int RANK = (K * K) / 2;
#pragma unroll
    for (int a = 0; a <= RANK; ++a) {
#pragma unroll
        for (int b = a + 1; b < K * K; ++b) {
            const float lo = fminf(v[a], v[b]), hi = fmaxf(v[a], v[b]);
            v[a] = lo;
            v[b] = hi;
        }
    }
    const size_t i = ((size_t)n * H + y) * W * C + xc;
    const float med = v[RANK], raw = in[i];
    out[i] = (fabsf(__fsub_rn(raw, med)) > threshold) ? med : raw;
}

// Single-channel 2x2 fast path (the reference's default sigma = 2): one thread = 4 consecutive pixels,
// 16-byte loads of rows y-1 and y plus the left neighbours; rank 2 of 4 = second largest.
__global__ void __launch_bounds__(256)
outliers2_c1_kernel(const float *__restrict__ in, float *__restrict__ out, int H, int W, float threshold)
{
    const int x4 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x4 >= W) return;
    const int y = blockIdx.y, n = blockIdx.z;
    const float *r1 = in + ((size_t)n * H + y) * W;                      // row y
    const float *r0 = in + ((size_t)n * H + (y > 0 ? y - 1 : 0)) * W;    // row y-1 (reflect: -1 -> 0)
    const float4 a = __ldg(reinterpret_cast<const float4 *>(r0 + x4));
    const float4 b = __ldg(reinterpret_cast<const float4 *>(r1 + x4));
    const int xl = x4 > 0 ? x4 - 1 : 0;                                  // column x-1 (reflect)
    const float top[5] = {__ldg(r0 + xl), a.x, a.y, a.z, a.w};
    const float bot[5] = {__ldg(r1 + xl), b.x, b.y, b.z, b.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float M1 = fmaxf(top[j], top[j + 1]), m1 = fminf(top[j], top[j + 1]);
        const float M2 = fmaxf(bot[j], bot[j + 1]), m2 = fminf(bot[j], bot[j + 1]);
        const float med = fmaxf(fminf(M1, M2), fmaxf(m1, m2));
        const float raw = bot[j + 1];
        o[j] = (fabsf(__fsub_rn(raw, med)) > threshold) ? med : raw;
    }
    *reinterpret_cast<float4 *>(out + ((size_t)n * H + y) * W + x4) = make_float4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------ ImageBGSubtract
// Least-squares fit of I(u,v) ~ k0 + k1 u + k2 v + k3 u^2 + k4 uv + k5 v^2 (u = column, v = row,
// pipeline.py:384-401).  The reference inverts the raw normal matrix; here the same surface is fitted in
// centred, scaled coordinates (s = (u - cu)/su, t = (v - cv)/sv), which is well conditioned, so the
// result agrees with the reference to ~1e-12 (tests) without reproducing its conditioning.
struct BgGram {
    double g[36];             // Gram matrix of the basis {1, s, t, s^2, st, t^2} over the pixel grid
    double cu, cv, su, sv;
};

__device__ __forceinline__ void bg_basis(double s, double t, double *b)
{
    b[0] = 1.0; b[1] = s; b[2] = t; b[3] = s * s; b[4] = s * t; b[5] = t * t;
}

// Each block sums whole rows: per row r0 = sum I, r1 = sum I s, r2 = sum I s^2 (3 FMAs per pixel), then
// the six moments pick up the row's t powers once per row.
__global__ void __launch_bounds__(PREP_THREADS)
bg_partial(const float *__restrict__ in, int H, int W, BgGram G, double *__restrict__ part)
{
    __shared__ double sm[PREP_THREADS / 32 * 6];
    const int n = blockIdx.y;
    const float *img = in + (size_t)n * H * W;
    const double isu = 1.0 / G.su, isv = 1.0 / G.sv;
    double v[6] = {0, 0, 0, 0, 0, 0};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // one warp per row, rows strided over all warps of the image
    for (int y = blockIdx.x * (PREP_THREADS / 32) + warp; y < H; y += PREP_BLOCKS * (PREP_THREADS / 32)) {
        const float *row = img + (size_t)y * W;
        double r0 = 0.0, r1 = 0.0, r2 = 0.0;
        if ((W & 3) == 0 && ((uintptr_t)img & 15) == 0) {
            // s = (x - cu) / su for this lane's four columns; the next group of columns is 128 further:
            // x is an exact small integer in fp64, so it is advanced by addition instead of a per-pixel
            // int -> double conversion (the conversion pipe, not HBM, bounded this pass)
            double xd = (double)(lane * 4);
            for (int x = lane * 4; x < W; x += 128, xd += 128.0) {
                const float4 f = __ldg(reinterpret_cast<const float4 *>(row + x));
                const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double I = (double)fv[j];
                    const double s = ((xd + (double)j) - G.cu) * isu;
                    const double Is = I * s;
                    r0 += I;
                    r1 += Is;
                    r2 = fma(Is, s, r2);
                }
            }
        } else
        for (int x = lane; x < W; x += 32) {
            const double I = (double)__ldg(row + x);
            const double s = ((double)x - G.cu) * isu;
            const double Is = I * s;
            r0 += I;
            r1 += Is;
            r2 = fma(Is, s, r2);
        }
        const double t = ((double)y - G.cv) * isv;
        v[0] += r0; v[1] += r1; v[3] += r2;
        v[2] = fma(t, r0, v[2]); v[4] = fma(t, r1, v[4]); v[5] = fma(t * t, r0, v[5]);
    }
    block_sum<6>(v, sm);
    if (threadIdx.x == 0)
        for (int k = 0; k < 6; ++k) part[((size_t)n * PREP_BLOCKS + blockIdx.x) * 6 + k] = v[k];
}

// one thread per image: reduce the partial moments, solve the 6x6 normal equations (Gaussian elimination
// with partial pivoting, fp64)
__global__ void bg_solve(const double *__restrict__ part, BgGram G, int nimg, double *__restrict__ coef)
{
    // one block of 6 warps per image: warp i reduces moment i, thread 0 solves
    __shared__ double rhs[6];
    const int n = blockIdx.x, i = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0, v[PREP_BLOCKS / 32];
#pragma unroll
    for (int k = 0; k < PREP_BLOCKS / 32; ++k) v[k] = part[((size_t)n * PREP_BLOCKS + lane + 32 * k) * 6 + i];
#pragma unroll
    for (int k = 0; k < PREP_BLOCKS / 32; ++k) s += v[k];
    s = warp_sum(s);
    if (lane == 0) rhs[i] = s;
    __syncthreads();
    if (threadIdx.x != 0) return;
    double a[6][7];
    for (int r = 0; r < 6; ++r) {
        for (int j = 0; j < 6; ++j) a[r][j] = G.g[r * 6 + j];
        a[r][6] = rhs[r];
    }
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        for (int r = c + 1; r < 6; ++r) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
        if (piv != c) for (int j = 0; j < 7; ++j) { const double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
        const double d = a[c][c];
        for (int r = c + 1; r < 6; ++r) {
            const double f = (d != 0.0) ? a[r][c] / d : 0.0;
            for (int j = c; j < 7; ++j) a[r][j] -= f * a[c][j];
        }
    }
    double k[6];
    for (int c = 5; c >= 0; --c) {
        double s = a[c][6];
        for (int j = c + 1; j < 6; ++j) s -= a[c][j] * k[j];
        k[c] = (a[c][c] != 0.0) ? s / a[c][c] : 0.0;
    }
    for (int i = 0; i < 6; ++i) coef[(size_t)n * 6 + i] = k[i];
}

template <typename OutT>
__global__ void __launch_bounds__(256)
bg_apply(const float *__restrict__ in, OutT *__restrict__ out, int H, int W, BgGram G,
         const double *__restrict__ coef)
{
    // grid (ceil(W/1024), H, n), 4 consecutive pixels per thread:
    // bg = (k0 + k2 t + k5 t^2) + s (k1 + k4 t + k3 s)
    const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x0 >= W) return;
    const int y = blockIdx.y, n = blockIdx.z;
    const double *k = coef + (size_t)n * 6;
    const double t = ((double)y - G.cv) * (1.0 / G.sv), isu = 1.0 / G.su;
    const double a = fma(t, fma(t, k[5], k[2]), k[0]), b = fma(t, k[4], k[1]), c = k[3];
    const size_t i0 = ((size_t)n * H + y) * W + x0;
    const bool vec = (W & 3) == 0 && ((uintptr_t)in & 15) == 0;
    float v[4];
    if (vec) {
        const float4 f = *reinterpret_cast<const float4 *>(in + i0);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
        for (int j = 0; j < 4; ++j) v[j] = (x0 + j < W) ? in[i0 + j] : 0.0f;
    }
    OutT o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double s = ((double)(x0 + j) - G.cu) * isu;
        o[j] = (OutT)((double)v[j] - fma(s, fma(s, c, b), a));
    }
    for (int j = 0; j < 4; ++j)
        if (x0 + j < W) out[i0 + j] = o[j];
}

template <typename T>
__global__ void cast_kernel(const T *__restrict__ in, float *__restrict__ out, long long count)
{
    // 4 values per thread (the caller guarantees 16-byte aligned buffers); scalar tail
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < count) {
        *reinterpret_cast<float4 *>(out + i) = load4<T>(in + i);
    } else {
        for (long long j = i; j < count; ++j) out[j] = (float)in[j];
    }
}

int check_stack(sq_handle_t h, int n, int hgt, int wid, int c)
{
    SQ_REQUIRE(h, SQ_EINVAL, "null handle");
    SQ_REQUIRE(n >= 1 && hgt >= 1 && wid >= 1 && c >= 1 && c <= 16, SQ_EINVAL,
               "image stack: bad geometry n=%d h=%d w=%d c=%d", n, hgt, wid, c);
    return SQ_OK;
}

}  // namespace

extern "C" int sq_image_cast(sq_handle_t h, const void *in, int in_dtype, float *out, long long count,
                             void *stream)
{
    SQ_REQUIRE(h && in && out && count >= 0, SQ_EINVAL, "image_cast: bad arguments");
    SQ_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((count + 1023) / 1024);
    if (count == 0) return SQ_OK;
    if (in_dtype == SQ_U8) cast_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)in, out, count);
    else if (in_dtype == SQ_U16) cast_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)in, out, count);
    else if (in_dtype == SQ_F32) SQ_CUDA(cudaMemcpyAsync(out, in, (size_t)count * 4, cudaMemcpyDeviceToDevice, st));
    else SQ_REQUIRE(false, SQ_EINVAL, "image_cast: in_dtype must be SQ_U8, SQ_U16 or SQ_F32");
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

extern "C" int sq_prep_workspace_bytes(sq_handle_t h, int n, int c, size_t *bytes)
{
    SQ_REQUIRE(h && bytes, SQ_EINVAL, "prep_workspace_bytes: null pointer");
    SQ_REQUIRE(n >= 1 && c >= 1, SQ_EINVAL, "prep_workspace_bytes: bad geometry");
    SqArena a(nullptr, 0);
    a.take<double>((size_t)n * c * PREP_BLOCKS * 6);
    a.take<double>((size_t)n * c * 8);
    *bytes = a.off;
    return SQ_OK;
}

template <typename T>
int norm_launch(const T *in, float *out, int n, long long npix, int c, double *part, float2 *stats, cudaStream_t st)
{
    norm_partial<T><<<dim3(PREP_BLOCKS, n, c), PREP_THREADS, 0, st>>>(in, npix, c, part);
    SQ_CHECK_LAUNCH();
    norm_final<<<(n * c + 3) / 4, 128, 0, st>>>(part, npix, n * c, stats);
    SQ_CHECK_LAUNCH();
    if (!out) return SQ_OK;                            // moments only: the consumer applies them itself
    norm_apply<T><<<dim3((unsigned)((npix * c + 1023) / 1024), n), 256, 0, st>>>(in, out, npix * c, c, stats);
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

// ImageNorm of single-channel raw frames with the result rounded to bf16 (RNE) -- what the UNet's first conv does
// with its float32 input anyway (bf16 contract): 2 B/px out instead of 4, and the fused first pair loads it as is.
template <typename T>
__global__ void norm_apply_bf16(const T *__restrict__ in, __nv_bfloat16 *__restrict__ out, long long per_image,
                                const float2 *__restrict__ stats)
{
    const int n = blockIdx.y;
    const float2 st = stats[n];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // four pixels per thread
    if (i >= (per_image >> 2)) return;
    float4 f = load4<T>(in + (size_t)n * per_image + 4 * i);
    f.x = __fdiv_rn(__fsub_rn(f.x, st.x), st.y);
    f.y = __fdiv_rn(__fsub_rn(f.y, st.x), st.y);
    f.z = __fdiv_rn(__fsub_rn(f.z, st.x), st.y);
    f.w = __fdiv_rn(__fsub_rn(f.w, st.x), st.y);
    __nv_bfloat162 lo = __floats2bfloat162_rn(f.x, f.y), hi = __floats2bfloat162_rn(f.z, f.w);
    uint2 o;
    o.x = *reinterpret_cast<unsigned *>(&lo);
    o.y = *reinterpret_cast<unsigned *>(&hi);
    reinterpret_cast<uint2 *>(out + (size_t)n * per_image)[i] = o;
}

// internal (the host pipeline of the UNet): uint16 frames -> normalised bf16 frames; hgt * wid must be a multiple of 4
int sq_image_norm_u16_to_bf16(sq_handle_t h, const uint16_t *in, void *out_bf16, int n, int hgt, int wid, void *ws,
                              size_t ws_bytes, cudaStream_t st)
{
    SqArena a(ws, ws_bytes);
    double *part = a.take<double>((size_t)n * PREP_BLOCKS * 6);
    float2 *stats = (float2 *)a.take<double>((size_t)n * 8);
    SQ_REQUIRE(a.ok(), SQ_ENOMEM, "image_norm: workspace %zu < %zu bytes", ws_bytes, a.off);
    const long long npix = (long long)hgt * wid;
    SQ_REQUIRE((npix & 3) == 0, SQ_EINVAL, "image_norm(bf16): frame size must be a multiple of 4 pixels");
    SQ_TRY(norm_launch(in, (float *)nullptr, n, npix, 1, part, stats, st));
    norm_apply_bf16<uint16_t><<<dim3((unsigned)((npix / 4 + 255) / 256), n), 256, 0, st>>>(in, (__nv_bfloat16 *)out_bf16, npix, stats);
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

// ImageNorm's moments of raw uint16 frames only (internal: the fused first pair of the UNet applies them in its
// loader); *stats = n float2 (mean, std) inside the workspace, valid in stream order
int sq_image_norm_stats_u16(sq_handle_t h, const uint16_t *in, int n, int hgt, int wid, void *ws, size_t ws_bytes,
                            cudaStream_t st, const float2 **stats_out)
{
    SqArena a(ws, ws_bytes);
    double *part = a.take<double>((size_t)n * PREP_BLOCKS * 6);
    float2 *stats = (float2 *)a.take<double>((size_t)n * 8);
    SQ_REQUIRE(a.ok(), SQ_ENOMEM, "image_norm: workspace %zu < %zu bytes", ws_bytes, a.off);
    *stats_out = stats;
    return norm_launch(in, (float *)nullptr, n, (long long)hgt * wid, 1, part, stats, st);
}

extern "C" int sq_image_norm_raw(sq_handle_t h, const void *in, int in_dtype, float *out, int n, int hgt,
                                 int wid, int c, void *ws, size_t ws_bytes, void *stream)
{
    SQ_TRY(check_stack(h, n, hgt, wid, c));
    SQ_REQUIRE(in && out && ws, SQ_EINVAL, "image_norm: null pointer");
    SQ_CUDA(cudaSetDevice(h->device));
    SqArena a(ws, ws_bytes);
    double *part = a.take<double>((size_t)n * c * PREP_BLOCKS * 6);
    float2 *stats = (float2 *)a.take<double>((size_t)n * c * 8);
    SQ_REQUIRE(a.ok(), SQ_ENOMEM, "image_norm: workspace %zu < %zu bytes", ws_bytes, a.off);
    cudaStream_t st = (cudaStream_t)stream;
    const long long npix = (long long)hgt * wid;
    if (in_dtype == SQ_F32) return norm_launch((const float *)in, out, n, npix, c, part, stats, st);
    if (in_dtype == SQ_U16) return norm_launch((const uint16_t *)in, out, n, npix, c, part, stats, st);
    if (in_dtype == SQ_U8) return norm_launch((const uint8_t *)in, out, n, npix, c, part, stats, st);
    SQ_REQUIRE(false, SQ_EINVAL, "image_norm: in_dtype must be SQ_F32, SQ_U16 or SQ_U8");
}

extern "C" int sq_image_norm(sq_handle_t h, const float *in, float *out, int n, int hgt, int wid, int c,
                             void *ws, size_t ws_bytes, void *stream)
{
    return sq_image_norm_raw(h, in, SQ_F32, out, n, hgt, wid, c, ws, ws_bytes, stream);
}

extern "C" int sq_image_outliers(sq_handle_t h, const float *in, float *out, int n, int hgt, int wid, int c,
                                 int size, double threshold, void *stream)
{
    SQ_TRY(check_stack(h, n, hgt, wid, c));
    SQ_REQUIRE(in && out && in != out, SQ_EINVAL, "image_outliers: needs distinct in / out buffers");
    SQ_REQUIRE(size >= 1 && size <= 5, SQ_EUNSUPPORTED, "image_outliers: median size %d not in 1..5", size);
    SQ_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    SQ_REQUIRE(hgt <= 65535 && n <= 65535, SQ_EINVAL, "image_outliers: more than 65535 rows / images");
    const dim3 grid((unsigned)((wid * c + 255) / 256), (unsigned)hgt, (unsigned)n);
    const float thr = (float)threshold;
    if (size == 2 && c == 1 && (wid & 3) == 0 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0) {
        outliers2_c1_kernel<<<dim3((unsigned)((wid / 4 + 255) / 256), (unsigned)hgt, (unsigned)n), 256, 0, st>>>(
            in, out, hgt, wid, thr);
        SQ_CHECK_LAUNCH();
        return SQ_OK;
    }
    switch (size) {
    case 1: outliers_kernel<1><<<grid, 256, 0, st>>>(in, out, hgt, wid, c, thr); break;
    case 2: outliers_kernel<2><<<grid, 256, 0, st>>>(in, out, hgt, wid, c, thr); break;
    case 3: outliers_kernel<3><<<grid, 256, 0, st>>>(in, out, hgt, wid, c, thr); break;
    case 4: outliers_kernel<4><<<grid, 256, 0, st>>>(in, out, hgt, wid, c, thr); break;
    default: outliers_kernel<5><<<grid, 256, 0, st>>>(in, out, hgt, wid, c, thr); break;
    }
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

extern "C" int sq_image_bgsubtract(sq_handle_t h, const float *in, void *out, int out_dtype, int n, int hgt,
                                   int wid, void *ws, size_t ws_bytes, void *stream)
{
    SQ_TRY(check_stack(h, n, hgt, wid, 1));
    SQ_REQUIRE(in && out && ws, SQ_EINVAL, "image_bgsubtract: null pointer");
    SQ_REQUIRE(out_dtype == SQ_F32 || out_dtype == SQ_F64, SQ_EINVAL, "image_bgsubtract: bad out_dtype");
    SQ_REQUIRE(hgt >= 3 && wid >= 3, SQ_EINVAL, "image_bgsubtract: a quadratic surface needs >= 3x3 pixels");
    SQ_CUDA(cudaSetDevice(h->device));
    SqArena a(ws, ws_bytes);
    double *part = a.take<double>((size_t)n * PREP_BLOCKS * 6);
    double *coef = a.take<double>((size_t)n * 8);
    SQ_REQUIRE(a.ok(), SQ_ENOMEM, "image_bgsubtract: workspace %zu < %zu bytes", ws_bytes, a.off);
    // Gram matrix of the centred basis: separable power sums over columns and rows
    BgGram G;
    G.cu = 0.5 * (wid - 1); G.cv = 0.5 * (hgt - 1);
    G.su = 0.5 * wid; G.sv = 0.5 * hgt;
    double px[5] = {0, 0, 0, 0, 0}, py[5] = {0, 0, 0, 0, 0};
    for (int x = 0; x < wid; ++x) { const double s = (x - G.cu) / G.su; double p = 1.0; for (int e = 0; e < 5; ++e) { px[e] += p; p *= s; } }
    for (int y = 0; y < hgt; ++y) { const double t = (y - G.cv) / G.sv; double p = 1.0; for (int e = 0; e < 5; ++e) { py[e] += p; p *= t; } }
    const int es[6] = {0, 1, 0, 2, 1, 0}, et[6] = {0, 0, 1, 0, 1, 2};   // basis i = s^es[i] * t^et[i]
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) G.g[i * 6 + j] = px[es[i] + es[j]] * py[et[i] + et[j]];
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)n * hgt * wid;
    bg_partial<<<dim3(PREP_BLOCKS, n), PREP_THREADS, 0, st>>>(in, hgt, wid, G, part);
    SQ_CHECK_LAUNCH();
    bg_solve<<<n, 192, 0, st>>>(part, G, n, coef);
    SQ_CHECK_LAUNCH();
    SQ_REQUIRE(hgt <= 65535 && n <= 65535, SQ_EINVAL, "image_bgsubtract: more than 65535 rows / images");
    const dim3 grid((unsigned)((wid + 1023) / 1024), (unsigned)hgt, (unsigned)n);
    if (out_dtype == SQ_F32) bg_apply<float><<<grid, 256, 0, st>>>(in, (float *)out, hgt, wid, G, coef);
    else bg_apply<double><<<grid, 256, 0, st>>>(in, (double *)out, hgt, wid, G, coef);
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

// Host-buffer form of the three pipes: which = 0 norm, 1 outliers, 2 background subtract.  `out_host`
// is float32 (n,h,w,c) except for background subtraction with out_dtype SQ_F64.
extern "C" int sq_image_pipe_host(sq_handle_t h, int which, const float *in_host, void *out_host, int out_dtype,
                                  int n, int hgt, int wid, int c, int size, double threshold)
{
    SQ_TRY(check_stack(h, n, hgt, wid, c));
    SQ_REQUIRE(in_host && out_host, SQ_EINVAL, "image_pipe_host: null pointer");
    SqHostCall call(h);
    SQ_REQUIRE(which >= 0 && which <= 2, SQ_EINVAL, "image_pipe_host: unknown pipe %d", which);
    SQ_REQUIRE(which != 2 || c == 1, SQ_EINVAL, "image_pipe_host: background subtraction takes one channel");
    SQ_CUDA(cudaSetDevice(h->device));
    const size_t count = (size_t)n * hgt * wid * c;
    const size_t osz = (which == 2 && out_dtype == SQ_F64) ? 8 : 4;
    size_t ws_bytes = 0;
    SQ_TRY(sq_prep_workspace_bytes(h, n, c, &ws_bytes));
    SQ_TRY(sq_reserve_device(h, sq_align_up(count * 4) + sq_align_up(count * osz) + ws_bytes + 1024));
    SqArena a(h->dev_arena, h->dev_arena_bytes);
    float *in = a.take<float>(count);
    char *out = a.take<char>(count * osz);
    void *ws = a.take<char>(ws_bytes);
    cudaStream_t st = h->stream;
    SQ_CUDA(cudaMemcpyAsync(in, in_host, count * 4, cudaMemcpyHostToDevice, st));
    if (which == 0) SQ_TRY(sq_image_norm(h, in, (float *)out, n, hgt, wid, c, ws, ws_bytes, st));
    else if (which == 1) SQ_TRY(sq_image_outliers(h, in, (float *)out, n, hgt, wid, c, size, threshold, st));
    else SQ_TRY(sq_image_bgsubtract(h, in, out, out_dtype, n, hgt, wid, ws, ws_bytes, st));
    SQ_CUDA(cudaMemcpyAsync(out_host, out, count * osz, cudaMemcpyDeviceToHost, st));
    SQ_CUDA(cudaStreamSynchronize(st));
    return SQ_OK;
}
