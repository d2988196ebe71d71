// Shared-memory tiled fp32 convolution on CUDA cores (SAME padding, channels-last, 2-D or 3-D), used by the exact
// inference path (unet.cu) and by the training step's data gradients (train.cu).
//
// Arithmetic contract (the one oracle/unet_ref.c fixes for the fp32 mode): every output is ONE chain
//     acc = fmaf(x[tap][ci], w[tap][ci][co], acc)   over taps (kz, ky, kx ascending), ci ascending (in0 then in1),
// followed by fmaf(acc, scale, shift) and the ReLU.  The tile walks K = (tap, ci) in that order without splitting
// it, so the result is bit-identical to the one-thread-per-pixel kernel; taps outside the frame contribute
// fmaf(0, w, acc) = acc.
//
// Block: 256 threads, 128 consecutive pixels x COT output channels; thread tile 4 pixels x COT/8 channels;
// K in chunks of 16 channels of one tap through shared memory.
#pragma once
#include <cuda_runtime.h>

namespace sqtile {

constexpr int PX = 128, KC = 16;

template <int COT>
__global__ void __launch_bounds__(256) conv_fp32_tile_kernel(const float *__restrict__ in0, int C0,
                                                             const float *__restrict__ in1, int C1, long long npix,
                                                             int D, int H, int W, const float *__restrict__ w, int KD,
                                                             int KH, int KW, int CO, const float *__restrict__ scale,
                                                             const float *__restrict__ shift, int relu,
                                                             float *__restrict__ out)
{
    constexpr int CT = COT / 8;
    __shared__ __align__(16) float As[KC][PX + 4];
    __shared__ __align__(16) float Ws[KC][COT];
    const int C = C0 + C1;
    const int t = threadIdx.x;
    const long long pbase = (long long)blockIdx.x * PX;
    const int co_base = blockIdx.y * COT;

    // the pixel this thread loads (two threads per pixel, 8 channels each)
    const int lp = t >> 1, lc = (t & 1) * 8;
    const long long p_ld = pbase + lp;
    const bool p_ok = p_ld < npix;
    int x = 0, y = 0, z = 0;
    long long nb = 0;
    if (p_ok) {
        const long long frame = (long long)W * H * D;
        const long long r = p_ld % frame;
        nb = p_ld - r;
        x = (int)(r % W);
        y = (int)((r / W) % H);
        z = (int)(r / ((long long)W * H));
    }
    // the outputs this thread owns
    const int pg = (t & 31) * 4, cg = (t >> 5) * CT;
    float acc[4][CT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CT; ++j) acc[i][j] = 0.0f;

    const int nchunk = (C + KC - 1) / KC;
    const int ntap = KD * KH * KW, niter = ntap * nchunk;
    constexpr int WPT = (KC * COT + 255) / 256;        // weight values a thread stages per chunk
    float av[8], wv[WPT];
    // global loads of iteration `it` (tap-major, then channel chunk: the oracle's order of K) into registers
    auto fetch = [&](int it) {
        const int tap = it / nchunk, c0 = (it - tap * nchunk) * KC;
        const int kx = tap % KW, ky = (tap / KW) % KH, kz = tap / (KW * KH);
        const int zz = z + kz - KD / 2, yy = y + ky - KH / 2, xx = x + kx - KW / 2;
        const bool inb = p_ok && zz >= 0 && zz < D && yy >= 0 && yy < H && xx >= 0 && xx < W;
        const long long pix = nb + ((long long)zz * H + yy) * W + xx;
        const float *wk = w + (size_t)tap * C * CO;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + lc + j;
            float v = 0.0f;
            if (inb && c < C) v = (c < C0) ? in0[pix * C0 + c] : in1[pix * C1 + (c - C0)];
            av[j] = v;
        }
#pragma unroll
        for (int j = 0; j < WPT; ++j) {
            const int e = t + j * 256;
            const int k = e / COT, co = e % COT;
            float v = 0.0f;
            if (e < KC * COT && c0 + k < C && co_base + co < CO) v = __ldg(wk + (size_t)(c0 + k) * CO + co_base + co);
            wv[j] = v;
        }
    };
    fetch(0);
    for (int it = 0; it < niter; ++it) {
        const int c0 = (it % nchunk) * KC;
        __syncthreads();                               // the previous chunk has been consumed
#pragma unroll
        for (int j = 0; j < 8; ++j) As[lc + j][lp] = av[j];
#pragma unroll
        for (int j = 0; j < WPT; ++j) {
            const int e = t + j * 256;
            if (e < KC * COT) Ws[e / COT][e % COT] = wv[j];
        }
        __syncthreads();
        if (it + 1 < niter) fetch(it + 1);             // in flight while this chunk is multiplied
        const int kn = min(KC, C - c0);
        auto step = [&](int k) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[k][pg]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float b[CT];
            if (CT == 2) {
                const float2 b2 = *reinterpret_cast<const float2 *>(&Ws[k][cg]);
                b[0] = b2.x; b[1] = b2.y;
            } else {
#pragma unroll
                for (int j = 0; j < CT; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(&Ws[k][cg + j]);
                    b[j] = b4.x; b[j + 1] = b4.y; b[j + 2] = b4.z; b[j + 3] = b4.w;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < CT; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        };
        if (kn == KC) {
#pragma unroll
            for (int k = 0; k < KC; ++k) step(k);
        } else {
            for (int k = 0; k < kn; ++k) step(k);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long p = pbase + pg + i;
        if (p >= npix) continue;
#pragma unroll
        for (int j = 0; j < CT; ++j) {
            const int co = co_base + cg + j;
            if (co < CO) {
                float v = acc[i][j];
                if (scale) v = fmaf(v, scale[co], shift[co]);
                if (relu && !(v > 0.0f)) v = 0.0f;
                out[p * CO + co] = v;
            }
        }
    }
}

// true when the tiled kernel serves this layer (otherwise the caller keeps its one-thread-per-pixel kernel)
static inline bool can_tile(int C, int CO) { return C >= 8 && CO >= 16; }

static inline cudaError_t launch(const float *in0, int C0, const float *in1, int C1, long long npix, int D, int H,
                                 int W, const float *w, int KD, int KH, int KW, int CO, const float *scale,
                                 const float *shift, int relu, float *out, cudaStream_t st)
{
    const unsigned gx = (unsigned)((npix + PX - 1) / PX);
    if (CO >= 64)
        conv_fp32_tile_kernel<64><<<dim3(gx, (CO + 63) / 64), 256, 0, st>>>(in0, C0, in1, C1, npix, D, H, W, w, KD, KH,
                                                                            KW, CO, scale, shift, relu, out);
    else if (CO >= 32)
        conv_fp32_tile_kernel<32><<<dim3(gx, (CO + 31) / 32), 256, 0, st>>>(in0, C0, in1, C1, npix, D, H, W, w, KD, KH,
                                                                            KW, CO, scale, shift, relu, out);
    else
        conv_fp32_tile_kernel<16><<<dim3(gx, (CO + 15) / 16), 256, 0, st>>>(in0, C0, in1, C1, npix, D, H, W, w, KD, KH,
                                                                            KW, CO, scale, shift, relu, out);
    return cudaGetLastError();
}

}  // namespace sqtile
