// Weighted softmax cross-entropy of the UNet head: loss and gradient w.r.t. the logits.
// Consumes the GPU weight map (sq_weightmap_*) as the per-pixel weights -- the training-step
// half of BASELINE config 5.  The reference prepares labels and weights in tr_augment
// (networks/unet.py:396-401) but does not ship the loss; definition in include/sequitr_b200.h.
// HBM-bound: (4K + 5) B/px in, 4K B/px out when the gradient is requested.
#include "sq_common.cuh"

namespace {

constexpr int CE_BLOCKS = 1024, CE_THREADS = 256;

__global__ void wce_kernel(const float *__restrict__ logits, const uint8_t *__restrict__ labels,
                           const float *__restrict__ weights, long long npix, int K,
                           double *__restrict__ partials, float *__restrict__ grad)
{
    __shared__ double sh[CE_THREADS / 32];
    const float inv_n = (float)(1.0 / (double)npix);
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * CE_THREADS + threadIdx.x; i < npix;
         i += (long long)CE_BLOCKS * CE_THREADS) {
        const float *l = logits + i * K;
        float v[16];
        float m = -INFINITY;
        for (int k = 0; k < K; ++k) { v[k] = l[k]; m = fmaxf(m, v[k]); }
        float s = 0.0f;
        for (int k = 0; k < K; ++k) { v[k] = expf(v[k] - m); s += v[k]; }
        const int y = labels[i];
        const float w = weights[i];
        const float lse = m + logf(s);
        acc += (double)w * (double)(lse - l[y]);
        if (grad) {
            const float g = w * inv_n / s;
            for (int k = 0; k < K; ++k) grad[i * K + k] = g * v[k] - (k == y ? w * inv_n : 0.0f);
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int wv = 0; wv < CE_THREADS / 32; ++wv) t += sh[wv];
        partials[blockIdx.x] = t;
    }
}

__global__ void wce_final(const double *__restrict__ partials, long long npix, double *__restrict__ loss)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < CE_BLOCKS; ++i) t += partials[i];      // fixed order
        *loss = t / (double)npix;
    }
}

}  // namespace

extern "C" int sq_weighted_ce_workspace_bytes(sq_handle_t h, size_t *bytes)
{
    SQ_REQUIRE(h && bytes, SQ_EINVAL, "weighted_ce: null pointer");
    *bytes = CE_BLOCKS * sizeof(double);
    return SQ_OK;
}

extern "C" int sq_weighted_ce(sq_handle_t h, const float *logits, const uint8_t *labels,
                              const float *weights, long long npix, int K, double *loss, float *grad,
                              void *ws, size_t ws_bytes, void *stream_)
{
    SQ_REQUIRE(h && logits && labels && weights && loss && ws, SQ_EINVAL, "weighted_ce: null pointer");
    SQ_REQUIRE(npix >= 1 && K >= 1 && K <= 16, SQ_EINVAL, "weighted_ce: need npix >= 1 and 1 <= K <= 16");
    SQ_REQUIRE(ws_bytes >= CE_BLOCKS * sizeof(double), SQ_ENOMEM, "weighted_ce: workspace too small");
    cudaStream_t st = (cudaStream_t)stream_;
    wce_kernel<<<CE_BLOCKS, CE_THREADS, 0, st>>>(logits, labels, weights, npix, K, (double *)ws, grad);
    wce_final<<<1, 32, 0, st>>>((const double *)ws, npix, loss);
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}
