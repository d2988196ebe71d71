// Training step of the UNet (BASELINE config 5 / reference networks/unet.py: the graph `build` :224-262 with
// `training` true (:170-172, dropout :274-276), fed by tr_augment :348-401 with (image, {'label', 'weights'})):
//
//     forward (fp32 path of unet.cu, every activation kept)  ->  weighted softmax cross-entropy (loss.cu)
//     ->  backward through head, up blocks, bridges, transposed convs, max pools, down blocks
//     ->  optimiser update of every kernel and bias, in place in the plan's device weights
//
// so the SAME plan object serves inference right after a step.  The reference ships neither the concrete layers nor
// a loss / optimiser for the UNet (conv_layer :326-329 raises NotImplementedError; the only optimiser in the repo is
// the GAN's Adam, gan.py:740-751); the layer definitions are the ones DESIGN.md §1 pins, the loss is the one loss.cu
// documents, the optimiser is TensorFlow's Adam update rule (or plain SGD).
//
// Arithmetic: fp32 on CUDA cores, fixed reduction orders (split partial sums, then one ordered pass) -- a first
// correct path, checked against a float64 autograd restatement (oracle/train_oracle.py).  The tensor-core version
// (dgrad = the forward tcgen05 conv on flipped weights, wgrad = an MN-major UMMA with split pixels) is not built.
#include "sq_common.cuh"
#include "unet_plan.cuh"
#include "conv_fp32_tile.cuh"

#include <algorithm>
#include <cmath>

struct sq_trainer_s {
    sq_unet_s *u = nullptr;
    int optimizer = 1;            // 0 = SGD, 1 = Adam
    float lr = 1e-3f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, dropout = 0.0f;
    unsigned long long seed = 0;
    long long step = 0;
    struct Slot { float *gw = nullptr, *gb = nullptr, *mw = nullptr, *vw = nullptr, *mb = nullptr, *vb = nullptr;
                  // a layer with a frozen per-channel affine keeps its trainable bias b and the frozen shift t here;
                  // the plan's epilogue shift is re-folded from them after every update
                  float *bias = nullptr, *tsh = nullptr;
                  bool affine = false;
                  size_t wcount = 0; };
    std::vector<Slot> slots;      // one per layer of u->layers
    float *wflip = nullptr;       // scratch for a layer's tap-reversed kernel (data gradients)
    float *garena = nullptr;      // all gradients, contiguous (one all-reduce for data-parallel training)
    size_t gcount = 0;
    std::vector<void *> allocs;
};

namespace {


// -------------------------------------------------------------------------------------------------- kernels
// d(pre-activation) from d(block output): ReLU mask, and the 1/(1-rate) of a dropout that followed it (the stored
// output is the dropped one: it is > 0 exactly where the unit was active AND kept)
// With a frozen per-channel affine (folded BN: y = relu((conv + b) * s + t)) the gradient of the pre-affine sum is
// that times s[c]: it feeds the weight gradient, the data gradient AND the bias gradient (dL/db = sum dz * s).
__global__ void relu_bwd_kernel(float *__restrict__ g, const float *__restrict__ y, long long n, float keep,
                                const float *__restrict__ scale, int C)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = (y[i] > 0.0f) ? __fdiv_rn(g[i], keep) : 0.0f;
    if (scale) v *= scale[(int)(i % C)];
    g[i] = v;
}

// the folded epilogue shift of a layer with a frozen affine, as sq_unet_finalize computes it: b * s + t, no fma
__global__ void refold_kernel(const float *__restrict__ b, const float *__restrict__ s, const float *__restrict__ t,
                              int C, float *__restrict__ shift)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) shift[c] = __fadd_rn(__fmul_rn(b[c], s[c]), t[c]);
}

// dx[p][ci] = sum_tap sum_co dz[p - off(tap)][co] * w[tap][ci][co]     (w HWIO / DHWIO, SAME padding)
template <int CIT>
__global__ void dgrad_kernel(const float *__restrict__ dz, int CO, long long npix, int D, int H, int W,
                             const float *__restrict__ w, int KD, int KH, int KW, int C, float *__restrict__ dx)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const int ci0 = blockIdx.y * CIT;
    const int x = (int)(p % W);
    const int y = (int)((p / W) % H);
    const int z = (int)((p / ((long long)W * H)) % D);
    const long long nb = p / ((long long)W * H * D) * ((long long)W * H * D);
    float acc[CIT];
#pragma unroll
    for (int j = 0; j < CIT; ++j) acc[j] = 0.0f;
    for (int kz = 0; kz < KD; ++kz) {
        const int zz = z - (kz - KD / 2);
        if (zz < 0 || zz >= D) continue;
        for (int ky = 0; ky < KH; ++ky) {
            const int yy = y - (ky - KH / 2);
            if (yy < 0 || yy >= H) continue;
            for (int kx = 0; kx < KW; ++kx) {
                const int xx = x - (kx - KW / 2);
                if (xx < 0 || xx >= W) continue;
                const float *q = dz + (nb + ((long long)zz * H + yy) * W + xx) * CO;
                const float *wk = w + ((size_t)((kz * KH + ky) * KW + kx) * C + ci0) * CO;
                for (int co = 0; co < CO; ++co) {
                    const float v = q[co];
#pragma unroll
                    for (int j = 0; j < CIT; ++j)
                        if (ci0 + j < C) acc[j] = fmaf(v, __ldg(wk + (size_t)j * CO + co), acc[j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CIT; ++j)
        if (ci0 + j < C) dx[p * C + ci0 + j] = acc[j];
}

// Weight gradients as split sums over pixels:  part[z][tap][a][b] = sum_{p in chunk z} A[pa(p, tap)][a] * B[p][b]
//   MODE 0 (3x3 / 1x1 conv):  A = the layer input (x0 | x1 concatenated), pa = p + off(tap), zero outside the frame;
//                             B = dz.                 -> (taps, Cin, Cout)   = HWIO
//   MODE 1 (2x2 stride-2 transposed conv):  A = d(output) at the fine pixel (2y+ky, 2x+kx) of coarse pixel p;
//                             B = the layer input.    -> (taps, Cout, Cin)   = TF (kh, kw, out, in)
// 256 threads work on a TA x TB tile of (a, b) with 4 x 4 values per thread; a tile smaller than 64 x 64 leaves
// threads over, which split the 64 pixels of a shared-memory stage between KG groups (summed in group order at the
// end), so that narrow layers (16 or 32 channels, where most of the pixels are) do not pay for a 64 x 64 tile.
template <int MODE, int TA, int TB>
__global__ void __launch_bounds__(256) wgrad_kernel(const float *__restrict__ A0, int CA0,
                                                    const float *__restrict__ A1, int CA1,
                                                    const float *__restrict__ B, int CB, long long npix, int D,
                                                    int H, int W, int KD, int KH, int KW, long long chunk,
                                                    float *__restrict__ part)
{
    constexpr int KS = 64;                         // pixels per stage
    constexpr int TPG = (TA / 4) * (TB / 4);       // threads per group
    constexpr int KG = 256 / TPG;                  // groups splitting the stage's pixels
    __shared__ __align__(16) float As[KS][TA + 4];
    __shared__ __align__(16) float Bs[KS][TB + 4];
    __shared__ long long pa_s[KS];
    __shared__ __align__(16) float red[(KG > 1) ? KG * TA * TB : 1];
    const int CA = CA0 + CA1;
    const int btiles = (CB + TB - 1) / TB;
    const int a0 = (blockIdx.x / btiles) * TA, b0 = (blockIdx.x % btiles) * TB;
    const int tap = blockIdx.y, ntap = gridDim.y;
    const int kx = tap % KW, ky = (tap / KW) % KH, kz = tap / (KW * KH);
    const long long p_lo = (long long)blockIdx.z * chunk, p_hi = min(npix, p_lo + chunk);
    const int t = threadIdx.x;
    const int grp = t / TPG, idx = t % TPG;
    const int ty = idx / (TB / 4), tx = idx % (TB / 4);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    const long long frame = (long long)W * H * D;
    constexpr int NA = KS * TA / 256, NB = KS * TB / 256;      // values a thread stages per stage
    float ra[NA], rb[NB];
    // the A-pixel of stage pixel k (or -1 outside the frame), computed by thread k and shared
    auto pixel_map = [&](long long p0) {
        if (t < KS) {
            const long long p = p0 + t;
            long long pa = -1;
            if (p < p_hi) {
                const long long n = p / frame;
                const int r = (int)(p - n * frame);
                const int x = r % W, y = (r / W) % H, z = r / (W * H);
                if (MODE == 0) {
                    const int xx = x + kx - KW / 2, yy = y + ky - KH / 2, zz = z + kz - KD / 2;
                    if (xx >= 0 && xx < W && yy >= 0 && yy < H && zz >= 0 && zz < D)
                        pa = ((n * D + zz) * H + yy) * W + xx;
                } else {
                    pa = ((n * (D * KD) + (long long)z * KD + kz) * (2 * H) + 2 * y + ky) * (2 * W) + 2 * x + kx;
                }
            }
            pa_s[t] = pa;
        }
    };
    auto fetch = [&](long long p0) {
#pragma unroll
        for (int q = 0; q < NA; ++q) {
            const int e = t + q * 256;
            const int k = e / TA, a = a0 + e % TA;
            const long long pa = pa_s[k];
            float v = 0.0f;
            if (pa >= 0) {
                if (a < CA0) v = A0[pa * CA0 + a];
                else if (a < CA) v = A1[pa * CA1 + (a - CA0)];
            }
            ra[q] = v;
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int e = t + q * 256;
            const int k = e / TB, b = b0 + e % TB;
            const long long p = p0 + k;
            rb[q] = (p < p_hi && b < CB) ? B[p * CB + b] : 0.0f;
        }
    };
    if (p_lo < p_hi) {
        pixel_map(p_lo);
        __syncthreads();
        fetch(p_lo);
    }
    for (long long p0 = p_lo; p0 < p_hi; p0 += KS) {
        __syncthreads();                           // the previous stage has been consumed (and pa_s read)
#pragma unroll
        for (int q = 0; q < NA; ++q) { const int e = t + q * 256; As[e / TA][e % TA] = ra[q]; }
#pragma unroll
        for (int q = 0; q < NB; ++q) { const int e = t + q * 256; Bs[e / TB][e % TB] = rb[q]; }
        const bool more = p0 + KS < p_hi;
        if (more) pixel_map(p0 + KS);
        __syncthreads();
        if (more) fetch(p0 + KS);                  // in flight while this stage is multiplied
#pragma unroll 4
        for (int k = grp; k < KS; k += KG) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    float *o = part + ((size_t)blockIdx.z * ntap + tap) * CA * CB;
    if (KG == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int a = a0 + ty * 4 + i, b = b0 + tx * 4 + j;
                if (a < CA && b < CB) o[(size_t)a * CB + b] = acc[i][j];
            }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[(grp * TA + ty * 4 + i) * TB + tx * 4 + j] = acc[i][j];
        __syncthreads();
        for (int e = t; e < TA * TB; e += 256) {
            float sum = 0.0f;
            for (int g2 = 0; g2 < KG; ++g2) sum += red[g2 * TA * TB + e];
            const int a = a0 + e / TB, b = b0 + e % TB;
            if (a < CA && b < CB) o[(size_t)a * CB + b] = sum;
        }
    }
}

// kernels with the taps reversed and the channel roles swapped: wT[taps-1-tap][co][ci] = w[tap][ci][co], so that the
// data gradient of a SAME convolution is the forward convolution of dz with wT
__global__ void flip_weights_kernel(const float *__restrict__ w, int taps, int C, int CO, float *__restrict__ wt)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)taps * C * CO) return;
    const int co = (int)(i % CO);
    const int ci = (int)((i / CO) % C);
    const int tap = (int)(i / ((size_t)CO * C));
    wt[((size_t)(taps - 1 - tap) * CO + co) * C + ci] = w[i];
}

// part[z][c] = sum over the pixels of chunk z of g[p][c]; 32 channels x 8 pixel lanes per block
__global__ void bias_partial_kernel(const float *__restrict__ g, int C, long long npix, long long chunk,
                                    float *__restrict__ part)
{
    __shared__ float sh[8][33];
    const int c = blockIdx.y * 32 + threadIdx.x;
    const long long p_lo = (long long)blockIdx.x * chunk, p_hi = min(npix, p_lo + chunk);
    float s = 0.0f;
    if (c < C)
        for (long long p = p_lo + threadIdx.y; p < p_hi; p += 8) s += g[p * C + c];
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float tsum = 0.0f;
        for (int k = 0; k < 8; ++k) tsum += sh[k][threadIdx.x];
        part[(size_t)blockIdx.x * C + c] = tsum;
    }
}

// out[i] = sum_z part[z][i] in the order of z
__global__ void reduce_splits_kernel(const float *__restrict__ part, size_t count, int nsplit, float *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s = 0.0f;
    for (int z = 0; z < nsplit; ++z) s += part[(size_t)z * count + i];
    out[i] = s;
}

// max-pool backward: the first maximum of the window (scan order dz, dy, dx) receives the gradient
__global__ void maxpool_bwd_kernel(const float *__restrict__ in, const float *__restrict__ gout, long long nout,
                                   int D, int H, int W, int C, int pool_d, int accumulate, float *__restrict__ gin)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nout) return;
    const int Do = D / pool_d, Ho = H / 2, Wo = W / 2;
    const int c = (int)(i % C);
    long long r = i / C;
    const int x = (int)(r % Wo); r /= Wo;
    const int y = (int)(r % Ho); r /= Ho;
    const int z = (int)(r % Do);
    const long long n = r / Do;
    float m = -INFINITY;
    int best = 0;
    for (int dz = 0; dz < pool_d; ++dz)
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                const float v = in[((((long long)n * D + z * pool_d + dz) * H + 2 * y + dy) * W + 2 * x + dx) * C + c];
                if (v > m) { m = v; best = (dz * 2 + dy) * 2 + dx; }
            }
    const float g = gout[i];
    for (int dz = 0; dz < pool_d; ++dz)
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                const long long q = ((((long long)n * D + z * pool_d + dz) * H + 2 * y + dy) * W + 2 * x + dx) * C + c;
                const float v = ((dz * 2 + dy) * 2 + dx == best) ? g : 0.0f;
                gin[q] = accumulate ? gin[q] + v : v;
            }
}

// transposed-conv backward to its input: gin[p][ci] = sum_tap sum_co gout[fine(p, tap)][co] * w[tap][co][ci]
__global__ void upconv_dgrad_kernel(const float *__restrict__ gout, long long nin, int D, int H, int W, int CI,
                                    int up_d, const float *__restrict__ w, int CO, float *__restrict__ gin)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nin) return;
    const int ci = (int)(i % CI);
    long long r = i / CI;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H); r /= H;
    const int z = (int)(r % D);
    const long long n = r / D;
    float acc = 0.0f;
    for (int kz = 0; kz < up_d; ++kz)
        for (int ky = 0; ky < 2; ++ky)
            for (int kx = 0; kx < 2; ++kx) {
                const int tap = (kz * 2 + ky) * 2 + kx;
                const float *q = gout + ((((long long)n * D * up_d + z * up_d + kz) * (2 * H) + 2 * y + ky) * (2 * W) +
                                         2 * x + kx) * CO;
                const float *wr = w + (size_t)tap * CO * CI + ci;
                for (int co = 0; co < CO; ++co) acc = fmaf(q[co], __ldg(wr + (size_t)co * CI), acc);
            }
    gin[i] = acc;
}

// bridge backward (networks/unet.py BRIDGE_TYPES): merged = a op b
__global__ void eltwise_bwd_kernel(const float *__restrict__ g, const float *__restrict__ a, const float *__restrict__ b,
                                   long long n, int op, float *__restrict__ ga, float *__restrict__ gb)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = g[i];
    if (op == SQ_BRIDGE_ADD) { ga[i] = v; gb[i] = v; }
    else if (op == SQ_BRIDGE_SUB) { ga[i] = v; gb[i] = -v; }
    else { ga[i] = v * b[i]; gb[i] = v * a[i]; }
}

__global__ void slice_kernel(const float *__restrict__ src, int Cs, int off, long long npix, int Cd,
                             float *__restrict__ dst)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix * Cd) return;
    const long long p = i / Cd;
    const int c = (int)(i % Cd);
    dst[i] = src[p * Cs + off + c];
}

// tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); m, v exponential averages;
// p -= lr_t * m / (sqrt(v) + eps).  optimizer 0: p -= lr * g.
__global__ void update_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                              float *__restrict__ v, size_t n, int optimizer, float lr, float lr_t, float b1, float b2,
                              float eps)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    if (optimizer == 0) { p[i] = p[i] - lr * gi; return; }
    const float mi = m[i] + (gi - m[i]) * (1.0f - b1);
    const float vi = v[i] + (gi * gi - v[i]) * (1.0f - b2);
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - __fdiv_rn(lr_t * mi, sqrtf(vi) + eps);
}

// -------------------------------------------------------------------------------------------------- host side
struct Geo { int n, d, h, w; };

long long level_px(const sq_unet_s *u, const Geo &g, int l)
{
    const int dl = (u->ndim == 3) ? (g.d >> l) : 1;
    return (long long)g.n * dl * (g.h >> l) * (g.w >> l);
}

int layer_index(const sq_unet_s *u, const std::string &scope)
{
    for (size_t i = 0; i < u->layers.size(); ++i)
        if (u->layers[i].scope == scope) return (int)i;
    return -1;
}

size_t layer_wcount(const sq_unet_s *u, const SqLayer &L)
{
    size_t taps;
    if (L.kind == SqLayer::UPCONV) taps = (u->ndim == 3) ? 8 : 4;
    else taps = (u->ndim == 3) ? (size_t)L.ksize * L.ksize * L.ksize : (size_t)L.ksize * L.ksize;
    return taps * (size_t)(L.cin0 + L.cin1) * L.cout;
}

// how many pixel chunks a weight / bias gradient is split into (>= ~4 blocks per SM overall, chunks >= 256 pixels)
int pick_splits(const sq_unet_s *u, long long npix, long long blocks_per_split)
{
    const long long want = (4LL * u->h->sm_count + blocks_per_split - 1) / blocks_per_split;
    const long long cap = std::max<long long>(1, npix / 256);
    return (int)std::max<long long>(1, std::min<long long>(std::min<long long>(want, cap), 512));
}

struct Scratch { float *part; size_t part_count; };

// weight + bias gradient of a conv-like layer.  conv: x0|x1 the inputs, dz the pre-activation gradient (both on the
// layer's pixel grid).  upconv: x0 = the coarse input, dz = the fine output gradient; (D,H,W) the COARSE grid.
int param_grads(sq_trainer_s *tr, int li, const float *x0, const float *x1, const float *dz, long long npix_x,
                int D, int H, int W, const Scratch &sc, cudaStream_t st)
{
    sq_unet_s *u = tr->u;
    const SqLayer &L = u->layers[li];
    sq_trainer_s::Slot &s = tr->slots[li];
    const bool up = L.kind == SqLayer::UPCONV;
    const int KD = up ? ((u->ndim == 3) ? 2 : 1) : ((u->ndim == 3) ? L.ksize : 1);
    const int KH = up ? 2 : L.ksize, KW = KH;
    const int taps = KD * KH * KW;
    const int CA = up ? L.cout : L.cin0 + L.cin1, CB = up ? L.cin0 : L.cout;
    const int TA = CA > 32 ? 64 : (CA > 16 ? 32 : 16), TB = CB > 32 ? 64 : (CB > 16 ? 32 : 16);
    const int tiles = ((CA + TA - 1) / TA) * ((CB + TB - 1) / TB);
    int nsplit = pick_splits(u, npix_x, (long long)tiles * taps);
    while (nsplit > 1 && (size_t)nsplit * s.wcount > sc.part_count) --nsplit;
    SQ_REQUIRE((size_t)nsplit * s.wcount <= sc.part_count, SQ_ENOMEM, "trainer: scratch too small for '%s'",
               L.scope.c_str());
    const long long chunk = ((npix_x + nsplit - 1) / nsplit + 63) / 64 * 64;
    dim3 grid((unsigned)tiles, (unsigned)taps, (unsigned)nsplit);
#define SQ_WGRAD(MODE_, TA_, TB_, ...) wgrad_kernel<MODE_, TA_, TB_><<<grid, 256, 0, st>>>(__VA_ARGS__)
#define SQ_WGRAD_TB(MODE_, TA_, ...)                                            \
    do {                                                                        \
        if (TB == 64) SQ_WGRAD(MODE_, TA_, 64, __VA_ARGS__);                    \
        else if (TB == 32) SQ_WGRAD(MODE_, TA_, 32, __VA_ARGS__);               \
        else SQ_WGRAD(MODE_, TA_, 16, __VA_ARGS__);                             \
    } while (0)
#define SQ_WGRAD_ANY(MODE_, ...)                                                \
    do {                                                                        \
        if (TA == 64) SQ_WGRAD_TB(MODE_, 64, __VA_ARGS__);                      \
        else if (TA == 32) SQ_WGRAD_TB(MODE_, 32, __VA_ARGS__);                 \
        else SQ_WGRAD_TB(MODE_, 16, __VA_ARGS__);                               \
    } while (0)
    if (up)
        SQ_WGRAD_ANY(1, dz, L.cout, nullptr, 0, x0, L.cin0, npix_x, D, H, W, KD, KH, KW, chunk, sc.part);
    else
        SQ_WGRAD_ANY(0, x0, L.cin0, x1, L.cin1, dz, L.cout, npix_x, D, H, W, KD, KH, KW, chunk, sc.part);
#undef SQ_WGRAD_ANY
#undef SQ_WGRAD_TB
#undef SQ_WGRAD
    SQ_CHECK_LAUNCH();
    reduce_splits_kernel<<<(unsigned)((s.wcount + 255) / 256), 256, 0, st>>>(sc.part, s.wcount, nsplit, s.gw);
    SQ_CHECK_LAUNCH();
    // bias: sum of dz over its own pixel grid (the fine grid for a transposed conv)
    const long long npix_z = up ? npix_x * taps : npix_x;
    int bsplit = pick_splits(u, npix_z, (L.cout + 31) / 32);
    while (bsplit > 1 && (size_t)bsplit * L.cout > sc.part_count) --bsplit;
    const long long bchunk = (npix_z + bsplit - 1) / bsplit;
    bias_partial_kernel<<<dim3((unsigned)bsplit, (unsigned)((L.cout + 31) / 32)), dim3(32, 8), 0, st>>>(
        dz, L.cout, npix_z, bchunk, sc.part);
    SQ_CHECK_LAUNCH();
    reduce_splits_kernel<<<(unsigned)((L.cout + 255) / 256), 256, 0, st>>>(sc.part, (size_t)L.cout, bsplit, s.gb);
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

int conv_dgrad(sq_trainer_s *tr, int li, const float *dz, long long npix, int D, int H, int W, float *dx,
               cudaStream_t st)
{
    sq_unet_s *u = tr->u;
    const SqLayer &L = u->layers[li];
    const int KD = (u->ndim == 3) ? L.ksize : 1;
    const int C = L.cin0 + L.cin1;
    const unsigned gx = (unsigned)((npix + 127) / 128);
    if (sqtile::can_tile(L.cout, C) && !getenv("SQ_FP32_NOTILE")) {
        // dx = SAME convolution of dz with the tap-reversed, channel-swapped kernel (conv_fp32_tile.cuh)
        const int taps = KD * L.ksize * L.ksize;
        const size_t cnt = (size_t)taps * C * L.cout;
        flip_weights_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(L.w, taps, C, L.cout, tr->wflip);
        SQ_CHECK_LAUNCH();
        SQ_CUDA(sqtile::launch(dz, L.cout, nullptr, 0, npix, D, H, W, tr->wflip, KD, L.ksize, L.ksize, C, nullptr,
                               nullptr, 0, dx, st));
        return SQ_OK;
    }
    if (C >= 8)
        dgrad_kernel<8><<<dim3(gx, (C + 7) / 8), 128, 0, st>>>(dz, L.cout, npix, D, H, W, L.w, KD, L.ksize, L.ksize, C, dx);
    else
        dgrad_kernel<1><<<dim3(gx, C), 128, 0, st>>>(dz, L.cout, npix, D, H, W, L.w, KD, L.ksize, L.ksize, C, dx);
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

// li: the conv layer whose output y is (its frozen affine scale, if any, is applied to the gradient)
int relu_bwd(sq_trainer_s *tr, int li, float *g, const float *y, long long n, float keep, cudaStream_t st)
{
    const bool aff = tr->slots[li].affine;
    relu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, y, n, keep, aff ? tr->u->layers[li].scale : nullptr,
                                                               tr->u->layers[li].cout);
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

// arena plan of one step: forward workspace, then gradient buffers that mirror the tape
struct StepPlan {
    size_t fwd_bytes = 0, total = 0;
    std::vector<float *> g_down, g_tmp, g_pooled, g_up, g_merged, g_upt, g_upo;
    float *g_logits = nullptr, *g_cat = nullptr, *g_in = nullptr;
    void *fwd = nullptr, *ce_ws = nullptr;
    size_t ce_bytes = 0;
    Scratch sc{nullptr, 0};
};

int plan_step(sq_trainer_s *tr, const Geo &g, void *ws, size_t ws_bytes, StepPlan *sp)
{
    sq_unet_s *u = tr->u;
    const int nl = u->nlev;
    SQ_TRY(sq_fp32_workspace(u, g.n, g.d, g.h, g.w, &sp->fwd_bytes));
    SQ_TRY(sq_weighted_ce_workspace_bytes(u->h, &sp->ce_bytes));
    SqArena a(ws, ws ? ws_bytes : 0);
    sp->fwd = a.take<char>(sp->fwd_bytes);
    sp->ce_ws = a.take<char>(sp->ce_bytes);
    sp->g_down.assign(nl, nullptr); sp->g_tmp.assign(nl, nullptr); sp->g_pooled.assign(nl, nullptr);
    sp->g_up.assign(nl, nullptr); sp->g_merged.assign(nl, nullptr); sp->g_upt.assign(nl, nullptr);
    sp->g_upo.assign(nl, nullptr);
    size_t cat_max = 0;
    for (int l = 0; l < nl; ++l) {
        const size_t px = (size_t)level_px(u, g, l), f = (size_t)u->filters[l];
        sp->g_tmp[l] = a.take<float>(px * f);
        sp->g_down[l] = a.take<float>(px * f);
        if (l > 0) sp->g_pooled[l] = a.take<float>(px * (size_t)u->filters[l - 1]);
        if (l < nl - 1) {
            sp->g_up[l] = a.take<float>(px * f);
            if (u->bridge >= SQ_BRIDGE_ADD && u->bridge <= SQ_BRIDGE_SUB) sp->g_merged[l] = a.take<float>(px * f);
            sp->g_upt[l] = a.take<float>(px * f);
            sp->g_upo[l] = a.take<float>(px * f);
            if (u->bridge == SQ_BRIDGE_CONCAT) cat_max = std::max(cat_max, px * 2 * f);
        }
    }
    sp->g_logits = a.take<float>((size_t)level_px(u, g, 0) * u->nout);
    if (cat_max) sp->g_cat = a.take<float>(cat_max);
    size_t wmax = 1024;
    for (const auto &s : tr->slots) wmax = std::max(wmax, s.wcount);
    // room for 8 splits of the largest kernel, or 512 splits of a small one
    sp->sc.part_count = std::max<size_t>(wmax * 8, (size_t)512 * 4096);
    sp->sc.part = a.take<float>(sp->sc.part_count);
    sp->total = a.off;
    return SQ_OK;
}

}  // namespace

// ================================================================================================== C ABI
extern "C" int sq_trainer_create(sq_unet_t u, int optimizer, float learning_rate, float beta1, float beta2,
                                 float epsilon, float dropout, unsigned long long seed, sq_trainer_t *out)
{
    SQ_REQUIRE(u && out, SQ_EINVAL, "trainer_create: null pointer");
    SQ_REQUIRE(u->finalized, SQ_ESTATE, "trainer_create: plan not finalised");
    SQ_REQUIRE(u->mode == SQ_MODE_FP32_EXACT, SQ_EUNSUPPORTED,
               "trainer_create: the training step runs on an fp32 plan (compute='fp32'); load the trained weights "
               "into a bf16 plan for tensor-core inference");
    SQ_REQUIRE(optimizer == 0 || optimizer == 1, SQ_EINVAL, "trainer_create: optimizer 0 (SGD) or 1 (Adam)");
    SQ_REQUIRE(dropout >= 0.0f && dropout < 1.0f, SQ_EINVAL, "trainer_create: dropout rate in [0, 1)");
    SQ_CUDA(cudaSetDevice(u->h->device));
    sq_trainer_s *tr = new sq_trainer_s();
    tr->u = u; tr->optimizer = optimizer; tr->lr = learning_rate; tr->beta1 = beta1; tr->beta2 = beta2;
    tr->eps = epsilon; tr->dropout = dropout; tr->seed = seed;
    tr->slots.resize(u->layers.size());
    auto zalloc = [&](size_t count, float **p) -> int {
        SQ_CUDA(cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(float)));
        tr->allocs.push_back(*p);
        SQ_CUDA(cudaMemset(*p, 0, std::max<size_t>(count, 1) * sizeof(float)));
        return SQ_OK;
    };
    // every gradient lives in ONE arena (kernel, bias, kernel, bias, ... in layer order; 64-float aligned pieces), so
    // that a data-parallel caller all-reduces the step's gradients with a single collective
    size_t gtotal = 0;
    for (size_t i = 0; i < u->layers.size(); ++i) {
        tr->slots[i].wcount = layer_wcount(u, u->layers[i]);
        gtotal += (tr->slots[i].wcount + 63) / 64 * 64 + ((size_t)u->layers[i].cout + 63) / 64 * 64;
    }
    if (zalloc(gtotal, &tr->garena) != SQ_OK) {
        delete tr;
        return SQ_ECUDA;
    }
    tr->gcount = gtotal;
    size_t goff = 0;
    for (size_t i = 0; i < u->layers.size(); ++i) {
        const SqLayer &L = u->layers[i];
        sq_trainer_s::Slot &s = tr->slots[i];
        s.gw = tr->garena + goff; goff += (s.wcount + 63) / 64 * 64;
        s.gb = tr->garena + goff; goff += ((size_t)L.cout + 63) / 64 * 64;
        int rc = SQ_OK;
        auto sc_it = u->host.find(L.scope + "/scale");
        if (sc_it != u->host.end()) {
            s.affine = true;
            rc = zalloc(L.cout, &s.bias);
            if (rc == SQ_OK) rc = zalloc(L.cout, &s.tsh);
            if (rc == SQ_OK) {
                const auto &b = u->host.at(L.scope + "/bias").data, &t = u->host.at(L.scope + "/shift").data;
                if (cudaMemcpy(s.bias, b.data(), L.cout * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
                    cudaMemcpy(s.tsh, t.data(), L.cout * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
                    sq_set_error("trainer_create: upload of '%s' failed", L.scope.c_str());
                    rc = SQ_ECUDA;
                }
            }
        }
        if (rc == SQ_OK && optimizer == 1) {
            rc = zalloc(s.wcount, &s.mw);
            if (rc == SQ_OK) rc = zalloc(s.wcount, &s.vw);
            if (rc == SQ_OK) rc = zalloc(L.cout, &s.mb);
            if (rc == SQ_OK) rc = zalloc(L.cout, &s.vb);
        }
        if (rc != SQ_OK) {
            for (void *p : tr->allocs) cudaFree(p);
            delete tr;
            return rc;
        }
    }
    size_t wmax = 1;
    for (const auto &s : tr->slots) wmax = std::max(wmax, s.wcount);
    if (zalloc(wmax, &tr->wflip) != SQ_OK) {
        for (void *p : tr->allocs) cudaFree(p);
        delete tr;
        return SQ_ECUDA;
    }
    *out = tr;
    return SQ_OK;
}

extern "C" int sq_trainer_destroy(sq_trainer_t tr)
{
    if (!tr) return SQ_OK;
    for (void *p : tr->allocs) cudaFree(p);
    delete tr;
    return SQ_OK;
}

extern "C" int sq_trainer_workspace_bytes(sq_trainer_t tr, int n, int d, int hgt, int wid, size_t *bytes)
{
    SQ_REQUIRE(tr && bytes, SQ_EINVAL, "trainer_workspace_bytes: null pointer");
    StepPlan sp;
    SQ_TRY(plan_step(tr, Geo{n, d, hgt, wid}, nullptr, 0, &sp));
    *bytes = sp.total;
    return SQ_OK;
}

extern "C" int sq_trainer_step(sq_trainer_t tr, const float *image_dev, const uint8_t *labels_dev,
                               const float *weights_dev, int n, int d, int hgt, int wid, int apply_update,
                               double *loss_dev, void *ws, size_t ws_bytes, void *stream_)
{
    SQ_REQUIRE(tr && image_dev && labels_dev && weights_dev && loss_dev && ws, SQ_EINVAL, "trainer_step: null pointer");
    sq_unet_s *u = tr->u;
    cudaStream_t st = (cudaStream_t)stream_;
    SQ_CUDA(cudaSetDevice(u->h->device));
    const Geo g{n, d, hgt, wid};
    StepPlan sp;
    SQ_TRY(plan_step(tr, g, ws, ws_bytes, &sp));
    SQ_REQUIRE(sp.total <= ws_bytes, SQ_ENOMEM, "trainer_step: workspace %zu < %zu bytes", ws_bytes, sp.total);
    const int nl = u->nlev;
    const int pd = (u->ndim == 3) ? 2 : 1;

    // ---- forward, every activation kept; dropout seeded per step
    SqTape tape;
    tape.drop_rate = tr->dropout;
    tape.seed = tr->seed + 0x9E3779B97F4A7C15ull * (unsigned long long)tr->step;
    SQ_TRY(sq_fp32_forward_tape(u, image_dev, n, d, hgt, wid, sp.fwd, sp.fwd_bytes, st, &tape));
    const float keep = 1.0f - tr->dropout;

    // ---- loss and d(loss)/d(logits)
    const long long px0 = level_px(u, g, 0);
    SQ_TRY(sq_weighted_ce(u->h, tape.logits, labels_dev, weights_dev, px0, u->nout, loss_dev, sp.g_logits, sp.ce_ws,
                          sp.ce_bytes, st));

    auto dims = [&](int l, int *D, int *H, int *W) {
        *D = (u->ndim == 3) ? (d >> l) : 1; *H = hgt >> l; *W = wid >> l;
    };
    auto idx = [&](const char *fmt, int l) {
        char scope[64];
        snprintf(scope, sizeof scope, fmt, l);
        return layer_index(u, scope);
    };
    int D, H, W;

    // ---- head (1x1 conv, no activation)
    const int ih = layer_index(u, "UNet/to_image");
    const float *top = (nl > 1) ? tape.upo[0] : tape.down[0];
    float *g_top = (nl > 1) ? sp.g_upo[0] : sp.g_down[0];
    dims(0, &D, &H, &W);
    SQ_TRY(param_grads(tr, ih, top, nullptr, sp.g_logits, px0, D, H, W, sp.sc, st));
    SQ_TRY(conv_dgrad(tr, ih, sp.g_logits, px0, D, H, W, g_top, st));

    // ---- up blocks, top to bottom (the reverse of the forward's bottom-to-top order)
    for (int l = 0; l <= nl - 2; ++l) {
        const long long px = level_px(u, g, l);
        const int f = u->filters[l];
        dims(l, &D, &H, &W);
        const int i2 = idx("UNet/up%d/conv2", l), i1 = idx("UNet/up%d/conv1", l), iu = idx("UNet/up%d/upscale", l);
        // conv2 (+ dropout)
        SQ_TRY(relu_bwd(tr, i2, sp.g_upo[l], tape.upo[l], px * f, keep, st));
        SQ_TRY(param_grads(tr, i2, tape.upt[l], nullptr, sp.g_upo[l], px, D, H, W, sp.sc, st));
        SQ_TRY(conv_dgrad(tr, i2, sp.g_upo[l], px, D, H, W, sp.g_upt[l], st));
        // conv1 over the bridged input
        SQ_TRY(relu_bwd(tr, i1, sp.g_upt[l], tape.upt[l], px * f, 1.0f, st));
        const unsigned eg = (unsigned)((px * f + 255) / 256);
        if (u->bridge == SQ_BRIDGE_CONCAT) {
            SQ_TRY(param_grads(tr, i1, tape.up[l], tape.down[l], sp.g_upt[l], px, D, H, W, sp.sc, st));
            SQ_TRY(conv_dgrad(tr, i1, sp.g_upt[l], px, D, H, W, sp.g_cat, st));
            slice_kernel<<<eg, 256, 0, st>>>(sp.g_cat, 2 * f, 0, px, f, sp.g_up[l]);
            slice_kernel<<<eg, 256, 0, st>>>(sp.g_cat, 2 * f, f, px, f, sp.g_down[l]);
            SQ_CHECK_LAUNCH();
        } else if (u->bridge == SQ_BRIDGE_NONE) {
            SQ_TRY(param_grads(tr, i1, tape.up[l], nullptr, sp.g_upt[l], px, D, H, W, sp.sc, st));
            SQ_TRY(conv_dgrad(tr, i1, sp.g_upt[l], px, D, H, W, sp.g_up[l], st));
            SQ_CUDA(cudaMemsetAsync(sp.g_down[l], 0, (size_t)px * f * sizeof(float), st));
        } else {
            SQ_TRY(param_grads(tr, i1, tape.merged[l], nullptr, sp.g_upt[l], px, D, H, W, sp.sc, st));
            SQ_TRY(conv_dgrad(tr, i1, sp.g_upt[l], px, D, H, W, sp.g_merged[l], st));
            eltwise_bwd_kernel<<<eg, 256, 0, st>>>(sp.g_merged[l], tape.up[l], tape.down[l], px * f, u->bridge,
                                                   sp.g_up[l], sp.g_down[l]);
            SQ_CHECK_LAUNCH();
        }
        // transposed conv: its input lives one level down
        const float *below = (l + 1 == nl - 1) ? tape.down[nl - 1] : tape.upo[l + 1];
        float *g_below = (l + 1 == nl - 1) ? sp.g_down[nl - 1] : sp.g_upo[l + 1];
        int Dc, Hc, Wc;
        dims(l + 1, &Dc, &Hc, &Wc);
        const long long pxc = level_px(u, g, l + 1);
        const SqLayer &LU = u->layers[iu];
        SQ_TRY(param_grads(tr, iu, below, nullptr, sp.g_up[l], pxc, Dc, Hc, Wc, sp.sc, st));
        const long long nin = pxc * LU.cin0;
        upconv_dgrad_kernel<<<(unsigned)((nin + 255) / 256), 256, 0, st>>>(sp.g_up[l], nin, Dc, Hc, Wc, LU.cin0, pd,
                                                                         LU.w, LU.cout, g_below);
        SQ_CHECK_LAUNCH();
    }

    // ---- down blocks, bottom to top
    for (int l = nl - 1; l >= 0; --l) {
        const long long px = level_px(u, g, l);
        const int f = u->filters[l];
        dims(l, &D, &H, &W);
        const int i2 = idx("UNet/down%d/conv2", l), i1 = idx("UNet/down%d/conv1", l);
        SQ_TRY(relu_bwd(tr, i2, sp.g_down[l], tape.down[l], px * f, keep, st));
        SQ_TRY(param_grads(tr, i2, tape.tmp[l], nullptr, sp.g_down[l], px, D, H, W, sp.sc, st));
        SQ_TRY(conv_dgrad(tr, i2, sp.g_down[l], px, D, H, W, sp.g_tmp[l], st));
        SQ_TRY(relu_bwd(tr, i1, sp.g_tmp[l], tape.tmp[l], px * f, 1.0f, st));
        const float *xin = (l == 0) ? image_dev : tape.pooled[l];
        SQ_TRY(param_grads(tr, i1, xin, nullptr, sp.g_tmp[l], px, D, H, W, sp.sc, st));
        if (l > 0) {
            SQ_TRY(conv_dgrad(tr, i1, sp.g_tmp[l], px, D, H, W, sp.g_pooled[l], st));
            const long long nout = px * u->filters[l - 1];
            int Df, Hf, Wf;
            dims(l - 1, &Df, &Hf, &Wf);
            // g_down[l-1] already holds the bridge's share (written by the up block of level l-1)
            maxpool_bwd_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(
                tape.down[l - 1], sp.g_pooled[l], nout, Df, Hf, Wf, u->filters[l - 1], pd, 1, sp.g_down[l - 1]);
            SQ_CHECK_LAUNCH();
        }
    }

    // ---- optimiser
    if (apply_update) SQ_TRY(sq_trainer_apply(tr, stream_));
    return SQ_OK;
}

// the optimiser update from the gradients that are in the arena now (a data-parallel caller runs the step with
// apply_update = 0, all-reduces the arena, then calls this)
extern "C" int sq_trainer_apply(sq_trainer_t tr, void *stream_)
{
    SQ_REQUIRE(tr, SQ_EINVAL, "trainer_apply: null pointer");
    sq_unet_s *u = tr->u;
    cudaStream_t st = (cudaStream_t)stream_;
    SQ_CUDA(cudaSetDevice(u->h->device));
    ++tr->step;
    const double t = (double)tr->step;
    const float lr_t = (float)((double)tr->lr * std::sqrt(1.0 - std::pow((double)tr->beta2, t)) /
                               (1.0 - std::pow((double)tr->beta1, t)));
    for (size_t i = 0; i < u->layers.size(); ++i) {
        SqLayer &L = u->layers[i];
        sq_trainer_s::Slot &s = tr->slots[i];
        update_kernel<<<(unsigned)((s.wcount + 255) / 256), 256, 0, st>>>(L.w, s.gw, s.mw, s.vw, s.wcount,
                                                                          tr->optimizer, tr->lr, lr_t, tr->beta1,
                                                                          tr->beta2, tr->eps);
        update_kernel<<<(unsigned)((L.cout + 255) / 256), 256, 0, st>>>(s.affine ? s.bias : L.shift, s.gb, s.mb, s.vb,
                                                                        (size_t)L.cout, tr->optimizer, tr->lr, lr_t,
                                                                        tr->beta1, tr->beta2, tr->eps);
        if (s.affine)
            refold_kernel<<<(unsigned)((L.cout + 255) / 256), 256, 0, st>>>(s.bias, L.scale, s.tsh, L.cout, L.shift);
    }
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

extern "C" int sq_trainer_grad_arena(sq_trainer_t tr, float **arena_dev, size_t *count)
{
    SQ_REQUIRE(tr && arena_dev && count, SQ_EINVAL, "trainer_grad_arena: null pointer");
    *arena_dev = tr->garena;
    *count = tr->gcount;
    return SQ_OK;
}

// value (what = 0) or last gradient (what = 1) of "<scope>/kernel" or "<scope>/bias", device -> host
extern "C" int sq_trainer_read(sq_trainer_t tr, const char *name, int what, float *out_host, size_t count)
{
    SQ_REQUIRE(tr && name && out_host, SQ_EINVAL, "trainer_read: null pointer");
    sq_unet_s *u = tr->u;
    const std::string full(name);
    const size_t slash = full.rfind('/');
    SQ_REQUIRE(slash != std::string::npos, SQ_EINVAL, "trainer_read: bad variable name '%s'", name);
    const int li = layer_index(u, full.substr(0, slash));
    const std::string var = full.substr(slash + 1);
    SQ_REQUIRE(li >= 0 && (var == "kernel" || var == "bias"), SQ_EINVAL, "trainer_read: unknown variable '%s'", name);
    const SqLayer &L = u->layers[li];
    const sq_trainer_s::Slot &s = tr->slots[li];
    const bool k = var == "kernel";
    const size_t have = k ? s.wcount : (size_t)L.cout;
    SQ_REQUIRE(count == have, SQ_EINVAL, "trainer_read: '%s' holds %zu values, not %zu", name, have, count);
    const float *src = what ? (k ? s.gw : s.gb) : (k ? L.w : (s.affine ? s.bias : L.shift));
    SQ_CUDA(cudaSetDevice(u->h->device));
    SQ_CUDA(cudaDeviceSynchronize());
    SQ_CUDA(cudaMemcpy(out_host, src, count * sizeof(float), cudaMemcpyDeviceToHost));
    return SQ_OK;
}
