// UNet plan management and the fp32 "exact" execution mode.
//
// Mirrors the reference graph builder: UNet.__init__ (networks/unet.py:126-146),
// build (:224-262), conv_block (:265-277), down_layer (:282-296), up_layer
// (:299-322), bridge (:182-202).  The four primitives the reference leaves
// abstract (:326-342) are defined in include/sequitr_b200.h.
//
// SQ_MODE_FP32_EXACT runs on CUDA cores with ONE fp32 accumulator per output and a
// fixed fmaf order (tap-major, then input channel; first input then skip input),
// which makes logits and masks bit-identical to oracle/unet_ref.c.  It is the
// verification mode; the throughput mode is SQ_MODE_BF16_TC (unet_tc.cu).
#include "unet_plan.cuh"
#include "conv_fp32_tile.cuh"
#include <algorithm>
#include <cmath>

// =============================================================== fp32 kernels
namespace {

// One thread = one output pixel x COT consecutive output channels.
// w: (taps, C0+C1, CO) [HWIO / DHWIO flattened], in0/in1 channels-last.
template <int COT>
__global__ void conv_fp32_kernel(const float *__restrict__ in0, int C0,
                                 const float *__restrict__ in1, int C1,
                                 long long npix, int D, int H, int W,
                                 const float *__restrict__ w, int KD, int KH, int KW, int CO,
                                 const float *__restrict__ scale, const float *__restrict__ shift,
                                 int relu, float *__restrict__ out)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const int co0 = blockIdx.y * COT;
    const int x = (int)(p % W);
    const int y = (int)((p / W) % H);
    const int z = (int)((p / ((long long)W * H)) % D);
    const long long nb = p / ((long long)W * H * D) * ((long long)W * H * D);   // frame base pixel
    const int C = C0 + C1;
    float acc[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) acc[j] = 0.0f;

    for (int kz = 0; kz < KD; ++kz) {
        const int zz = z + kz - KD / 2;
        if (zz < 0 || zz >= D) continue;
        for (int ky = 0; ky < KH; ++ky) {
            const int yy = y + ky - KH / 2;
            if (yy < 0 || yy >= H) continue;
            for (int kx = 0; kx < KW; ++kx) {
                const int xx = x + kx - KW / 2;
                if (xx < 0 || xx >= W) continue;
                const long long pix = nb + ((long long)zz * H + yy) * W + xx;
                const float *wk = w + (size_t)((kz * KH + ky) * KW + kx) * C * CO + co0;
                const float *p0 = in0 + pix * C0;
                for (int ci = 0; ci < C0; ++ci) {
                    const float v = p0[ci];
                    const float *wr = wk + (size_t)ci * CO;
#pragma unroll
                    for (int j = 0; j < COT; ++j)
                        if (co0 + j < CO) acc[j] = fmaf(v, __ldg(wr + j), acc[j]);
                }
                if (C1 > 0) {
                    const float *p1 = in1 + pix * C1;
                    for (int ci = 0; ci < C1; ++ci) {
                        const float v = p1[ci];
                        const float *wr = wk + (size_t)(C0 + ci) * CO;
#pragma unroll
                        for (int j = 0; j < COT; ++j)
                            if (co0 + j < CO) acc[j] = fmaf(v, __ldg(wr + j), acc[j]);
                    }
                }
            }
        }
    }
    float *o = out + p * CO + co0;
#pragma unroll
    for (int j = 0; j < COT; ++j)
        if (co0 + j < CO) {
            float v = fmaf(acc[j], scale[co0 + j], shift[co0 + j]);
            if (relu && !(v > 0.0f)) v = 0.0f;
            o[j] = v;
        }
}

// 2x2(x2) max pool, channels-last.  One thread per output element.
__global__ void maxpool_fp32_kernel(const float *__restrict__ in, long long nout, int D, int H,
                                    int W, int C, int pool_d, float *__restrict__ out)
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
    for (int dz = 0; dz < pool_d; ++dz)
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx)
                m = fmaxf(m, in[((((long long)n * D + z * pool_d + dz) * H + 2 * y + dy) * W +
                                 2 * x + dx) * C + c]);
    out[i] = m;
}

// 2x2(x2) stride-2 transposed conv + bias.  One thread per output element.
// w: (taps, CO, CI) with tap = (kz*2 + ky)*2 + kx (kz absent in 2-D).
__global__ void upconv_fp32_kernel(const float *__restrict__ in, long long nout, int D, int H,
                                   int W, int CI, int up_d, const float *__restrict__ w,
                                   const float *__restrict__ bias, int CO, float *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nout) return;
    const int Do = D * up_d, Ho = 2 * H, Wo = 2 * W;
    const int co = (int)(i % CO);
    long long r = i / CO;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho); r /= Ho;
    const int oz = (int)(r % Do);
    const long long n = r / Do;
    const int kz = (up_d == 2) ? (oz & 1) : 0;
    const int tap = (kz * 2 + (oy & 1)) * 2 + (ox & 1);
    const float *p = in + ((((long long)n * D + oz / up_d) * H + oy / 2) * W + ox / 2) * CI;
    const float *wr = w + ((size_t)tap * CO + co) * CI;
    float acc = 0.0f;
    for (int ci = 0; ci < CI; ++ci) acc = fmaf(p[ci], __ldg(wr + ci), acc);
    out[i] = acc + bias[co];
}

// The same transposed conv with one thread per INPUT pixel: all TAPS fine pixels x COT channels in registers, the
// weight address is uniform across a warp (the per-output kernel above reads weights CI floats apart between
// neighbouring threads).  Every output is still the chain fmaf(in[ci], w[tap][co][ci], acc) over ci, plus the bias.
template <int TAPS, int COT>
__global__ void upconv_fp32_px_kernel(const float *__restrict__ in, long long npix, int D, int H, int W, int CI,
                                      const float *__restrict__ w, const float *__restrict__ bias, int CO,
                                      float *__restrict__ out)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const int co0 = blockIdx.y * COT;
    float acc[TAPS][COT];
#pragma unroll
    for (int tp = 0; tp < TAPS; ++tp)
#pragma unroll
        for (int j = 0; j < COT; ++j) acc[tp][j] = 0.0f;
    const float *q = in + p * CI;
    for (int ci = 0; ci < CI; ++ci) {
        const float v = q[ci];
#pragma unroll
        for (int tp = 0; tp < TAPS; ++tp)
#pragma unroll
            for (int j = 0; j < COT; ++j)
                if (co0 + j < CO) acc[tp][j] = fmaf(v, __ldg(w + ((size_t)tp * CO + co0 + j) * CI + ci), acc[tp][j]);
    }
    const int x = (int)(p % W);
    const int y = (int)((p / W) % H);
    const int z = (int)((p / ((long long)W * H)) % D);
    const long long n = p / ((long long)W * H * D);
    constexpr int UD = TAPS / 4;
#pragma unroll
    for (int tp = 0; tp < TAPS; ++tp) {
        const int kx = tp & 1, ky = (tp >> 1) & 1, kz = tp >> 2;
        float *o = out + ((((long long)n * D * UD + z * UD + kz) * (2 * H) + 2 * y + ky) * (2 * W) + 2 * x + kx) * CO;
#pragma unroll
        for (int j = 0; j < COT; ++j)
            if (co0 + j < CO) o[co0 + j] = acc[tp][j] + bias[co0 + j];
    }
}

__global__ void eltwise_fp32_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                    long long n, int op, float *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = a[i], y = b[i];
    out[i] = op == SQ_BRIDGE_ADD ? __fadd_rn(x, y) : (op == SQ_BRIDGE_MUL ? __fmul_rn(x, y)
                                                                           : __fsub_rn(x, y));
}

// per-pixel softmax + first-max argmax over K <= 16 logits
__global__ void softmax_argmax_kernel(const float *__restrict__ logits, long long npix, int K,
                                      float *__restrict__ probs, uint8_t *__restrict__ mask)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float *l = logits + i * K;
    float v[16];
    int best = 0;
    float m = l[0];
    v[0] = m;
    for (int k = 1; k < K; ++k) {
        v[k] = l[k];
        if (v[k] > m) { m = v[k]; best = k; }
    }
    if (mask) mask[i] = (uint8_t)best;
    if (probs) {
        float s = 0.0f;
        for (int k = 0; k < K; ++k) { v[k] = expf(v[k] - m); s += v[k]; }
        for (int k = 0; k < K; ++k) probs[i * K + k] = __fdiv_rn(v[k], s);
    }
}

// tf.layers.dropout (reference networks/unet.py:274-276) with a counter-based generator, so that the oracle can
// restate the mask: u = top 24 bits of splitmix64(splitmix64(seed ^ (block << 48)) + index) / 2^24, the element is
// kept when u >= rate and divided by (1 - rate).
__device__ __forceinline__ unsigned long long sq_mix64(unsigned long long z)
{
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

__global__ void dropout_fp32_kernel(float *__restrict__ x, long long n, float rate, unsigned long long seed,
                                    int block_id)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long base = sq_mix64(seed ^ ((unsigned long long)block_id << 48));
    const float u = (float)(sq_mix64(base + (unsigned long long)i) >> 40) * (1.0f / 16777216.0f);
    x[i] = (u >= rate) ? __fdiv_rn(x[i], 1.0f - rate) : 0.0f;
}

}  // namespace

// ================================================================== plumbing
void sq_timer_mark(sq_unet_s *u, cudaStream_t st, const char *name, double flops)
{
    if (!u->timer.enabled) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    u->timer.ev.push_back(e);
    if (name) { u->timer.names.push_back(name); u->timer.flops.push_back(flops); }
}

namespace {

struct Geo { int n, d, h, w; };

long long level_pixels(const sq_unet_s *u, const Geo &g, int l)
{
    const int dl = (u->ndim == 3) ? (g.d >> l) : 1;
    return (long long)g.n * dl * (g.h >> l) * (g.w >> l);
}

int check_geometry(const sq_unet_s *u, int n, int d, int hgt, int wid)
{
    SQ_REQUIRE(u && u->finalized, SQ_ESTATE, "unet: plan not finalised");
    SQ_REQUIRE(n >= 1 && d >= 1 && hgt >= 1 && wid >= 1, SQ_EINVAL, "unet: bad shape");
    SQ_REQUIRE(u->ndim == 3 || d == 1, SQ_EINVAL, "unet: 2-D plan needs d == 1");
    const int m = 1 << (u->nlev - 1);
    // utils.divisible_by_two_n_times (reference utils.py:234-240)
    SQ_REQUIRE(hgt % m == 0 && wid % m == 0 && (u->ndim == 2 || d % m == 0), SQ_EINVAL,
               "unet: spatial dims (%d,%d,%d) must be divisible by %d for %d pooling levels", d,
               hgt, wid, m, u->nlev - 1);
    return SQ_OK;
}

const SqLayer *find_layer(const sq_unet_s *u, const std::string &scope)
{
    for (const SqLayer &l : u->layers)
        if (l.scope == scope) return &l;
    return nullptr;
}

int launch_conv(sq_unet_s *u, const SqLayer &L, const float *in0, const float *in1, long long npix,
                int D, int H, int W, int relu, float *out, cudaStream_t st)
{
    const int KD = (u->ndim == 3) ? L.ksize : 1;
    const int threads = 128;
    const unsigned gx = (unsigned)((npix + threads - 1) / threads);
    if (sqtile::can_tile(L.cin0 + L.cin1, L.cout) && !getenv("SQ_FP32_NOTILE")) {
        // same fmaf chain per output, staged through shared memory (conv_fp32_tile.cuh)
        SQ_CUDA(sqtile::launch(in0, L.cin0, in1, L.cin1, npix, D, H, W, L.w, KD, L.ksize, L.ksize, L.cout, L.scale,
                               L.shift, relu, out, st));
    } else if (L.cout % 16 == 0) {
        conv_fp32_kernel<16><<<dim3(gx, L.cout / 16), threads, 0, st>>>(
            in0, L.cin0, in1, L.cin1, npix, D, H, W, L.w, KD, L.ksize, L.ksize, L.cout, L.scale,
            L.shift, relu, out);
    } else {
        conv_fp32_kernel<4><<<dim3(gx, (L.cout + 3) / 4), threads, 0, st>>>(
            in0, L.cin0, in1, L.cin1, npix, D, H, W, L.w, KD, L.ksize, L.ksize, L.cout, L.scale,
            L.shift, relu, out);
    }
    ++u->last_launches;
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

// Shared by the workspace query (dry = true: no launches, null arena) and the real run.
int fp32_run(sq_unet_s *u, bool dry, const float *in, int n, int d, int hgt, int wid, float *probs,
             uint8_t *mask, float *logits, void *ws, size_t ws_bytes, cudaStream_t st,
             size_t *need, SqTape *tape = nullptr)
{
    const Geo g{n, d, hgt, wid};
    SqArena a(dry ? nullptr : ws, dry ? 0 : ws_bytes);
    const int nl = u->nlev;
    std::vector<float *> down(nl), tmp(nl), pooled(nl, nullptr);
    std::vector<float *> up(nl, nullptr), merged(nl, nullptr), upt(nl, nullptr), upo(nl, nullptr);
    for (int l = 0; l < nl; ++l) {
        const size_t px = (size_t)level_pixels(u, g, l);
        const size_t f = (size_t)u->filters[l];
        tmp[l] = a.take<float>(px * f);
        down[l] = a.take<float>(px * f);
        if (l > 0) pooled[l] = a.take<float>(px * (size_t)u->filters[l - 1]);
        if (l < nl - 1) {
            up[l] = a.take<float>(px * f);
            if (u->bridge >= SQ_BRIDGE_ADD && u->bridge <= SQ_BRIDGE_SUB)
                merged[l] = a.take<float>(px * f);
            upt[l] = a.take<float>(px * f);
            upo[l] = a.take<float>(px * f);
        }
    }
    float *logit_buf = logits;
    if (!logit_buf) logit_buf = a.take<float>((size_t)level_pixels(u, g, 0) * u->nout);
    if (need) *need = a.off;
    if (dry) return SQ_OK;
    SQ_REQUIRE(a.ok(), SQ_ENOMEM, "unet: workspace %zu < %zu bytes", ws_bytes, a.off);
    if (tape) {
        tape->down = down; tape->tmp = tmp; tape->pooled = pooled; tape->up = up; tape->merged = merged;
        tape->upt = upt; tape->upo = upo; tape->logits = logit_buf;
    }
    auto dropout = [&](float *x, long long count, int block_id) -> int {
        if (!tape || !(tape->drop_rate > 0.0f)) return SQ_OK;
        dropout_fp32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(x, count, tape->drop_rate, tape->seed,
                                                                            block_id);
        ++u->last_launches;
        SQ_CHECK_LAUNCH();
        return SQ_OK;
    };

    u->last_launches = 0;
    sq_timer_mark(u, st, nullptr, 0);
    const int threads = 256;
    const float *x = in;
    for (int l = 0; l < nl; ++l) {
        const int D = (u->ndim == 3) ? (d >> l) : 1, H = hgt >> l, W = wid >> l;
        const long long px = level_pixels(u, g, l);
        char scope[64];
        if (l > 0) {
            const long long nout = px * u->filters[l - 1];
            maxpool_fp32_kernel<<<(unsigned)((nout + threads - 1) / threads), threads, 0, st>>>(
                down[l - 1], nout, (u->ndim == 3) ? (d >> (l - 1)) : 1, hgt >> (l - 1),
                wid >> (l - 1), u->filters[l - 1], (u->ndim == 3) ? 2 : 1, pooled[l]);
            ++u->last_launches;
            SQ_CHECK_LAUNCH();
            sq_timer_mark(u, st, "maxpool", 0);
            x = pooled[l];
        }
        snprintf(scope, sizeof scope, "UNet/down%d/conv1", l);
        const SqLayer *c1 = find_layer(u, scope);
        snprintf(scope, sizeof scope, "UNet/down%d/conv2", l);
        const SqLayer *c2 = find_layer(u, scope);
        SQ_TRY(launch_conv(u, *c1, x, nullptr, px, D, H, W, 1, tmp[l], st));
        sq_timer_mark(u, st, c1->scope.c_str(), c1->flops_per_px * px);
        SQ_TRY(launch_conv(u, *c2, tmp[l], nullptr, px, D, H, W, 1, down[l], st));
        sq_timer_mark(u, st, c2->scope.c_str(), c2->flops_per_px * px);
        SQ_TRY(dropout(down[l], px * u->filters[l], l));
    }
    const float *cur = down[nl - 1];
    for (int l = nl - 2; l >= 0; --l) {
        const int D = (u->ndim == 3) ? (d >> l) : 1, H = hgt >> l, W = wid >> l;
        const long long px = level_pixels(u, g, l);
        char scope[64];
        snprintf(scope, sizeof scope, "UNet/up%d/upscale", l);
        const SqLayer *us = find_layer(u, scope);
        snprintf(scope, sizeof scope, "UNet/up%d/conv1", l);
        const SqLayer *c1 = find_layer(u, scope);
        snprintf(scope, sizeof scope, "UNet/up%d/conv2", l);
        const SqLayer *c2 = find_layer(u, scope);
        const long long nout = px * us->cout;
        {
            const long long pin = level_pixels(u, g, l + 1);
            const int Dc = (u->ndim == 3) ? (d >> (l + 1)) : 1, Hc = hgt >> (l + 1), Wc = wid >> (l + 1);
            const unsigned gxp = (unsigned)((pin + 127) / 128);
            if (u->ndim == 3)
                upconv_fp32_px_kernel<8, 4><<<dim3(gxp, (us->cout + 3) / 4), 128, 0, st>>>(
                    cur, pin, Dc, Hc, Wc, us->cin0, us->w, us->shift, us->cout, up[l]);
            else
                upconv_fp32_px_kernel<4, 8><<<dim3(gxp, (us->cout + 7) / 8), 128, 0, st>>>(
                    cur, pin, Dc, Hc, Wc, us->cin0, us->w, us->shift, us->cout, up[l]);
        }
        ++u->last_launches;
        SQ_CHECK_LAUNCH();
        sq_timer_mark(u, st, us->scope.c_str(), us->flops_per_px * px);
        const float *in0 = up[l], *in1 = nullptr;
        if (u->bridge == SQ_BRIDGE_CONCAT) {
            in1 = down[l];
        } else if (u->bridge != SQ_BRIDGE_NONE) {
            eltwise_fp32_kernel<<<(unsigned)((nout + threads - 1) / threads), threads, 0, st>>>(
                up[l], down[l], nout, u->bridge, merged[l]);
            ++u->last_launches;
            SQ_CHECK_LAUNCH();
            sq_timer_mark(u, st, "bridge", 0);
            in0 = merged[l];
        }
        SQ_TRY(launch_conv(u, *c1, in0, in1, px, D, H, W, 1, upt[l], st));
        sq_timer_mark(u, st, c1->scope.c_str(), c1->flops_per_px * px);
        SQ_TRY(launch_conv(u, *c2, upt[l], nullptr, px, D, H, W, 1, upo[l], st));
        sq_timer_mark(u, st, c2->scope.c_str(), c2->flops_per_px * px);
        SQ_TRY(dropout(upo[l], px * u->filters[l], nl + l));
        cur = upo[l];
    }
    const SqLayer *head = find_layer(u, "UNet/to_image");
    const long long px0 = level_pixels(u, g, 0);
    SQ_TRY(launch_conv(u, *head, cur, nullptr, px0, (u->ndim == 3) ? d : 1, hgt, wid, 0, logit_buf, st));
    if (probs || mask) {
        softmax_argmax_kernel<<<(unsigned)((px0 + threads - 1) / threads), threads, 0, st>>>(
            logit_buf, px0, u->nout, probs, mask);
        ++u->last_launches;
        SQ_CHECK_LAUNCH();
    }
    sq_timer_mark(u, st, head->scope.c_str(), head->flops_per_px * px0);
    return SQ_OK;
}

int upload(sq_unet_s *u, const float *src, size_t count, float **dst)
{
    SQ_CUDA(cudaMalloc((void **)dst, std::max<size_t>(count, 1) * sizeof(float)));
    u->dev_allocs.push_back(*dst);
    SQ_CUDA(cudaMemcpy(*dst, src, count * sizeof(float), cudaMemcpyHostToDevice));
    return SQ_OK;
}

}  // namespace

int sq_fp32_workspace(sq_unet_s *u, int n, int d, int hgt, int wid, size_t *need)
{
    SQ_TRY(check_geometry(u, n, d, hgt, wid));
    return fp32_run(u, true, nullptr, n, d, hgt, wid, nullptr, nullptr, nullptr, nullptr, 0, nullptr, need);
}

int sq_fp32_forward_tape(sq_unet_s *u, const float *in, int n, int d, int hgt, int wid, void *ws, size_t ws_bytes,
                         cudaStream_t st, SqTape *tape)
{
    SQ_TRY(check_geometry(u, n, d, hgt, wid));
    return fp32_run(u, false, in, n, d, hgt, wid, nullptr, nullptr, nullptr, ws, ws_bytes, st, nullptr, tape);
}

// ================================================================== C ABI
extern "C" int sq_unet_create(sq_handle_t h, int ndim, int num_inputs, int num_outputs,
                              const int *filters, int nlev, int bridge, int mode, sq_unet_t *out)
{
    SQ_REQUIRE(h && filters && out, SQ_EINVAL, "unet_create: null pointer");
    SQ_REQUIRE(ndim == 2 || ndim == 3, SQ_EINVAL, "unet_create: ndim must be 2 or 3");
    SQ_REQUIRE(nlev >= 1 && nlev <= 8, SQ_EINVAL, "unet_create: 1..8 levels supported");
    SQ_REQUIRE(num_inputs >= 1 && num_inputs <= 64, SQ_EINVAL, "unet_create: bad num_inputs");
    SQ_REQUIRE(num_outputs >= 1 && num_outputs <= 16, SQ_EINVAL, "unet_create: 1..16 outputs");
    // reference networks/unet.py:186-187
    SQ_REQUIRE(bridge >= SQ_BRIDGE_NONE && bridge <= SQ_BRIDGE_CONCAT, SQ_EINVAL,
               "Bridge type not recognized");
    SQ_REQUIRE(mode == SQ_MODE_FP32_EXACT || mode == SQ_MODE_BF16_TC, SQ_EINVAL,
               "unet_create: unknown mode");
    for (int i = 0; i < nlev; ++i)
        SQ_REQUIRE(filters[i] >= 1 && filters[i] <= 1024, SQ_EINVAL, "unet_create: bad filter count");
    sq_unet_s *u = new sq_unet_s();
    u->h = h;
    u->ndim = ndim; u->cin = num_inputs; u->nout = num_outputs; u->nlev = nlev;
    u->bridge = bridge; u->mode = mode;
    u->filters.assign(filters, filters + nlev);
    const double taps = (ndim == 3) ? 27.0 : 9.0;
    auto add = [&](SqLayer::Kind k, const std::string &scope, int level, int c0, int c1, int co,
                   int ks, double fl) {
        SqLayer L;
        L.kind = k; L.scope = scope; L.level = level; L.cin0 = c0; L.cin1 = c1; L.cout = co;
        L.ksize = ks; L.flops_per_px = fl;
        u->layers.push_back(L);
    };
    int cin = num_inputs;
    for (int i = 0; i < nlev; ++i) {
        const int f = filters[i];
        add(SqLayer::CONV, "UNet/down" + std::to_string(i) + "/conv1", i, cin, 0, f, 3, 2 * taps * cin * f);
        add(SqLayer::CONV, "UNet/down" + std::to_string(i) + "/conv2", i, f, 0, f, 3, 2 * taps * f * f);
        cin = f;
    }
    for (int i = nlev - 2; i >= 0; --i) {
        const int f = filters[i];
        add(SqLayer::UPCONV, "UNet/up" + std::to_string(i) + "/upscale", i, cin, 0, f, 2, 2.0 * cin * f);
        const int c1 = (bridge == SQ_BRIDGE_CONCAT) ? f : 0;
        add(SqLayer::CONV, "UNet/up" + std::to_string(i) + "/conv1", i, f, c1, f, 3, 2 * taps * (f + c1) * f);
        add(SqLayer::CONV, "UNet/up" + std::to_string(i) + "/conv2", i, f, 0, f, 3, 2 * taps * f * f);
        cin = f;
    }
    add(SqLayer::HEAD, "UNet/to_image", 0, cin, 0, num_outputs, 1, 2.0 * cin * num_outputs);
    *out = u;
    return SQ_OK;
}

extern "C" int sq_unet_destroy(sq_unet_t u)
{
    if (!u) return SQ_OK;
    sq_tc_destroy(u);
    for (void *p : u->dev_allocs) cudaFree(p);
    for (cudaEvent_t e : u->timer.ev) cudaEventDestroy(e);
    delete u;
    return SQ_OK;
}

extern "C" int sq_unet_load_weights(sq_unet_t u, const char *name, const float *data,
                                    const int64_t *shape, int rank)
{
    SQ_REQUIRE(u && name && data && shape, SQ_EINVAL, "unet_load_weights: null pointer");
    SQ_REQUIRE(!u->finalized, SQ_ESTATE, "unet_load_weights: plan already finalised");
    SQ_REQUIRE(rank >= 1 && rank <= 5, SQ_EINVAL, "unet_load_weights: bad rank");
    const std::string full(name);
    const size_t slash = full.rfind('/');
    SQ_REQUIRE(slash != std::string::npos, SQ_EINVAL, "unet_load_weights: bad name '%s'", name);
    const std::string scope = full.substr(0, slash), var = full.substr(slash + 1);
    const SqLayer *L = find_layer(u, scope);
    SQ_REQUIRE(L, SQ_EINVAL, "unet_load_weights: unknown scope '%s'", scope.c_str());
    std::vector<int64_t> want;
    const int nd = u->ndim;
    if (var == "kernel") {
        for (int i = 0; i < nd; ++i) want.push_back(L->kind == SqLayer::CONV ? 3 : (L->kind == SqLayer::UPCONV ? 2 : 1));
        if (L->kind == SqLayer::UPCONV) { want.push_back(L->cout); want.push_back(L->cin0); }
        else { want.push_back(L->cin0 + L->cin1); want.push_back(L->cout); }
    } else if (var == "bias" || var == "scale" || var == "shift") {
        SQ_REQUIRE(var == "bias" || L->kind == SqLayer::CONV, SQ_EINVAL,
                   "unet_load_weights: '%s' only valid for conv layers", var.c_str());
        want.push_back(L->cout);
    } else {
        SQ_REQUIRE(false, SQ_EINVAL, "unet_load_weights: unknown variable '%s'", var.c_str());
    }
    bool ok = (int)want.size() == rank;
    size_t count = 1;
    for (int i = 0; ok && i < rank; ++i) { ok = want[i] == shape[i]; count *= (size_t)shape[i]; }
    if (!ok) {
        std::string ws, gs;
        for (auto v : want) ws += std::to_string(v) + ",";
        for (int i = 0; i < rank; ++i) gs += std::to_string(shape[i]) + ",";
        SQ_REQUIRE(false, SQ_EINVAL, "unet_load_weights: '%s' has shape (%s) expected (%s)", name,
                   gs.c_str(), ws.c_str());
    }
    SqHostTensor t;
    t.data.assign(data, data + count);
    t.shape.assign(shape, shape + rank);
    u->host[full] = std::move(t);
    return SQ_OK;
}

extern "C" int sq_unet_finalize(sq_unet_t u)
{
    SQ_REQUIRE(u, SQ_EINVAL, "unet_finalize: null plan");
    SQ_REQUIRE(!u->finalized, SQ_ESTATE, "unet_finalize: already finalised");
    SQ_CUDA(cudaSetDevice(u->h->device));
    for (SqLayer &L : u->layers) {
        auto k = u->host.find(L.scope + "/kernel");
        auto b = u->host.find(L.scope + "/bias");
        SQ_REQUIRE(k != u->host.end() && b != u->host.end(), SQ_ESTATE,
                   "unet_finalize: missing kernel/bias for '%s'", L.scope.c_str());
        auto s = u->host.find(L.scope + "/scale");
        auto t = u->host.find(L.scope + "/shift");
        SQ_REQUIRE((s == u->host.end()) == (t == u->host.end()), SQ_ESTATE,
                   "unet_finalize: '%s' needs both scale and shift or neither", L.scope.c_str());
        std::vector<float> sc(L.cout, 1.0f), sh(b->second.data);
        if (s != u->host.end())
            for (int c = 0; c < L.cout; ++c) {
                sc[c] = s->second.data[c];
                // y = (conv + bias)*scale + shift  ==  conv*scale + (bias*scale + shift)
                volatile float prod = b->second.data[c] * sc[c];
                sh[c] = prod + t->second.data[c];
            }
        SQ_TRY(upload(u, k->second.data.data(), k->second.data.size(), &L.w));
        SQ_TRY(upload(u, sc.data(), sc.size(), &L.scale));
        SQ_TRY(upload(u, sh.data(), sh.size(), &L.shift));
        // keep the folded epilogue on the host for the tensor-core re-layout
        SqHostTensor fs, ft;
        fs.data = sc; fs.shape = {L.cout};
        ft.data = sh; ft.shape = {L.cout};
        u->host[L.scope + "/_scale"] = fs;
        u->host[L.scope + "/_shift"] = ft;
    }
    if (u->mode == SQ_MODE_BF16_TC) SQ_TRY(sq_tc_finalize(u));
    u->finalized = true;
    return SQ_OK;
}

extern "C" int sq_unet_workspace_bytes(sq_unet_t u, int n, int d, int hgt, int wid, size_t *bytes)
{
    SQ_REQUIRE(bytes, SQ_EINVAL, "unet_workspace_bytes: null pointer");
    SQ_TRY(check_geometry(u, n, d, hgt, wid));
    if (u->mode == SQ_MODE_BF16_TC) return sq_tc_workspace_bytes(u, n, d, hgt, wid, bytes);
    return fp32_run(u, true, nullptr, n, d, hgt, wid, nullptr, nullptr, nullptr, nullptr, 0,
                    nullptr, bytes);
}

extern "C" int sq_unet_forward(sq_unet_t u, const float *in, int n, int d, int hgt, int wid,
                               float *probs, uint8_t *mask, float *logits, void *ws,
                               size_t ws_bytes, void *stream_)
{
    SQ_TRY(check_geometry(u, n, d, hgt, wid));
    SQ_REQUIRE(in && ws, SQ_EINVAL, "unet_forward: null pointer");
    cudaStream_t st = (cudaStream_t)stream_;
    if (u->mode == SQ_MODE_BF16_TC)
        return sq_tc_forward(u, in, n, d, hgt, wid, probs, mask, logits, ws, ws_bytes, st);
    return fp32_run(u, false, in, n, d, hgt, wid, probs, mask, logits, ws, ws_bytes, st, nullptr);
}

extern "C" int sq_unet_last_launches(sq_unet_t u, int *launches)
{
    SQ_REQUIRE(u && launches, SQ_EINVAL, "unet_last_launches: null pointer");
    *launches = u->last_launches;
    return SQ_OK;
}

extern "C" int sq_unet_profile(sq_unet_t u, const float *in, int n, int d, int hgt, int wid,
                               void *ws, size_t ws_bytes, void *stream_, const char **names,
                               float *ms, double *flops, int max_layers, int *n_layers)
{
    SQ_REQUIRE(u && names && ms && flops && n_layers, SQ_EINVAL, "unet_profile: null pointer");
    for (cudaEvent_t e : u->timer.ev) cudaEventDestroy(e);
    u->timer.ev.clear(); u->timer.names.clear(); u->timer.flops.clear();
    u->timer.enabled = true;
    const int s = sq_unet_forward(u, in, n, d, hgt, wid, nullptr, nullptr, nullptr, ws, ws_bytes, stream_);
    u->timer.enabled = false;
    SQ_TRY(s);
    SQ_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
    const int nl = (int)u->timer.names.size();
    *n_layers = nl;
    for (int i = 0; i < nl && i < max_layers; ++i) {
        names[i] = u->timer.names[i];
        flops[i] = u->timer.flops[i];
        SQ_CUDA(cudaEventElapsedTime(&ms[i], u->timer.ev[i], u->timer.ev[i + 1]));
    }
    return SQ_OK;
}

static int segment_localise_impl(sq_unet_t u, const void *frames_host, int in_dtype, int normalise,
                                 int n, int hgt, int wid, int frame0, float *table_host,
                                 int32_t *counts_host, int max_rows, uint8_t *mask_host);

extern "C" int sq_segment_localise_raw_host(sq_unet_t u, const void *frames_host, int in_dtype, int normalise,
                                            int n, int hgt, int wid, int frame0, float *table_host,
                                            int32_t *counts_host, int max_rows, uint8_t *mask_host)
{
    SQ_REQUIRE(u && u->h, SQ_EINVAL, "segment_localise_host: null plan");
    SqHostCall call(u->h);          // one host call per handle at a time; streams drained on every exit path
    return segment_localise_impl(u, frames_host, in_dtype, normalise, n, hgt, wid, frame0, table_host,
                                 counts_host, max_rows, mask_host);
}

static int segment_localise_impl(sq_unet_t u, const void *frames_host, int in_dtype, int normalise,
                                 int n, int hgt, int wid, int frame0, float *table_host,
                                 int32_t *counts_host, int max_rows, uint8_t *mask_host)
{
    SQ_TRY(check_geometry(u, n, 1, hgt, wid));
    SQ_REQUIRE(u->ndim == 2, SQ_EUNSUPPORTED, "segment_localise_host: planar stacks only");
    SQ_REQUIRE(frames_host && table_host && counts_host, SQ_EINVAL, "segment_localise_host: null pointer");
    SQ_REQUIRE(in_dtype == SQ_F32 || in_dtype == SQ_U8 || in_dtype == SQ_U16, SQ_EINVAL,
               "segment_localise_host: in_dtype must be SQ_F32, SQ_U8 or SQ_U16");
    sq_handle_s *h = u->h;
    SQ_CUDA(cudaSetDevice(h->device));
    // Frames stream through in chunks: the H2D copy of chunk c+1 (copy stream, double-buffered
    // input) overlaps the UNet of chunk c (compute stream); label-and-localise then runs once
    // over the whole batch of masks and only the small centroid tables travel back.
    const size_t px1 = (size_t)hgt * wid, px = (size_t)n * px1;
    const size_t esz = in_dtype == SQ_U8 ? 1 : (in_dtype == SQ_U16 ? 2 : 4);
    const bool fuse_norm = normalise && in_dtype == SQ_U16 && u->mode == SQ_MODE_BF16_TC && sq_tc_can_take_raw_u16(u, hgt, wid);
    const bool staged = !fuse_norm && (in_dtype != SQ_F32 || normalise);      // raw chunk -> float32 chunk on the device
    // Chunk schedule 1, 1, 2, 4, 8, 8, ...: the first copy (the only one nothing can hide) is one frame;
    // steady-state chunks are as large as the bench batch (the net runs ~5 % faster on 8 frames than on 4).
    // Chunk size is bounded by memory: 2 input buffers + the net's workspace per chunk.
    int ch = n >= 24 ? 8 : (n >= 8 ? 4 : (n >= 2 ? n / 2 : 1));
    while (ch > 1 && (size_t)ch * hgt * wid > ((size_t)1 << 25)) ch >>= 1;      // <= 32 Mpx per chunk
    std::vector<int> chunk_of;
    for (int done = 0; done < n;) {
        int c = ch;
        if (n >= 8) {
            const size_t k = chunk_of.size();
            c = k < 2 ? 1 : (k == 2 ? 2 : (k == 3 ? std::min(ch, 4) : ch));
        }
        c = std::min(c, n - done);
        chunk_of.push_back(c);
        done += c;
    }
    const int nchunks = (int)chunk_of.size();
    size_t unet_ws = 0, lab_ws = 0, prep_ws = 0;
    SQ_TRY(sq_unet_workspace_bytes(u, ch, 1, hgt, wid, &unet_ws));
    SQ_TRY(sq_label_workspace_bytes(h, n, 1, hgt, wid, max_rows, &lab_ws));
    SQ_TRY(sq_prep_workspace_bytes(h, ch, u->cin, &prep_ws));
    const size_t chunk_elems = (size_t)ch * px1 * u->cin;
    SqArena probe(nullptr, 0);
    probe.take<char>(2 * chunk_elems * esz);
    if (staged || fuse_norm) probe.take<float>(chunk_elems);
    probe.take<uint8_t>(px);
    probe.take<float>((size_t)n * max_rows * 5);
    probe.take<int32_t>(n);
    SQ_TRY(sq_reserve_device(h, probe.off + sq_align_up(unet_ws) + sq_align_up(lab_ws) + sq_align_up(prep_ws) + 1024));
    SqArena a(h->dev_arena, h->dev_arena_bytes);
    char *frames = a.take<char>(2 * chunk_elems * esz);
    float *stage = (staged || fuse_norm) ? a.take<float>(chunk_elems) : nullptr;
    uint8_t *mask = a.take<uint8_t>(px);
    float *table = a.take<float>((size_t)n * max_rows * 5);
    int32_t *counts = a.take<int32_t>(n);
    void *w1 = a.take<char>(unet_ws);
    void *w2 = a.take<char>(lab_ws);
    void *w3 = a.take<char>(prep_ws);
    cudaStream_t st = h->stream, cs = h->copy_stream;
    for (int c = 0, f0 = 0; c < nchunks; f0 += chunk_of[c], ++c) {
        const int b = c & 1;
        const int nc = chunk_of[c];
        const size_t elems = (size_t)nc * px1 * u->cin;
        char *buf = frames + (size_t)b * chunk_elems * esz;
        if (c >= 2) SQ_CUDA(cudaStreamWaitEvent(cs, h->ev_done[b], 0));     // buffer b is free again
        SQ_CUDA(cudaMemcpyAsync(buf, (const char *)frames_host + (size_t)f0 * px1 * u->cin * esz, elems * esz,
                                cudaMemcpyHostToDevice, cs));
        SQ_CUDA(cudaEventRecord(h->ev_h2d[b], cs));
        SQ_CUDA(cudaStreamWaitEvent(st, h->ev_h2d[b], 0));
        const float *net_in = (const float *)buf;
        if (fuse_norm) {
            // uint16 frames + ImageNorm: widened, normalised and rounded to bf16 ONCE (2 B/px, into the stage buffer); the
            // fused first pair of the UNet loads that as is -- no float32 copy of the chunk.  SQ_QNORM=2: only the
            // moments are computed here and the first pair's loader normalises the raw values itself (slower).
            const char *qn = getenv("SQ_QNORM");
            if (qn && atoi(qn) == 2) {
                const float2 *stats = nullptr;
                SQ_TRY(sq_image_norm_stats_u16(h, (const uint16_t *)buf, nc, hgt, wid, w3, prep_ws, st, &stats));
                SQ_TRY(sq_tc_forward_raw_u16(u, (const uint16_t *)buf, stats, nc, hgt, wid, mask + (size_t)f0 * px1, w1, unet_ws, st));
            } else {
                SQ_TRY(sq_image_norm_u16_to_bf16(h, (const uint16_t *)buf, stage, nc, hgt, wid, w3, prep_ws, st));
                SQ_TRY(sq_tc_forward_raw_u16(u, (const uint16_t *)stage, nullptr, nc, hgt, wid, mask + (size_t)f0 * px1, w1, unet_ws, st));
            }
            SQ_CUDA(cudaEventRecord(h->ev_done[b], st));
            continue;
        }
        if (staged) {
            // widen (exact for 8/16-bit integers) and optionally normalise; the raw buffer is free as
            // soon as the cast has run, the float32 stage is consumed in stream order
            if (normalise) SQ_TRY(sq_image_norm_raw(h, buf, in_dtype, stage, nc, hgt, wid, u->cin, w3, prep_ws, st));
            else SQ_TRY(sq_image_cast(h, buf, in_dtype, stage, (long long)elems, st));
            net_in = stage;
        }
        SQ_TRY(sq_unet_forward(u, net_in, nc, 1, hgt, wid, nullptr, mask + (size_t)f0 * px1, nullptr, w1,
                               unet_ws, st));
        SQ_CUDA(cudaEventRecord(h->ev_done[b], st));
    }
    SQ_TRY(sq_label_centroids(h, mask, n, 1, hgt, wid, frame0, nullptr, table, counts, max_rows, w2,
                              lab_ws, st));
    SQ_CUDA(cudaMemcpyAsync(counts_host, counts, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SQ_CUDA(cudaMemcpyAsync(table_host, table, (size_t)n * max_rows * 5 * sizeof(float),
                            cudaMemcpyDeviceToHost, st));
    if (mask_host) SQ_CUDA(cudaMemcpyAsync(mask_host, mask, px, cudaMemcpyDeviceToHost, st));
    SQ_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n; ++i)
        SQ_REQUIRE(counts_host[i] <= max_rows, SQ_EOVERFLOW,
                   "segment_localise: frame %d has %d components > max_rows=%d", frame0 + i,
                   counts_host[i], max_rows);
    return SQ_OK;
}

extern "C" int sq_segment_localise_host(sq_unet_t u, const float *frames_host, int n, int hgt,
                                        int wid, int frame0, float *table_host,
                                        int32_t *counts_host, int max_rows, uint8_t *mask_host)
{
    return sq_segment_localise_raw_host(u, frames_host, SQ_F32, 0, n, hgt, wid, frame0, table_host, counts_host,
                                        max_rows, mask_host);
}
