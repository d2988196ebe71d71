/* CPU oracle for the UNet layer arithmetic -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the four layer primitives the reference leaves
 * abstract (/root/reference/sequitr/networks/unet.py:326-342) and of the skip
 * "bridge" (unet.py:182-202), with TensorFlow semantics: NHWC / NDHWC
 * activations, HWIO / DHWIO kernels, SAME zero padding, ReLU (unet.py:142).
 * The topology (unet.py:224-322) is driven from Python (oracle/unet_c.py).
 *
 * "Parity unpinned": the reference ships neither the concrete layers nor any
 * golden vector; the arithmetic lives in TensorFlow 1.x (not installable
 * here).  This file is pinned to torch-CPU conv semantics by
 * tests/test_oracle_unet.py.
 *
 * NUMERIC CONTRACT (what the GPU "fp32 exact" mode reproduces bit for bit):
 * every output value owns ONE fp32 accumulator, started at +0 and updated with
 * fmaf() in the fixed order  (kz,) ky, kx, ci  -- ci running over the first
 * input then the second (folded concat: [upsampled, skip], unet.py:197) --
 * skipping out-of-image taps; the epilogue is y = fmaf(acc, scale[co],
 * shift[co]) then ReLU if requested.  Parallelism is only ACROSS outputs.
 *
 * Threading: every function works on the flattened outer-row range [r0, r1)
 * (row = n*H + y in 2-D, (n*D + z)*H + y in 3-D; element ranges for the flat
 * ops) so that oracle/unet_c.py can fan rows out over Python threads (ctypes
 * drops the GIL; libgomp is not installed in this image).  Outputs are
 * disjoint per row, so results do not depend on the thread count.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define MAXC 1024

/* one output pixel of a 2-D/3-D conv: accumulate tap (wk = kernel slice for
 * this tap, pix = flattened input pixel index) into acc[0..CO) */
static inline void tap_accumulate(const float *in0, int C0, const float *in1, int C1,
                                  size_t pix, const float *wk, int CO, float *acc)
{
    const float *p0 = in0 + pix * C0;
    for (int ci = 0; ci < C0; ++ci) {
        const float v = p0[ci];
        const float *wr = wk + (size_t)ci * CO;
        for (int co = 0; co < CO; ++co) acc[co] = fmaf(v, wr[co], acc[co]);
    }
    if (C1 > 0) {
        const float *p1 = in1 + pix * C1;
        for (int ci = 0; ci < C1; ++ci) {
            const float v = p1[ci];
            const float *wr = wk + (size_t)(C0 + ci) * CO;
            for (int co = 0; co < CO; ++co) acc[co] = fmaf(v, wr[co], acc[co]);
        }
    }
}

static inline void epilogue(const float *acc, const float *scale, const float *shift,
                            int relu, int CO, float *o)
{
    for (int co = 0; co < CO; ++co) {
        float v = fmaf(acc[co], scale[co], shift[co]);
        if (relu && !(v > 0.0f)) v = 0.0f;
        o[co] = v;
    }
}

/* 2-D KHxKW SAME convolution over the channel-concatenation of in0|in1.
 * in0: (N,H,W,C0)  in1: (N,H,W,C1) or NULL  w: (KH,KW,C0+C1,CO)  out: (N,H,W,CO) */
void sqref_conv2d(const float *in0, int C0, const float *in1, int C1,
                  int H, int W, const float *w, int KH, int KW, int CO,
                  const float *scale, const float *shift, int relu, float *out,
                  long r0, long r1)
{
    const int C = C0 + C1;
    const int ph = KH / 2, pw = KW / 2;
    float acc[MAXC];
    for (long r = r0; r < r1; ++r) {
        const long n = r / H;
        const int y = (int)(r % H);
        for (int x = 0; x < W; ++x) {
            for (int co = 0; co < CO; ++co) acc[co] = 0.0f;
            for (int ky = 0; ky < KH; ++ky) {
                const int yy = y + ky - ph;
                if (yy < 0 || yy >= H) continue;
                for (int kx = 0; kx < KW; ++kx) {
                    const int xx = x + kx - pw;
                    if (xx < 0 || xx >= W) continue;
                    tap_accumulate(in0, C0, in1, C1, ((size_t)n * H + yy) * W + xx,
                                   w + (size_t)(ky * KW + kx) * C * CO, CO, acc);
                }
            }
            epilogue(acc, scale, shift, relu, CO, out + (((size_t)n * H + y) * W + x) * CO);
        }
    }
}

/* 3-D KDxKHxKW SAME convolution, NDHWC / DHWIO, same contract. */
void sqref_conv3d(const float *in0, int C0, const float *in1, int C1,
                  int D, int H, int W, const float *w, int KD, int KH, int KW, int CO,
                  const float *scale, const float *shift, int relu, float *out,
                  long r0, long r1)
{
    const int C = C0 + C1;
    const int pd = KD / 2, ph = KH / 2, pw = KW / 2;
    float acc[MAXC];
    for (long r = r0; r < r1; ++r) {
        const int y = (int)(r % H);
        const int z = (int)((r / H) % D);
        const long n = r / ((long)H * D);
        for (int x = 0; x < W; ++x) {
            for (int co = 0; co < CO; ++co) acc[co] = 0.0f;
            for (int kz = 0; kz < KD; ++kz) {
                const int zz = z + kz - pd;
                if (zz < 0 || zz >= D) continue;
                for (int ky = 0; ky < KH; ++ky) {
                    const int yy = y + ky - ph;
                    if (yy < 0 || yy >= H) continue;
                    for (int kx = 0; kx < KW; ++kx) {
                        const int xx = x + kx - pw;
                        if (xx < 0 || xx >= W) continue;
                        tap_accumulate(in0, C0, in1, C1,
                                       (((size_t)n * D + zz) * H + yy) * W + xx,
                                       w + (size_t)((kz * KH + ky) * KW + kx) * C * CO, CO, acc);
                    }
                }
            }
            epilogue(acc, scale, shift, relu, CO,
                     out + ((((size_t)n * D + z) * H + y) * W + x) * CO);
        }
    }
}

/* 2x2 stride-2 max pool, NHWC; rows index the OUTPUT (n*Ho + y). */
void sqref_maxpool2d(const float *in, int H, int W, int C, float *out, long r0, long r1)
{
    const int Ho = H / 2, Wo = W / 2;
    for (long r = r0; r < r1; ++r) {
        const long n = r / Ho;
        const int y = (int)(r % Ho);
        for (int x = 0; x < Wo; ++x)
            for (int c = 0; c < C; ++c) {
                const float *p = in + (((size_t)n * H + 2 * y) * W + 2 * x) * C + c;
                float m = p[0];
                m = fmaxf(m, p[C]);
                m = fmaxf(m, p[(size_t)W * C]);
                m = fmaxf(m, p[(size_t)W * C + C]);
                out[(((size_t)n * Ho + y) * Wo + x) * C + c] = m;
            }
    }
}

/* 2x2x2 stride-2 max pool, NDHWC; rows index the OUTPUT ((n*Do + z)*Ho + y). */
void sqref_maxpool3d(const float *in, int D, int H, int W, int C, float *out, long r0, long r1)
{
    const int Do = D / 2, Ho = H / 2, Wo = W / 2;
    for (long r = r0; r < r1; ++r) {
        const int y = (int)(r % Ho);
        const int z = (int)((r / Ho) % Do);
        const long n = r / ((long)Ho * Do);
        for (int x = 0; x < Wo; ++x)
            for (int c = 0; c < C; ++c) {
                float m = -INFINITY;
                for (int dz = 0; dz < 2; ++dz)
                    for (int dy = 0; dy < 2; ++dy)
                        for (int dx = 0; dx < 2; ++dx)
                            m = fmaxf(m, in[((((size_t)n * D + 2 * z + dz) * H + 2 * y + dy) * W
                                             + 2 * x + dx) * C + c]);
                out[((((size_t)n * Do + z) * Ho + y) * Wo + x) * C + c] = m;
            }
    }
}

/* 2x2 stride-2 transposed convolution + bias (no activation).
 * in: (N,H,W,CI)  w: (2,2,CO,CI) [tf.layers.conv2d_transpose kernel layout]
 * out: (N,2H,2W,CO); out[2y+ky,2x+kx,co] = (fmaf chain over ci from +0) + bias.
 * rows index the INPUT (n*H + y). */
void sqref_upconv2d(const float *in, int H, int W, int CI,
                    const float *w, const float *bias, int CO, float *out, long r0, long r1)
{
    for (long r = r0; r < r1; ++r) {
        const long n = r / H;
        const int y = (int)(r % H);
        for (int x = 0; x < W; ++x) {
            const float *p = in + (((size_t)n * H + y) * W + x) * CI;
            for (int ky = 0; ky < 2; ++ky)
                for (int kx = 0; kx < 2; ++kx) {
                    float *o = out + (((size_t)n * 2 * H + 2 * y + ky) * 2 * W + 2 * x + kx) * CO;
                    for (int co = 0; co < CO; ++co) {
                        const float *wr = w + ((size_t)(ky * 2 + kx) * CO + co) * CI;
                        float acc = 0.0f;
                        for (int ci = 0; ci < CI; ++ci) acc = fmaf(p[ci], wr[ci], acc);
                        o[co] = acc + bias[co];
                    }
                }
        }
    }
}

/* 2x2x2 stride-2 transposed convolution + bias.  w: (2,2,2,CO,CI).
 * rows index the INPUT ((n*D + z)*H + y). */
void sqref_upconv3d(const float *in, int D, int H, int W, int CI,
                    const float *w, const float *bias, int CO, float *out, long r0, long r1)
{
    for (long r = r0; r < r1; ++r) {
        const int y = (int)(r % H);
        const int z = (int)((r / H) % D);
        const long n = r / ((long)H * D);
        for (int x = 0; x < W; ++x) {
            const float *p = in + ((((size_t)n * D + z) * H + y) * W + x) * CI;
            for (int k = 0; k < 8; ++k) {
                const int kz = k >> 2, ky = (k >> 1) & 1, kx = k & 1;
                float *o = out + ((((size_t)n * 2 * D + 2 * z + kz) * 2 * H + 2 * y + ky) * 2 * W
                                  + 2 * x + kx) * CO;
                for (int co = 0; co < CO; ++co) {
                    const float *wr = w + ((size_t)k * CO + co) * CI;
                    float acc = 0.0f;
                    for (int ci = 0; ci < CI; ++ci) acc = fmaf(p[ci], wr[ci], acc);
                    o[co] = acc + bias[co];
                }
            }
        }
    }
}

/* Element-wise bridges (unet.py:190-195): op 0 add, 1 mul, 2 sub; out = a op b. */
void sqref_eltwise(const float *a, const float *b, int op, float *out, long r0, long r1)
{
    for (long i = r0; i < r1; ++i)
        out[i] = op == 0 ? a[i] + b[i] : (op == 1 ? a[i] * b[i] : a[i] - b[i]);
}

/* Head: per-pixel softmax over K logits + first-max argmax (np.argmax rule). */
void sqref_softmax_argmax(const float *logits, int K, float *probs, uint8_t *mask,
                          long r0, long r1)
{
    for (long i = r0; i < r1; ++i) {
        const float *l = logits + (size_t)i * K;
        int best = 0;
        float m = l[0];
        for (int k = 1; k < K; ++k) if (l[k] > m) { m = l[k]; best = k; }
        float s = 0.0f, e[16];
        for (int k = 0; k < K; ++k) { e[k] = expf(l[k] - m); s += e[k]; }
        if (probs) for (int k = 0; k < K; ++k) probs[(size_t)i * K + k] = e[k] / s;
        if (mask) mask[i] = (uint8_t)best;
    }
}

/* Round-to-nearest-even fp32 -> bf16 -> fp32 (the tensor-core path's storage rounding). */
void sqref_round_bf16(const float *in, float *out, long r0, long r1)
{
    for (long i = r0; i < r1; ++i) {
        uint32_t u;
        memcpy(&u, &in[i], 4);
        if ((u & 0x7fffffffu) > 0x7f800000u) { u |= 0x00400000u; u &= 0xffff0000u; }
        else { u += 0x7fffu + ((u >> 16) & 1u); u &= 0xffff0000u; }
        memcpy(&out[i], &u, 4);
    }
}
