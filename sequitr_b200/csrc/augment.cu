// tr_augment (reference networks/unet.py:348-401) as ONE gather kernel: rotate (image bilinear,
// label nearest, weights bilinear + 1 outside the rotated frame) -> crop -> one-hot labels.
// The reference builds four full-size rotated tensors with tf.contrib.image.rotate and crops them
// afterwards; here only the crop is ever computed.  The projective sampling follows TensorFlow 1.x
// contrib/image ImageProjectiveTransform (third-party, absent; restated in oracle/augment_oracle.py):
// float32 arithmetic without fused multiply-add, std::round for NEAREST, zero fill per bilinear corner.
// HBM-bound: (4C + 5) B/px gathered in, (4C + K + 4) B/px out per cropped pixel.
#include "sq_common.cuh"

namespace {

constexpr int AUG_MAX_BATCH = 32;   // frames per launch (parameters travel by value)

struct AugFrame {
    float a0, a1, a2, b0, b1, b2;   // input_x = a0*x + a1*y + a2; input_y = b0*x + b1*y + b2
    int rh, rw;                     // crop origin in the rotated frame
};
struct AugParams {
    AugFrame f[AUG_MAX_BATCH];
};

__device__ __forceinline__ float aug_coord(float c0, float c1, float c2, float x, float y)
{
    // (c0*x + c1*y) + c2 with every product and sum rounded on its own (TensorFlow's CPU kernel)
    return __fadd_rn(__fadd_rn(__fmul_rn(c0, x), __fmul_rn(c1, y)), c2);
}

// the four corner taps of a zero-filled bilinear sample: offsets (or -1 outside) and the float32 weights
struct AugTaps {
    long long o00, o01, o10, o11;
    float wx0, wx1, wy0, wy1;
};

__device__ __forceinline__ AugTaps aug_taps(int hgt, int wid, float y, float x)
{
    AugTaps T;
    const float yf = floorf(y), xf = floorf(x);
    const float yc = __fadd_rn(yf, 1.0f), xc = __fadd_rn(xf, 1.0f);
    // clamp before the int conversion: coordinates far outside the frame read the fill value anyway
    const int y0 = (int)fminf(fmaxf(yf, -2.0f), (float)hgt + 1.0f), x0 = (int)fminf(fmaxf(xf, -2.0f), (float)wid + 1.0f);
    const bool r0 = y0 >= 0 && y0 < hgt, r1 = y0 + 1 >= 0 && y0 + 1 < hgt;
    const bool c0 = x0 >= 0 && x0 < wid, c1 = x0 + 1 >= 0 && x0 + 1 < wid;
    const long long base = (long long)y0 * wid + x0;
    T.o00 = (r0 && c0) ? base : -1;
    T.o01 = (r0 && c1) ? base + 1 : -1;
    T.o10 = (r1 && c0) ? base + wid : -1;
    T.o11 = (r1 && c1) ? base + wid + 1 : -1;
    T.wx1 = __fsub_rn(xc, x); T.wx0 = __fsub_rn(x, xf);
    T.wy1 = __fsub_rn(yc, y); T.wy0 = __fsub_rn(y, yf);
    return T;
}

__device__ __forceinline__ float aug_tap(const float *__restrict__ img, long long o, int c, int ch)
{
    return o >= 0 ? __ldg(img + o * c + ch) : 0.0f;
}

// TensorFlow's order: (x_ceil-x)*f(x_floor) + (x-x_floor)*f(x_ceil) per row, then the same over rows
__device__ __forceinline__ float aug_bilinear(const float *__restrict__ img, const AugTaps &T, int c, int ch)
{
    const float v_floor = __fadd_rn(__fmul_rn(T.wx1, aug_tap(img, T.o00, c, ch)), __fmul_rn(T.wx0, aug_tap(img, T.o01, c, ch)));
    const float v_ceil = __fadd_rn(__fmul_rn(T.wx1, aug_tap(img, T.o10, c, ch)), __fmul_rn(T.wx0, aug_tap(img, T.o11, c, ch)));
    return __fadd_rn(__fmul_rn(T.wy1, v_floor), __fmul_rn(T.wy0, v_ceil));
}

// A block owns a 32x32 tile of the crop (one warp per row, four passes): the rotated footprint of a
// square tile touches ~1.5x its own sectors, a flat 32x8 strip more than 2x.
constexpr int AUG_TILE = 32;

__global__ void __launch_bounds__(256)
augment_kernel(const float *__restrict__ image, const uint8_t *__restrict__ label, const float *__restrict__ weights,
               int hgt, int wid, int c, int ch, int cw, int k, AugParams P, float *__restrict__ image_out,
               uint8_t *__restrict__ label_out, float *__restrict__ weights_out)
{
    const int ox = blockIdx.x * AUG_TILE + (threadIdx.x & 31);
    const int n = blockIdx.z;
    if (ox >= cw) return;
    const AugFrame F = P.f[n];
    const long long frame = (long long)hgt * wid;
    const float *img = image + (long long)n * frame * c;
    const float *wgt = weights + (long long)n * frame;
    const uint8_t *lab = label + (long long)n * frame;
    const float x = (float)(ox + F.rw);
    for (int oy = blockIdx.y * AUG_TILE + (threadIdx.x >> 5); oy < min(ch, (int)(blockIdx.y + 1) * AUG_TILE); oy += 8) {
        const float y = (float)(oy + F.rh);
        const float ix = aug_coord(F.a0, F.a1, F.a2, x, y);
        const float iy = aug_coord(F.b0, F.b1, F.b2, x, y);
        const AugTaps T = aug_taps(hgt, wid, iy, ix);
        const long long o = ((long long)n * ch + oy) * cw + ox;
        for (int q = 0; q < c; ++q) image_out[o * c + q] = aug_bilinear(img, T, c, q);
        // NEAREST: std::round = half away from zero
        const float ry = roundf(iy), rx = roundf(ix);
        const bool inside = ry >= 0.0f && ry < (float)hgt && rx >= 0.0f && rx < (float)wid;
        const int lv = inside ? lab[(long long)ry * wid + (long long)rx] : 0;
        for (int q = 0; q < k; ++q) label_out[o * k + q] = (uint8_t)(lv == q);
        weights_out[o] = __fadd_rn(aug_bilinear(wgt, T, 1, 0), inside ? 0.0f : 1.0f);
    }
}

}  // namespace

extern "C" int sq_tr_augment(sq_handle_t h, const float *image, const uint8_t *label, const float *weights, int n,
                             int hgt, int wid, int c, const float *transforms_host, const int *crop_host, int ch,
                             int cw, int num_outputs, float *image_out, uint8_t *label_out, float *weights_out,
                             void *stream_)
{
    SQ_REQUIRE(h && image && label && weights && transforms_host && crop_host && image_out && label_out && weights_out,
               SQ_EINVAL, "tr_augment: null pointer");
    SQ_REQUIRE(n >= 1 && hgt >= 1 && wid >= 1 && c >= 1 && ch >= 1 && cw >= 1, SQ_EINVAL,
               "tr_augment: sizes must be positive");
    SQ_REQUIRE(num_outputs >= 1 && num_outputs <= 5, SQ_EINVAL,
               "tr_augment: 1 <= num_outputs <= 5 (the reference expands at most five label channels)");
    for (int i = 0; i < n; ++i)
        SQ_REQUIRE(crop_host[2 * i] >= 0 && crop_host[2 * i + 1] >= 0 && crop_host[2 * i] + ch <= hgt &&
                       crop_host[2 * i + 1] + cw <= wid,
                   SQ_EINVAL, "tr_augment: crop window of frame %d leaves the image", i);
    cudaStream_t st = (cudaStream_t)stream_;
    const long long in_frame = (long long)hgt * wid, out_frame = (long long)ch * cw;
    for (int n0 = 0; n0 < n; n0 += AUG_MAX_BATCH) {
        const int nb = n - n0 < AUG_MAX_BATCH ? n - n0 : AUG_MAX_BATCH;
        AugParams P;
        for (int i = 0; i < nb; ++i) {
            const float *t = transforms_host + 6 * (size_t)(n0 + i);
            P.f[i] = AugFrame{t[0], t[1], t[2], t[3], t[4], t[5], crop_host[2 * (n0 + i)], crop_host[2 * (n0 + i) + 1]};
        }
        for (int i = nb; i < AUG_MAX_BATCH; ++i) P.f[i] = AugFrame{1, 0, 0, 0, 1, 0, 0, 0};
        dim3 grid((unsigned)sq_div_up(cw, AUG_TILE), (unsigned)sq_div_up(ch, AUG_TILE), (unsigned)nb);
        augment_kernel<<<grid, 256, 0, st>>>(image + n0 * in_frame * c, label + n0 * in_frame, weights + n0 * in_frame,
                                             hgt, wid, c, ch, cw, num_outputs, P, image_out + n0 * out_frame * c,
                                             label_out + n0 * out_frame * num_outputs, weights_out + n0 * out_frame);
    }
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}
