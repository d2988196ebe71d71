// SQ_MODE_BF16_TC: the UNet on 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Data layout in HBM: activations are bf16, channel-blocked "NC/8HW8" (volumes: N D C/8 H W 8):
//     act[n][z][c/8][y][x][c%8]
// so that (a) a pixel's 8-channel slice is one 16-byte core-matrix row, (b) a TMA
// box (8*PW, PH, 2 blocks, 1 slice, 1 image) lands in shared memory as [2][PH][PW][8] --
// already a canonical K-major, no-swizzle UMMA operand whose rows are pixels -- and
// (c) shifting the descriptor start address by (ky*PW + kx)*16 bytes selects the
// 3x3 tap (ky,kx).  Each halo patch is therefore read from L2/HBM ONCE per 16 input
// channels and multiplied 9 times from shared memory; SAME zero padding is TMA's
// out-of-bounds fill (also across slices of a volume); the skip-concat is just a second
// tensor map (k-steps ks0..ks0+ks1 read the skip tensor), no concatenated tensor ever exists.
//
// Two persistent, warp-specialised kernel templates do every dense layer:
//   warp 0      TMA producer   (patch via cp.async.bulk.tensor.5d, weights via cp.async.bulk)
//   warp 1      MMA issuer     (tcgen05.mma kind::f16, M=128 pixels, K=16 per instruction)
//   warps 2-9   epilogue       (tcgen05.ld -> scale/shift/ReLU -> bf16 -> 16-byte stores;
//                               two groups of four warps, one warp per TMEM lane quarter)
// with an smem ring (full/empty mbarriers) and double-buffered TMEM accumulators
// (tmem_full/tmem_empty) so the epilogue of tile i overlaps the MMAs of tile i+1.
//   conv_tc_kernel  conv 3x3(x3): tile = 8 x (16*S) pixels, N = Cout, 9 taps accumulate into one
//                   accumulator per sub-tile; up-conv 2x2 stride 2 = 4 independent 1x1 GEMMs (one per
//                   output sub-position), scattered 2x upsampled store
//   conv_xc_kernel  conv 3x3(x3) for Cout <= 32 with several k-steps per tile: the three horizontal
//                   taps folded into N = 3*Cout (3 MMAs per k-step), combined by lane shuffles in the
//                   epilogue; weights resident in shared memory
//   conv_qd_kernel  level 0 of planar nets with 16 filters on the QUAD (space-to-depth) layout: a 2x2 pixel block is
//                   one pixel of a half-resolution image with 64 channels, a 3x3 conv 16 -> 16 becomes four 128 x 64 x 16
//                   MMAs per input parity (the other weight blocks are zero), the up-conv a 1x1 conv, pool / head in-thread
//   conv_qf_kernel  down0/conv1 + down0/conv2 + pool in ONE launch (the first conv as an im2col GEMM, its output stays
//                   in shared memory as conv2's halo patch)
//   conv_qu_kernel  up0/upscale + up0/conv1 in ONE launch (the up-sampled patch stays in shared memory, the skip
//                   tensor streams through a TMA ring)
// Epilogue variants: plain store, store + fused 2x2 max-pool, fused 1x1 head + softmax + argmax.
// Run-time switches (A/B measurements and tests; defaults are the measured optimum): SQ_QUAD=0 keeps level 0 on the
// full-resolution kernels, SQ_QFUSE=0 / SQ_QUP=0 split the fused level-0 pairs, SQ_XC=0/2 disables / forces the
// x-combined kernel, SQ_CLUSTER=1 runs the Cout >= 128 convs as weight-multicasting CTA pairs, SQ_FUSE_FIRST=1 /
// SQ_PAIR=1 are the (slower) on-chip hand-offs of rounds 1-2 on the full-resolution layout.
// The first conv (Cin <= 4: K = 9..108 -- warp-level mma.sync on a staged halo tile), the depth
// half of the 2x2x2 pool, the element-wise bridges and the stand-alone head are bandwidth-bound
// CUDA-core kernels on the same layout.
//
// Numeric contract (oracle/unet_c.py contract='bf16'): inputs, weights and every
// stored activation are bf16 (RNE); accumulation and the scale/shift epilogue are fp32.
#include "unet_plan.cuh"
#include "tc_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------ tile configuration
template <int COUT, int S, bool UP>
struct Cfg {
    static constexpr int TH = 16 * S;                    // tile height (pixels)
    static constexpr int PW = UP ? 8 : 10;               // patch width incl. halo
    static constexpr int PH = UP ? TH : TH + 2;
    static constexpr int NT = UP ? 4 : 9;                // taps per k-step
    static constexpr int NQ = UP ? 4 : 1;                // accumulators per sub-tile
    static constexpr int A_BYTES = 2 * PH * PW * 16;     // 16 input channels of the patch
    static constexpr int B_BYTES = NT * 2 * COUT * 16;   // 16 input channels of the weights
    static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES + 127) / 128 * 128;
    static constexpr int ACC_COLS = S * NQ * COUT;
    // conv with an even number of sub-tiles: sub-tiles 2p and 2p+1 take the even / odd image rows
    // of a 32-row block, so vertically adjacent pixels sit in the SAME accumulator lane (same
    // thread) of two sub-tiles and the fused 2x2 max-pool needs no vertical shuffle
    static constexpr bool ILV = !UP && (S % 2 == 0);
    static constexpr uint32_t LBO_A = PH * PW * 16, SBO_A = (ILV ? 2 : 1) * PW * 16;
    static constexpr uint32_t LBO_B = COUT * 16, SBO_B = 128;
    static_assert(A_BYTES % 128 == 0, "TMA destination alignment");
};

constexpr int MAX_STAGES = 8;
constexpr int EPI_GROUPS = 2;                 // epilogue warp groups (4 warps each, one per TMEM lane quarter)
constexpr int TC_THREADS = 64 + 128 * EPI_GROUPS;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

__device__ __forceinline__ uint32_t bf162_max(uint32_t a, uint32_t b)
{
    __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162 *>(&a), y = *reinterpret_cast<__nv_bfloat162 *>(&b);
    __nv_bfloat162 m = __hmax2(x, y);
    return *reinterpret_cast<uint32_t *>(&m);
}

__device__ __forceinline__ void mma_m16n8k16_bf16(float *c, const uint32_t *a, uint32_t b0, uint32_t b1);

// Fused first conv (NBLD > 0): the fp32 single-channel input and the first conv's parameters; the A half of
// every smem stage is then BUILT by NBLD extra warps (conv1 of the halo patch) instead of fetched by TMA.
struct FirstArgs {
    const float *in;           // (n, H, W) fp32
    const float *wf;           // [9][16] fp32 holding bf16-rounded weights
    const float *scale, *shift;
};

// Epilogue variants of the 3x3 conv kernel
constexpr int EPI_STORE = 0;   // activation -> bf16 blocked tensor
constexpr int EPI_POOL = 1;    // same + the 2x2 max-pooled tensor (fused max_pool_layer)
constexpr int EPI_HEAD = 2;    // fused 1x1 head + softmax + argmax; the activation is never stored

struct HeadArgs {
    const float *w;            // [COUT][K] then bias[K] (bf16-rounded weights, fp32 bias)
    int K;
    float *logits, *probs;     // optional, (n,H,W,K) fp32
    uint8_t *mask;             // optional, (n,H,W)
};

// NMMA = MMA-issuing warps (warps 1..NMMA; sub-tile j belongs to warp 1 + j % NMMA): one thread sustains
// one MMA per ~45-65 clk, which a single-CTA-per-SM kernel with a 64 clk pipe time (Cout = 128) cannot hide.
// CL = launched as clusters of two CTAs that walk their (different) pixel tiles in lock-step and share
// every weight stage: each CTA fetches HALF of the stage's weights and multicasts it into both CTAs'
// shared memory, halving the L2 -> SM weight traffic that otherwise caps the Cout >= 128 layers (36.8 /
// 73.7 KB of weights per k-step and 128-256 pixel tile).  A stage is free when the MMAs of BOTH CTAs
// have retired (tcgen05.commit multicast onto both CTAs' `empty` barriers).
// NBLD > 0 = level-0 fusion: the layer's input is the FIRST conv's output (Cin = 1 -> 16 channels), which
// never goes to HBM: warps FIRST_BLD.. compute it for each 66 x 10 halo patch (mma.sync.m16n8k16 on the
// L1-cached fp32 frame, exactly the fragments and epilogue of first_conv_kernel, zeros outside the image =
// this layer's SAME padding) and write it into the stage in the layout TMA would have produced.
// The folded epilogue of a 3x3 conv with 32..128 output channels, passed BY VALUE next to the device arrays: the plain
// store epilogue then reads scale / shift as constant-bank operands of its FFMAs instead of 8 LDS.128 per work item
// (ncu: 26 % short-scoreboard stalls in these kernels).  Work items are split between the two epilogue groups by
// sub-tile, so the channel chunk -- and with it every table index -- is a compile-time constant.
struct TcEpi {
    float sc[128], sh[128];
};

template <int COUT, int S, bool UP, int NBUF, int MINB, int EPI, int HK, int NMMA, bool CL, int NBLD = 0>
__global__ void __launch_bounds__(TC_THREADS + 32 * (NMMA - 1) + 32 * NBLD, MINB)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               int ks0, int ks1, const bf16 *__restrict__ wts, const float *__restrict__ scale,
               const float *__restrict__ shift, bf16 *__restrict__ out, bf16 *__restrict__ out_pool,
               HeadArgs head, int nimg, int H, int W, int relu, int nstages, int D, int KZ,
               int out_mul, int out_off, FirstArgs first, const __grid_constant__ TcEpi epc)
{
    constexpr bool CONST_EPI = !UP && EPI == EPI_STORE && COUT >= 32 && COUT <= 128 && (S % EPI_GROUPS) == 0;
    static_assert(NBLD == 0 || (!UP && !CL), "fused first conv: plain 3x3 conv launches only");
    constexpr int FIRST_BLD = 1 + NMMA + 4 * EPI_GROUPS;       // first builder warp
    // Volumes: activations are [n][z][c/8][y][x][8]; a (z-slice, image) pair is one "image" np =
    // n*D + z for tiling and for the epilogue, and a 3x3x3 conv is the same 9-tap stage run for
    // KZ = 3 z-offsets: k-step q = kz*ksteps + ks loads the halo patch of slice z + kz - 1 (TMA
    // zero-fills slices outside the volume) and the weights of depth tap kz.  Planar: D = KZ = 1.
    using C = Cfg<COUT, S, UP>;
    constexpr int TMEM_COLS = (NBUF * C::ACC_COLS <= 32) ? 32 : (NBUF * C::ACC_COLS <= 64) ? 64
                            : (NBUF * C::ACC_COLS <= 128) ? 128 : (NBUF * C::ACC_COLS <= 256) ? 256 : 512;
    static_assert(NBUF * C::ACC_COLS <= 512, "TMEM overflow");
    extern __shared__ uint8_t smem_raw[];
    // the runtime only guarantees 16-byte alignment of dynamic shared memory: align by hand
    uint8_t *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_sh;
    __shared__ __align__(16) float s_scale[COUT], s_shift[COUT];
    __shared__ __align__(16) float s_head[EPI == EPI_HEAD ? COUT * HK + HK : 4];   // [k][c] then bias[k]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (W + 7) >> 3, tiles_y = (H + C::TH - 1) / C::TH;
    const int ntiles = nimg * D * tiles_x * tiles_y;
    const int ksteps = ks0 + ks1, qsteps = KZ * ksteps;

    if (threadIdx.x == 0) {
        for (int i = 0; i < nstages; ++i) { tc::mbar_init(&full_bar[i], 1 + NBLD); tc::mbar_init(&empty_bar[i], NMMA * (CL ? 2 : 1)); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull_bar[i], NMMA); tc::mbar_init(&tempty_bar[i], 4 * EPI_GROUPS); }
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&mapA0);
        tc::tma_prefetch_desc(&mapA1);
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_base_sh, TMEM_COLS); tc::tmem_relinquish(); }
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) { s_scale[i] = scale[i]; s_shift[i] = shift[i]; }
    if constexpr (EPI == EPI_HEAD) {
        // global layout is [c][k] (+ bias); keep it transposed so one class is 16 contiguous floats
        for (int i = threadIdx.x; i < COUT * HK; i += blockDim.x) s_head[(i % HK) * COUT + i / HK] = head.w[i];
        for (int i = threadIdx.x; i < HK; i += blockDim.x) s_head[COUT * HK + i] = head.w[COUT * HK + i];
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;
    // lock-step tile loop for clusters: both CTAs run the same number of iterations; an iteration whose
    // tile index is past the end loads zero-filled patches (image coordinate out of range) and stores nothing
    const int iters = CL ? (ntiles + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t crank = CL ? tc::cluster_ctarank() : 0u;
    if (CL) tc::cluster_sync_all();                         // peer barriers are initialised

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x, i = 0; CL ? (i < iters) : (t < ntiles); t += gridDim.x, ++i) {
                const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, np = t / (tiles_x * tiles_y);
                const int z = np % D, n = np / D;
                const int x0 = tx * 8 - (UP ? 0 : 1), y0 = ty * C::TH - (UP ? 0 : 1);
                for (int q = 0; q < qsteps; ++q) {
                    const int kz = q / ksteps, ks = q - kz * ksteps;
                    const int zc = z + kz - (KZ >> 1);
                    tc::mbar_wait(&empty_bar[stage], phase ^ 1);
                    tc::mbar_arrive_expect_tx(&full_bar[stage], (NBLD ? 0 : C::A_BYTES) + C::B_BYTES);
                    uint8_t *sA = smem + (size_t)stage * C::STAGE_BYTES;
                    if (NBLD) { /* the builder warps write the patch */ }
                    else if (ks < ks0) tc::tma_load_5d(sA, &mapA0, &full_bar[stage], x0 * 8, y0, ks * 2, zc, n);
                    else          tc::tma_load_5d(sA, &mapA1, &full_bar[stage], x0 * 8, y0, (ks - ks0) * 2, zc, n);
                    if (CL)      // my half of the weights, into both CTAs
                        tc::bulk_load_mc(sA + C::A_BYTES + crank * (C::B_BYTES / 2),
                                         wts + (size_t)q * (C::B_BYTES / 2) + crank * (C::B_BYTES / 4), C::B_BYTES / 2,
                                         &full_bar[stage], (uint16_t)3);
                    else
                        tc::bulk_load(sA + C::A_BYTES, wts + (size_t)q * (C::B_BYTES / 2), C::B_BYTES,
                                      &full_bar[stage]);
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp <= NMMA) {
        // ======================================================= MMA issuer(s)
        // The whole warp walks the loop converged (all lanes poll the barriers); one elected
        // lane issues.  Descriptors = per-stage base + compile-time offset (one add each).
        const uint32_t idesc = tc::instr_desc_bf16(128, COUT);
        const uint32_t a_hi = ((C::SBO_A >> 4) & 0x3FFFu) | (1u << 14);
        const uint32_t b_hi = ((C::SBO_B >> 4) & 0x3FFFu) | (1u << 14);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int t = blockIdx.x; CL ? (it < iters) : (t < ntiles); t += gridDim.x, ++it) {
            const int buf = it % NBUF;
            tc::mbar_wait(&tempty_bar[buf], ((it / NBUF) & 1) ^ 1);
            tc::tc_fence_after();
            for (int ks = 0; ks < qsteps; ++ks) {
                tc::mbar_wait(&full_bar[stage], phase);
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t a_base = tc::smem_u32(smem + (size_t)stage * C::STAGE_BYTES);
                    const uint32_t a_lo = ((a_base >> 4) & 0x3FFFu) | (((C::LBO_A >> 4) & 0x3FFFu) << 16);
                    const uint32_t b_lo = (((a_base + C::A_BYTES) >> 4) & 0x3FFFu) |
                                          (((C::LBO_B >> 4) & 0x3FFFu) << 16);
                    const uint32_t d0 = tmem_base + buf * C::ACC_COLS;
                    const uint32_t first = (ks > 0) ? 1u : 0u;
#pragma unroll
                    for (int j = 0; j < S; ++j) {
                        if (NMMA > 1 && (j % NMMA) != warp - 1) continue;
#pragma unroll
                        for (int tp = 0; tp < C::NT; ++tp) {
                            const int row0 = C::ILV ? 32 * (j / 2) + (j & 1) : j * 16;   // first image row
                            const uint32_t a_off = UP ? (uint32_t)(row0 * C::PW)
                                                      : (uint32_t)((row0 + tp / 3) * C::PW + tp % 3);
                            const int q = UP ? tp : 0;
                            tc::umma_bf16_parts(d0 + (j * C::NQ + q) * COUT, a_lo + a_off, a_hi,
                                                b_lo + tp * 2 * COUT, b_hi, idesc,
                                                (UP || tp == 0) ? first : 1u);
                        }
                    }
                    // smem slot reusable once these MMAs retire (in both CTAs of a cluster)
                    if (CL) tc::umma_commit_mc(&empty_bar[stage], (uint16_t)3);
                    else tc::umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
            if (tc::elect_one()) tc::umma_commit(&tfull_bar[buf]);   // accumulators of this tile complete
            __syncwarp();
        }
    } else if (NBLD > 0 && warp >= FIRST_BLD) {
        // ================================================== first-conv builders
        if constexpr (NBLD > 0) {
            const int bw = warp - FIRST_BLD, g = lane >> 2, t = lane & 3;
            uint32_t bwf[2][2];
            float sc1[2][2], sh1[2][2];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k0 = 2 * t + 8 * h;
                    const float w0 = k0 < 9 ? first.wf[k0 * 16 + nt * 8 + g] : 0.0f;
                    const float w1 = k0 + 1 < 9 ? first.wf[(k0 + 1) * 16 + nt * 8 + g] : 0.0f;
                    bwf[nt][h] = pack_bf16(w0, w1);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) { sc1[nt][e] = first.scale[nt * 8 + 2 * t + e]; sh1[nt][e] = first.shift[nt * 8 + 2 * t + e]; }
            }
            // this lane's taps: k = 2t, 2t+1 (always < 9) and, for t = 0 only, k = 8
            constexpr int RW = C::PW + 2, RH = 16 * ((C::PH + 15) / 16) + 2, NRAW = RW * RH;   // raw fp32 halo of the patch
            // (rows rounded up to whole 16-row m-tiles: the short last m-tile of a column reads them, stores nothing)
            constexpr int RPT = (NRAW + 32 * NBLD - 1) / (32 * NBLD);         // raw values per builder thread
            __shared__ float raw[2][NRAW];
            const int ta = 2 * t, tb = 2 * t + 1;
            const int oa = (ta / 3) * RW + ta % 3, ob = (tb / 3) * RW + tb % 3, oc = 2 * RW + 2;
            const int bt = threadIdx.x - FIRST_BLD * 32;                      // 0 .. 32 NBLD - 1
            // global -> registers for one tile's raw halo (zero outside the image): issued one tile ahead so
            // that the L2 latency hides behind the previous tile's conv1
            float nxt[RPT];
            auto fetch = [&](int tl) {
                const int tx = tl % tiles_x, ty = (tl / tiles_x) % tiles_y, np = tl / (tiles_x * tiles_y);
                const int xr = tx * 8 - 2, yr = ty * C::TH - 2;
                const float *img = first.in + (size_t)np * H * W;
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const int i = bt + k * 32 * NBLD;
                    const int rr = i / RW, cc = i - rr * RW;
                    const int Y = yr + rr, X = xr + cc;
                    nxt[k] = (i < NRAW && Y >= 0 && Y < H && X >= 0 && X < W) ? __ldg(img + (size_t)Y * W + X) : 0.0f;
                }
            };
            int stage = 0, buf = 0;
            uint32_t phase = 0;
            if ((int)blockIdx.x < ntiles) fetch(blockIdx.x);
            for (int tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
                const int tx = tl % tiles_x, ty = (tl / tiles_x) % tiles_y;
                const int x0 = tx * 8 - 1, y0 = ty * C::TH - 1;
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const int i = bt + k * 32 * NBLD;
                    if (i < NRAW) raw[buf][i] = nxt[k];
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * NBLD) : "memory");   // the builders' own barrier
                if (tl + (int)gridDim.x < ntiles) fetch(tl + gridDim.x);
                tc::mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t *sA = smem + (size_t)stage * C::STAGE_BYTES;
                // m-tiles run DOWN the patch columns (16 consecutive rows of one column, CT per column): no
                // division per pixel, and every lane's offsets are a per-lane constant plus a warp-uniform term
                constexpr int CT = (C::PH + 15) / 16, MTC = CT * C::PW;
                const float *rb = raw[buf] + g * RW;                              // row g of the raw halo
                uint8_t *sl = sA + g * (C::PW * 16) + 4 * t;
                const bool interior = y0 >= 0 && y0 + C::PH <= H && x0 >= 0 && x0 + C::PW <= W;
#pragma unroll 2
                for (int m = bw; m < MTC; m += NBLD) {
                    const int c = m / CT, r0 = (m - c * CT) * 16;                 // warp-uniform
                    const float *q0 = rb + r0 * RW + c, *q1 = q0 + 8 * RW;        // tap (0,0) of pixels (r0+g, c), (r0+g+8, c)
                    uint32_t a[4];
                    a[0] = pack_bf16(q0[oa], q0[ob]);
                    a[1] = pack_bf16(q1[oa], q1[ob]);
                    a[2] = (t == 0) ? pack_bf16(q0[oc], 0.0f) : 0u;               // k = 8: tap (2, 2)
                    a[3] = (t == 0) ? pack_bf16(q1[oc], 0.0f) : 0u;
                    bool keep[2], live[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int r = r0 + g + 8 * e;
                        live[e] = r < C::PH;                                      // the last m-tile of a column is short
                        const int Y = y0 + r, X = x0 + c;
                        keep[e] = interior || (Y >= 0 && Y < H && X >= 0 && X < W);   // outside the image: SAME padding
                    }
                    uint8_t *d0 = sl + (r0 * C::PW + c) * 16;
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                        mma_m16n8k16_bf16(acc, a, bwf[nt][0], bwf[nt][1]);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const float v0 = fmaxf(fmaf(acc[2 * e], sc1[nt][0], sh1[nt][0]), 0.0f);
                            const float v1 = fmaxf(fmaf(acc[2 * e + 1], sc1[nt][1], sh1[nt][1]), 0.0f);
                            if (live[e])
                                *reinterpret_cast<uint32_t *>(d0 + nt * (C::PH * C::PW * 16) + e * (8 * C::PW * 16)) =
                                    keep[e] ? pack_bf16(v0, v1) : 0u;
                        }
                    }
                }
                tc::fence_proxy_async();                    // generic-proxy writes -> visible to the UMMA reads
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&full_bar[stage]);
                if (++stage == nstages) { stage = 0; phase ^= 1; }
                buf ^= 1;
            }
        }
    } else {
        // ========================================================= epilogue
        const int q4 = warp & 3;                            // TMEM lane quarter this warp may read
        const int half = (warp - 1 - NMMA) >> 2;            // which epilogue group: splits the work items
        const int r = q4 * 32 + lane;                       // accumulator row = pixel of the sub-tile
        const int ph = r >> 3, pw = r & 7;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        const int CBo = COUT / 8;
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.0f, 0.0f);
        // tile coordinates advance incrementally (no division in the loop)
        const int tpi = tiles_x * tiles_y, gstep = (int)gridDim.x;
        int tx = (int)blockIdx.x % tiles_x, ty = ((int)blockIdx.x / tiles_x) % tiles_y, np = (int)blockIdx.x / tpi;
        const int dtx = gstep % tiles_x, dty = (gstep / tiles_x) % tiles_y, dnp = gstep / tpi;
        int it = 0;
        for (int t = blockIdx.x; CL ? (it < iters) : (t < ntiles); t += gridDim.x, ++it,
                 tx += dtx, ty += dty + (tx >= tiles_x ? 1 : 0), tx -= (tx >= tiles_x ? tiles_x : 0),
                 np += dnp + (ty >= tiles_y ? 1 : 0), ty -= (ty >= tiles_y ? tiles_y : 0)) {
            const int buf = it % NBUF;
            const bool live = !CL || t < ntiles;
            const int n = np * out_mul + out_off;                          // output "image" (n*D + z)
            tc::mbar_wait(&tfull_bar[buf], (it / NBUF) & 1);
            tc::tc_fence_after();
            if (EPI == EPI_POOL) {
                // sub-tile pair (2p, 2p+1) = image rows (y, y+1) in the same lane: store both rows of
                // the activation, take the vertical max in registers, the horizontal max with one
                // shuffle (lane^1) and let the even-x lanes store the pooled tensor
                static_assert(EPI != EPI_POOL || C::ILV, "fused pool needs interleaved sub-tiles");
#pragma unroll 1
                for (int jp = 0; jp < S / 2; ++jp) {
                    const int y = ty * C::TH + 32 * jp + 2 * ph, x = tx * 8 + pw;
                    const bool valid = live && (y < H) && (x < W);
                    // ONE 64-bit address chain per sub-tile pair; the 16-channel chunks are a constant byte stride apart
                    // (the per-item `out + (((n*CBo + 2*c16)*H + y)*W + x)*8` cost ~70 integer instructions per item)
                    const int Hp = H >> 1, Wp = W >> 1;
                    const size_t plane_b = (size_t)H * W * 8 * sizeof(bf16), pplane_b = (size_t)Hp * Wp * 8 * sizeof(bf16);
                    char *const pb = reinterpret_cast<char *>(out + (((size_t)n * CBo * H + y) * W + x) * 8);
                    char *const qb = reinterpret_cast<char *>(out_pool + (((size_t)n * CBo * Hp + (y >> 1)) * Wp + (x >> 1)) * 8);
#pragma unroll 1
                    for (int c16 = 0; c16 < COUT / 16; ++c16) {
                        if (((jp * (COUT / 16) + c16) % EPI_GROUPS) != half) continue;
                        uint32_t v0[16], v1[16];
                        const uint32_t col = tmem_base + lane_addr + buf * C::ACC_COLS + (2 * jp) * COUT + c16 * 16;
                        tc::tmem_ld16(col, v0);
                        tc::tmem_ld16(col + COUT, v1);
                        tc::tmem_ld_wait();
                        uint32_t o0[8], o1[8];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 sc = *reinterpret_cast<const float4 *>(s_scale + c16 * 16 + 4 * g);
                            const float4 sh = *reinterpret_cast<const float4 *>(s_shift + c16 * 16 + 4 * g);
                            // ReLU on the packed bf16 pair (max commutes with the monotone rounding): one HMNMX2 per pair
                            // instead of two FMNMX
                            o0[2 * g] = bf162_max(pack_bf16(fmaf(__uint_as_float(v0[4 * g]), sc.x, sh.x),
                                                            fmaf(__uint_as_float(v0[4 * g + 1]), sc.y, sh.y)), 0u);
                            o0[2 * g + 1] = bf162_max(pack_bf16(fmaf(__uint_as_float(v0[4 * g + 2]), sc.z, sh.z),
                                                                fmaf(__uint_as_float(v0[4 * g + 3]), sc.w, sh.w)), 0u);
                            o1[2 * g] = bf162_max(pack_bf16(fmaf(__uint_as_float(v1[4 * g]), sc.x, sh.x),
                                                            fmaf(__uint_as_float(v1[4 * g + 1]), sc.y, sh.y)), 0u);
                            o1[2 * g + 1] = bf162_max(pack_bf16(fmaf(__uint_as_float(v1[4 * g + 2]), sc.z, sh.z),
                                                                fmaf(__uint_as_float(v1[4 * g + 3]), sc.w, sh.w)), 0u);
                        }
                        if (valid) {
                            char *p = pb + (size_t)c16 * (2 * plane_b);
                            *reinterpret_cast<uint4 *>(p) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
                            *reinterpret_cast<uint4 *>(p + plane_b) = make_uint4(o0[4], o0[5], o0[6], o0[7]);
                            p += (size_t)W * 8 * sizeof(bf16);                 // row y + 1 (H is even)
                            *reinterpret_cast<uint4 *>(p) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
                            *reinterpret_cast<uint4 *>(p + plane_b) = make_uint4(o1[4], o1[5], o1[6], o1[7]);
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const uint32_t m = bf162_max(o0[e], o1[e]);
                            o0[e] = bf162_max(m, __shfl_xor_sync(0xffffffffu, m, 1));
                        }
                        if (valid && (lane & 1) == 0) {
                            char *q = qb + (size_t)c16 * (2 * pplane_b);
                            *reinterpret_cast<uint4 *>(q) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
                            *reinterpret_cast<uint4 *>(q + pplane_b) = make_uint4(o0[4], o0[5], o0[6], o0[7]);
                        }
                    }
                }
            } else if constexpr (CONST_EPI) {
                // group `half` owns the sub-tiles j = half, half + 2, ...; per sub-tile the 16-channel chunks go two per
                // TMEM round trip; scale / shift are constant-bank operands (epc), ReLU on the packed bf16 pair
                constexpr int NC16 = COUT / 16;
                const int x = tx * 8 + pw;
                const size_t plane = (size_t)H * W * 8;
#pragma unroll 1
                for (int j = half; j < S; j += EPI_GROUPS) {
                    const int y = ty * C::TH + (C::ILV ? 32 * (j / 2) + 2 * ph + (j & 1) : j * 16 + ph);
                    const bool valid = live && (y < H) && (x < W);
                    bf16 *prow = out + (((size_t)n * CBo * H + y) * W + x) * 8;
                    const uint32_t tj = tmem_base + lane_addr + buf * C::ACC_COLS + j * COUT;
#pragma unroll
                    for (int cp = 0; cp < NC16; cp += 2) {
                        uint32_t v[2][16];
                        tc::tmem_ld16(tj + cp * 16, v[0]);
                        tc::tmem_ld16(tj + cp * 16 + 16, v[1]);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            uint32_t o[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const int c = (cp + u) * 16 + 2 * e;            // compile-time after unrolling
                                __nv_bfloat162 p2 = __floats2bfloat162_rn(fmaf(__uint_as_float(v[u][2 * e]), epc.sc[c], epc.sh[c]),
                                                                          fmaf(__uint_as_float(v[u][2 * e + 1]), epc.sc[c + 1], epc.sh[c + 1]));
                                if (relu) p2 = __hmax2(p2, zero2);
                                o[e] = *reinterpret_cast<uint32_t *>(&p2);
                            }
                            if (valid) {
                                bf16 *p = prow + (size_t)(2 * (cp + u)) * plane;
                                *reinterpret_cast<uint4 *>(p) = make_uint4(o[0], o[1], o[2], o[3]);
                                *reinterpret_cast<uint4 *>(p + plane) = make_uint4(o[4], o[5], o[6], o[7]);
                            }
                        }
                    }
                }
            } else if (!UP && EPI == EPI_STORE) {
                // work items (sub-tile j, 16-channel chunk c16), two per round trip: both TMEM loads in flight
                // before the single wait; ReLU on the packed bf16 pair (max commutes with the monotone rounding)
                constexpr int NC16 = COUT / 16, NI = S * NC16;
                static_assert(UP || EPI != EPI_STORE || NI % (2 * EPI_GROUPS) == 0, "items come in pairs per group");
                const int x = tx * 8 + pw;
                const size_t plane = (size_t)H * W * 8;
#pragma unroll 1
                for (int i0 = half; i0 < NI; i0 += 2 * EPI_GROUPS) {
                    uint32_t v[2][16];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int i = i0 + u * EPI_GROUPS, j = i / NC16, c16 = i % NC16;
                        tc::tmem_ld16(tmem_base + lane_addr + buf * C::ACC_COLS + j * COUT + c16 * 16, v[u]);
                    }
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int i = i0 + u * EPI_GROUPS, j = i / NC16, c16 = i % NC16;
                        const int y = ty * C::TH + (C::ILV ? 32 * (j / 2) + 2 * ph + (j & 1) : j * 16 + ph);
                        uint32_t o[8];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 sc = *reinterpret_cast<const float4 *>(s_scale + c16 * 16 + 4 * g);
                            const float4 sh = *reinterpret_cast<const float4 *>(s_shift + c16 * 16 + 4 * g);
                            __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaf(__uint_as_float(v[u][4 * g]), sc.x, sh.x),
                                                                      fmaf(__uint_as_float(v[u][4 * g + 1]), sc.y, sh.y));
                            __nv_bfloat162 p1 = __floats2bfloat162_rn(fmaf(__uint_as_float(v[u][4 * g + 2]), sc.z, sh.z),
                                                                      fmaf(__uint_as_float(v[u][4 * g + 3]), sc.w, sh.w));
                            if (relu) { p0 = __hmax2(p0, zero2); p1 = __hmax2(p1, zero2); }
                            o[2 * g] = *reinterpret_cast<uint32_t *>(&p0);
                            o[2 * g + 1] = *reinterpret_cast<uint32_t *>(&p1);
                        }
                        if (live && (y < H) && (x < W)) {
                            bf16 *p = out + ((((size_t)n * CBo + c16 * 2) * H + y) * W + x) * 8;
                            *reinterpret_cast<uint4 *>(p) = make_uint4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<uint4 *>(p + plane) = make_uint4(o[4], o[5], o[6], o[7]);
                        }
                    }
                }
            } else
#pragma unroll 1
            for (int j = 0; j < S; ++j) {
                const int y = ty * C::TH + (C::ILV ? 32 * (j / 2) + 2 * ph + (j & 1) : j * 16 + ph);
                const int x = tx * 8 + pw;
                const bool valid = live && (y < H) && (x < W);
                if (!UP) {
                    // fused head (EPI_HEAD): the activation is consumed as it would have been stored (bf16)
                    if ((j % EPI_GROUPS) != half) continue;
                    constexpr int NC16 = COUT / 16;
                    float hl[HK > 0 ? HK : 1];                     // running logits
#pragma unroll
                    for (int k = 0; k < HK; ++k) hl[k] = 0.0f;
                    uint32_t v[NC16][16];
#pragma unroll
                    for (int c16 = 0; c16 < NC16; ++c16)
                        tc::tmem_ld16(tmem_base + lane_addr + buf * C::ACC_COLS + j * COUT + c16 * 16, v[c16]);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int c16 = 0; c16 < NC16; ++c16) {
                        float f[16];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 sc = *reinterpret_cast<const float4 *>(s_scale + c16 * 16 + 4 * g);
                            const float4 sh = *reinterpret_cast<const float4 *>(s_shift + c16 * 16 + 4 * g);
                            __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaf(__uint_as_float(v[c16][4 * g]), sc.x, sh.x),
                                                                      fmaf(__uint_as_float(v[c16][4 * g + 1]), sc.y, sh.y));
                            __nv_bfloat162 p1 = __floats2bfloat162_rn(fmaf(__uint_as_float(v[c16][4 * g + 2]), sc.z, sh.z),
                                                                      fmaf(__uint_as_float(v[c16][4 * g + 3]), sc.w, sh.w));
                            if (relu) { p0 = __hmax2(p0, zero2); p1 = __hmax2(p1, zero2); }
                            const float2 t0 = __bfloat1622float2(p0), t1 = __bfloat1622float2(p1);
                            f[4 * g] = t0.x; f[4 * g + 1] = t0.y; f[4 * g + 2] = t1.x; f[4 * g + 3] = t1.y;
                        }
#pragma unroll
                        for (int k = 0; k < HK; ++k) {
                            const float4 *wk = reinterpret_cast<const float4 *>(s_head + k * COUT + c16 * 16);
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const float4 w4 = wk[g];
                                hl[k] = fmaf(f[4 * g], w4.x, hl[k]);
                                hl[k] = fmaf(f[4 * g + 1], w4.y, hl[k]);
                                hl[k] = fmaf(f[4 * g + 2], w4.z, hl[k]);
                                hl[k] = fmaf(f[4 * g + 3], w4.w, hl[k]);
                            }
                        }
                    }
                    {
                        const size_t p = ((size_t)n * H + y) * W + x;
                        int best = 0;
                        float m = -INFINITY;
#pragma unroll
                        for (int k = 0; k < HK; ++k) {
                            hl[k] += s_head[COUT * HK + k];
                            if (hl[k] > m) { m = hl[k]; best = k; }
                        }
                        if (head.mask) {
                            // 8 pixels of one image row sit in 8 consecutive lanes: gather their class
                            // bytes so that one lane writes 8 contiguous bytes
                            uint32_t lo = (uint32_t)best << (8 * (lane & 3));
                            lo |= __shfl_xor_sync(0xffffffffu, lo, 1);
                            lo |= __shfl_xor_sync(0xffffffffu, lo, 2);
                            const uint32_t hi = __shfl_down_sync(0xffffffffu, lo, 4);
                            if (valid && (lane & 7) == 0) {
                                if (x + 8 <= W && (W & 7) == 0)
                                    *reinterpret_cast<uint2 *>(head.mask + p) = make_uint2(lo, hi);
                                else
                                    for (int i = 0; i < 8 && x + i < W; ++i)
                                        head.mask[p + i] = (uint8_t)((i < 4 ? lo >> (8 * i) : hi >> (8 * (i - 4))) & 0xff);
                            }
                        }
                        if (valid && head.logits) {
#pragma unroll
                            for (int k = 0; k < HK; ++k) head.logits[p * HK + k] = hl[k];
                        }
                        if (valid && head.probs) {
                            float sum = 0.0f;
#pragma unroll
                            for (int k = 0; k < HK; ++k) { hl[k] = expf(hl[k] - m); sum += hl[k]; }
#pragma unroll
                            for (int k = 0; k < HK; ++k) head.probs[p * HK + k] = hl[k] / sum;
                        }
                    }
                } else {
                    const int Ho = 2 * H, Wo = 2 * W;
#pragma unroll 1
                    for (int ky = 0; ky < 2; ++ky) {
                        if (((j * 2 + ky) % EPI_GROUPS) != half) continue;
#pragma unroll 1
                        for (int c8 = 0; c8 < COUT / 8; ++c8) {
                            uint32_t v0[8], v1[8];
                            const uint32_t col = tmem_base + lane_addr + buf * C::ACC_COLS +
                                                 (j * 4 + ky * 2) * COUT + c8 * 8;
                            tc::tmem_ld8(col, v0);
                            tc::tmem_ld8(col + COUT, v1);
                            tc::tmem_ld_wait();
                            uint32_t o0[4], o1[4];
#pragma unroll
                            for (int g = 0; g < 2; ++g) {
                                const float4 sh = *reinterpret_cast<const float4 *>(s_shift + c8 * 8 + 4 * g);
                                o0[2 * g] = pack_bf16(__uint_as_float(v0[4 * g]) + sh.x, __uint_as_float(v0[4 * g + 1]) + sh.y);
                                o0[2 * g + 1] = pack_bf16(__uint_as_float(v0[4 * g + 2]) + sh.z, __uint_as_float(v0[4 * g + 3]) + sh.w);
                                o1[2 * g] = pack_bf16(__uint_as_float(v1[4 * g]) + sh.x, __uint_as_float(v1[4 * g + 1]) + sh.y);
                                o1[2 * g + 1] = pack_bf16(__uint_as_float(v1[4 * g + 2]) + sh.z, __uint_as_float(v1[4 * g + 3]) + sh.w);
                            }
                            if (valid) {
                                bf16 *p = out + ((((size_t)n * CBo + c8) * Ho + 2 * y + ky) * Wo + 2 * x) * 8;
                                *reinterpret_cast<uint4 *>(p) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
                                *reinterpret_cast<uint4 *>(p + 8) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
                            }
                        }
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty_bar[buf]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
    if (CL) tc::cluster_sync_all();                         // the peer may still signal this CTA's barriers
}

// ------------------------------------------------- x-combined 3x3 conv (Cout <= 32)
// With N = Cout <= 32 the 128 x N x 16 MMA above is bound by its A-operand fetch (4 KB from shared
// memory per instruction, ~32 clk) rather than by the tensor math (N/2 clk), and a 3x3 conv pays
// that fetch nine times per 16 input channels.  This variant folds the three horizontal taps into
// the N dimension instead: B_ky = [(kx, co)][ci] has N = 3*Cout rows, so ONE instruction per row
// tap ky computes, for every INPUT pixel of the tile, its contribution to the three outputs it
// feeds (D'[xi][kx][co]); three instructions per k-step instead of nine, each with 3x the math per
// fetched byte.  The epilogue finishes the conv with two lane shuffles:
//     out[x][co] = D'[x-1][0][co] + D'[x][1][co] + D'[x+1][2][co]
// which is why a sub-tile is 8 rows x 16 INPUT columns (one warp = 2 rows of 16 lanes; the outer
// two columns are the x halo and produce no output): 8 x 14 outputs per 128-lane accumulator.
template <int COUT, int S, bool PADACC = true>
struct XCfg {
    static constexpr int N = 3 * COUT;                   // MMA N: (kx, co)
    static constexpr int TW = 14, TH = 8 * S;            // output tile (pixels)
    static constexpr int PW = 16, PH = TH + 2;           // input patch incl. halo
    static constexpr int A_BYTES = 2 * PH * PW * 16;     // 16 input channels of the patch
    static constexpr int B_BYTES = 3 * 2 * N * 16;       // 3 row taps x 16 input channels x N
    static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES + 127) / 128 * 128;   // streaming weights
    static constexpr int ACC_STRIDE = !PADACC ? N : (N <= 64) ? 64 : 128;   // TMEM columns per sub-tile
    static constexpr int ACC_COLS = S * ACC_STRIDE;
    static constexpr uint32_t LBO_A = PH * PW * 16, SBO_A = 128;   // 8-pixel groups are contiguous
    static constexpr uint32_t LBO_B = N * 16, SBO_B = 128;
    static_assert(A_BYTES % 128 == 0, "TMA destination alignment");
};

constexpr int XC_MAX_STAGES = 12;

template <int COUT, int S, int NBUF, int MINB, int EPI, int HK, bool PADACC, int KPS, int ZPS>
__global__ void __launch_bounds__(TC_THREADS, MINB)
conv_xc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               int ks0, int ks1, const bf16 *__restrict__ wts, const float *__restrict__ scale,
               const float *__restrict__ shift, bf16 *__restrict__ out, bf16 *__restrict__ out_pool,
               HeadArgs head, int nimg, int H, int W, int nstages, int D, int KZ, int wres, long long *phase_dbg)
{
    // phase_dbg (build with -DSQ_XC_PHASE_DIAG, run with SQ_XC_PHASE=1; diagnostics only): per-CTA
    // clocks each role spends waiting / working (scripts/xc_phase.py, profiles/r1_xc_phase.log)
#ifdef SQ_XC_PHASE_DIAG
#define XC_T0() const long long t_ = phase_dbg ? clock64() : 0
#define XC_ACC(var) if (phase_dbg) var += clock64() - t_
#else
#define XC_T0() do { } while (0)
#define XC_ACC(var) do { } while (0)
#endif
    // The weights of ALL k-steps (wres bytes) are loaded once per CTA and stay in shared memory; the
    // ring stages carry activation patches only (with tiles this small, re-fetching the weights per
    // tile would make the kernel L2-bandwidth bound).
    // One TMA instruction costs ~450-500 clk of issue time per SM whatever the box size
    // (profiles/r1_tma_rate_probe_*), so small boxes starve the MMA pipe.  A ring stage therefore
    // holds ZPS x KPS 16-channel k-steps fetched by ONE box: KPS pairs of channel blocks (box
    // channel-block extent 2*KPS) and, for volumes, all ZPS = 3 depth taps (box z extent 3, slices
    // outside the volume zero-filled).  ZPS = 1: depth taps (if any) are separate stages.
    using C = XCfg<COUT, S, PADACC>;
    constexpr int A_STAGE = ZPS * KPS * C::A_BYTES;
    constexpr int TMEM_COLS = (NBUF * C::ACC_COLS <= 64) ? 64 : (NBUF * C::ACC_COLS <= 128) ? 128
                            : (NBUF * C::ACC_COLS <= 256) ? 256 : 512;
    static_assert(NBUF * C::ACC_COLS <= 512, "TMEM overflow");
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full_bar[XC_MAX_STAGES], empty_bar[XC_MAX_STAGES], tfull_bar[2], tempty_bar[2], w_bar;
    __shared__ uint32_t tmem_base_sh;
    __shared__ __align__(16) float s_scale[COUT], s_shift[COUT];
    __shared__ __align__(16) float s_head[EPI == EPI_HEAD ? COUT * HK + HK : 4];   // [k][c] then bias[k]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (W + C::TW - 1) / C::TW, tiles_y = (H + C::TH - 1) / C::TH;
    const int ntiles = nimg * D * tiles_x * tiles_y;
    const int ksteps = ks0 + ks1, qsteps = KZ * ksteps;
    const int zgroups = KZ / ZPS;                           // depth-tap groups fetched as separate stages
    uint8_t *ring = smem + ((wres + 127) & ~127);           // stages start after the resident weights

    if (threadIdx.x == 0) {
        for (int i = 0; i < nstages; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull_bar[i], 1); tc::mbar_init(&tempty_bar[i], 4 * EPI_GROUPS); }
        tc::mbar_init(&w_bar, 1);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&mapA0);
        tc::tma_prefetch_desc(&mapA1);
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_base_sh, TMEM_COLS); tc::tmem_relinquish(); }
    for (int i = threadIdx.x; i < COUT; i += TC_THREADS) { s_scale[i] = scale[i]; s_shift[i] = shift[i]; }
    if constexpr (EPI == EPI_HEAD) {
        for (int i = threadIdx.x; i < COUT * HK; i += TC_THREADS) s_head[(i % HK) * COUT + i / HK] = head.w[i];
        for (int i = threadIdx.x; i < HK; i += TC_THREADS) s_head[COUT * HK + i] = head.w[COUT * HK + i];
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
#ifdef SQ_XC_PHASE_DIAG
            long long d_wait0 = 0, d_work = 0;
#endif
            tc::mbar_arrive_expect_tx(&w_bar, (uint32_t)wres);
            for (int q = 0; q < qsteps; ++q)
                tc::bulk_load(smem + (size_t)q * C::B_BYTES, wts + (size_t)q * (C::B_BYTES / 2), C::B_BYTES, &w_bar);
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, np = t / (tiles_x * tiles_y);
                const int z = np % D, n = np / D;
                const int x0 = tx * C::TW - 1, y0 = ty * C::TH - 1;
                for (int zg = 0; zg < zgroups; ++zg) {
                    const int zc = z + zg - (KZ >> 1);      // first slice of the box (ZPS = 3: z - 1)
                    for (int ks = 0; ks < ksteps; ks += KPS) {
                        { XC_T0(); tc::mbar_wait(&empty_bar[stage], phase ^ 1); XC_ACC(d_wait0); }
                        XC_T0();
                        tc::mbar_arrive_expect_tx(&full_bar[stage], A_STAGE);
                        uint8_t *sA = ring + (size_t)stage * A_STAGE;
                        if (ks < ks0) tc::tma_load_5d(sA, &mapA0, &full_bar[stage], x0 * 8, y0, ks * 2, zc, n);
                        else          tc::tma_load_5d(sA, &mapA1, &full_bar[stage], x0 * 8, y0, (ks - ks0) * 2, zc, n);
                        XC_ACC(d_work);
                        if (++stage == nstages) { stage = 0; phase ^= 1; }
                    }
                }
            }
#ifdef SQ_XC_PHASE_DIAG
            if (phase_dbg) { phase_dbg[blockIdx.x * 8 + 0] = d_wait0; phase_dbg[blockIdx.x * 8 + 1] = d_work; }
#endif
        }
    } else if (warp == 1) {
        // ======================================================= MMA issuer
        const uint32_t idesc = tc::instr_desc_bf16(128, C::N);
        const uint32_t a_hi = ((C::SBO_A >> 4) & 0x3FFFu) | (1u << 14);
        const uint32_t b_hi = ((C::SBO_B >> 4) & 0x3FFFu) | (1u << 14);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        tc::mbar_wait(&w_bar, 0);
#ifdef SQ_XC_PHASE_DIAG
        long long d_wait0 = 0, d_wait1 = 0, d_work = 0;
#endif
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int buf = it % NBUF;
            { XC_T0(); tc::mbar_wait(&tempty_bar[buf], ((it / NBUF) & 1) ^ 1); XC_ACC(d_wait0); }
            tc::tc_fence_after();
            for (int zg = 0; zg < zgroups; ++zg) {
                for (int ks = 0; ks < ksteps; ks += KPS) {
                    { XC_T0(); tc::mbar_wait(&full_bar[stage], phase); XC_ACC(d_wait1); }
                    tc::tc_fence_after();
                    XC_T0();
                    if (tc::elect_one()) {
                        const uint32_t a_base = tc::smem_u32(ring + (size_t)stage * A_STAGE);
                        const uint32_t a_lo = ((a_base >> 4) & 0x3FFFu) | (((C::LBO_A >> 4) & 0x3FFFu) << 16);
                        const uint32_t d0 = tmem_base + buf * C::ACC_COLS;
                        const uint32_t first = (zg > 0 || ks > 0) ? 1u : 0u;
#pragma unroll
                        for (int dz = 0; dz < ZPS; ++dz) {
                            // weights of k-step q = (depth tap) * ksteps + ks (+ kk)
                            const uint32_t b_base = tc::smem_u32(smem) + (uint32_t)(((zg * ZPS + dz) * ksteps + ks) * C::B_BYTES);
                            const uint32_t b_lo = ((b_base >> 4) & 0x3FFFu) | (((C::LBO_B >> 4) & 0x3FFFu) << 16);
#pragma unroll
                            for (int kk = 0; kk < KPS; ++kk) {
#pragma unroll
                                for (int j = 0; j < S; ++j) {
#pragma unroll
                                    for (int ky = 0; ky < 3; ++ky)
                                        tc::umma_bf16_parts(d0 + j * C::ACC_STRIDE,
                                                            a_lo + (uint32_t)((dz * KPS + kk) * (C::A_BYTES >> 4) + (j * 8 + ky) * C::PW), a_hi,
                                                            b_lo + (uint32_t)(kk * (C::B_BYTES >> 4) + ky * 2 * C::N), b_hi, idesc,
                                                            (ky == 0 && kk == 0 && dz == 0) ? first : 1u);
                                }
                            }
                        }
                        tc::umma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    XC_ACC(d_work);
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
            }
            if (tc::elect_one()) tc::umma_commit(&tfull_bar[buf]);
            __syncwarp();
        }
#ifdef SQ_XC_PHASE_DIAG
        if (phase_dbg && lane == 0) {
            phase_dbg[blockIdx.x * 8 + 2] = d_wait0; phase_dbg[blockIdx.x * 8 + 3] = d_wait1; phase_dbg[blockIdx.x * 8 + 4] = d_work;
        }
#endif
    } else {
        // ========================================================= epilogue
        const int q4 = warp & 3;                            // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;                   // epilogue group
        const int ph = q4 * 2 + (lane >> 4), pw = lane & 15;   // sub-tile row, INPUT column of this lane
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        constexpr int CBo = COUT / 8, NC16 = COUT / 16;
        const bool colok = (pw >= 1) && (pw <= C::TW);
        const size_t plane = (size_t)H * W * 8;
        // tile coordinates advance incrementally (no division in the loop)
        const int tpi = tiles_x * tiles_y, gstep = (int)gridDim.x;
        int tx = (int)blockIdx.x % tiles_x, ty = ((int)blockIdx.x / tiles_x) % tiles_y, n = (int)blockIdx.x / tpi;
        const int dtx = gstep % tiles_x, dty = (gstep / tiles_x) % tiles_y, dn = gstep / tpi;
        int it = 0;
#ifdef SQ_XC_PHASE_DIAG
        long long d_wait0 = 0, d_work = 0;
#endif
        for (int t = blockIdx.x; t < ntiles; t += gstep, ++it) {
            const int buf = it % NBUF;
            { XC_T0(); tc::mbar_wait(&tfull_bar[buf], (it / NBUF) & 1); XC_ACC(d_wait0); }
            tc::tc_fence_after();
            XC_T0();
            const int x = tx * C::TW + pw - 1;
            const bool xok = colok && (x < W);
            const uint32_t tbase = tmem_base + lane_addr + buf * C::ACC_COLS;
            if (EPI != EPI_HEAD) {
#pragma unroll
                for (int i = 0; i < (S * NC16 + EPI_GROUPS - 1) / EPI_GROUPS; ++i) {
                    const int item = i * EPI_GROUPS + half;             // (sub-tile, 16-channel chunk)
                    if (item >= S * NC16) break;
                    const int j = item / NC16, c16 = item % NC16;
                    const int y = ty * C::TH + j * 8 + ph;
                    const bool valid = xok && (y < H);
                    uint32_t a0[16], a1[16], a2[16];
                    const uint32_t col = tbase + j * C::ACC_STRIDE + c16 * 16;
                    tc::tmem_ld16(col, a0);
                    tc::tmem_ld16(col + COUT, a1);
                    tc::tmem_ld16(col + 2 * COUT, a2);
                    tc::tmem_ld_wait();
                    uint32_t o[8];
                    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.0f, 0.0f);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 sc = *reinterpret_cast<const float4 *>(s_scale + c16 * 16 + 4 * g);
                        const float4 sh = *reinterpret_cast<const float4 *>(s_shift + c16 * 16 + 4 * g);
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            v[e] = __shfl_up_sync(0xffffffffu, __uint_as_float(a0[4 * g + e]), 1) +
                                   __uint_as_float(a1[4 * g + e]) +
                                   __shfl_down_sync(0xffffffffu, __uint_as_float(a2[4 * g + e]), 1);
                        // ReLU after the bf16 rounding (max commutes with the monotone rounding)
                        __nv_bfloat162 p0 = __hmax2(__floats2bfloat162_rn(fmaf(v[0], sc.x, sh.x), fmaf(v[1], sc.y, sh.y)), zero2);
                        __nv_bfloat162 p1 = __hmax2(__floats2bfloat162_rn(fmaf(v[2], sc.z, sh.z), fmaf(v[3], sc.w, sh.w)), zero2);
                        o[2 * g] = *reinterpret_cast<uint32_t *>(&p0);
                        o[2 * g + 1] = *reinterpret_cast<uint32_t *>(&p1);
                    }
                    if (valid) {
                        bf16 *p = out + ((((size_t)n * CBo + c16 * 2) * H + y) * W + x) * 8;
                        *reinterpret_cast<uint4 *>(p) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4 *>(p + plane) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                    if (EPI == EPI_POOL) {
                        // rows (y, y+1) sit 16 lanes apart, columns (x, x+1) in lanes (pw, pw+1), pw odd
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const uint32_t m = bf162_max(o[e], __shfl_xor_sync(0xffffffffu, o[e], 16));
                            o[e] = bf162_max(m, __shfl_down_sync(0xffffffffu, m, 1));
                        }
                        if (valid && lane < 16 && (pw & 1)) {
                            const int Hp = H >> 1, Wp = W >> 1;
                            bf16 *p = out_pool + ((((size_t)n * CBo + c16 * 2) * Hp + (y >> 1)) * Wp + (x >> 1)) * 8;
                            *reinterpret_cast<uint4 *>(p) = make_uint4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<uint4 *>(p + (size_t)Hp * Wp * 8) = make_uint4(o[4], o[5], o[6], o[7]);
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int j = half; j < S; j += EPI_GROUPS) {
                    const int y = ty * C::TH + j * 8 + ph;
                    const bool valid = xok && (y < H);
                    float hl[HK > 0 ? HK : 1];
#pragma unroll
                    for (int k = 0; k < HK; ++k) hl[k] = 0.0f;
#pragma unroll 1
                    for (int c16 = 0; c16 < NC16; ++c16) {
                        uint32_t a0[16], a1[16], a2[16];
                        const uint32_t col = tbase + j * C::ACC_STRIDE + c16 * 16;
                        tc::tmem_ld16(col, a0);
                        tc::tmem_ld16(col + COUT, a1);
                        tc::tmem_ld16(col + 2 * COUT, a2);
                        tc::tmem_ld_wait();
                        float f[16];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 sc = *reinterpret_cast<const float4 *>(s_scale + c16 * 16 + 4 * g);
                            const float4 sh = *reinterpret_cast<const float4 *>(s_shift + c16 * 16 + 4 * g);
                            float v[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                v[e] = __shfl_up_sync(0xffffffffu, __uint_as_float(a0[4 * g + e]), 1) +
                                       __uint_as_float(a1[4 * g + e]) +
                                       __shfl_down_sync(0xffffffffu, __uint_as_float(a2[4 * g + e]), 1);
                            // the head consumes the activation as it would have been stored (bf16)
                            const float2 t0 = __bfloat1622float2(__floats2bfloat162_rn(
                                fmaxf(fmaf(v[0], sc.x, sh.x), 0.0f), fmaxf(fmaf(v[1], sc.y, sh.y), 0.0f)));
                            const float2 t1 = __bfloat1622float2(__floats2bfloat162_rn(
                                fmaxf(fmaf(v[2], sc.z, sh.z), 0.0f), fmaxf(fmaf(v[3], sc.w, sh.w), 0.0f)));
                            f[4 * g] = t0.x; f[4 * g + 1] = t0.y; f[4 * g + 2] = t1.x; f[4 * g + 3] = t1.y;
                        }
#pragma unroll
                        for (int k = 0; k < HK; ++k) {
                            const float4 *wk = reinterpret_cast<const float4 *>(s_head + k * COUT + c16 * 16);
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const float4 w4 = wk[g];
                                hl[k] = fmaf(f[4 * g], w4.x, hl[k]);
                                hl[k] = fmaf(f[4 * g + 1], w4.y, hl[k]);
                                hl[k] = fmaf(f[4 * g + 2], w4.z, hl[k]);
                                hl[k] = fmaf(f[4 * g + 3], w4.w, hl[k]);
                            }
                        }
                    }
                    const size_t p = ((size_t)n * H + y) * W + x;
                    int best = 0;
                    float m = -INFINITY;
#pragma unroll
                    for (int k = 0; k < HK; ++k) {
                        hl[k] += s_head[COUT * HK + k];
                        if (hl[k] > m) { m = hl[k]; best = k; }
                    }
                    if (valid && head.mask) head.mask[p] = (uint8_t)best;
                    if (valid && head.logits) {
#pragma unroll
                        for (int k = 0; k < HK; ++k) head.logits[p * HK + k] = hl[k];
                    }
                    if (valid && head.probs) {
                        float sum = 0.0f;
#pragma unroll
                        for (int k = 0; k < HK; ++k) { hl[k] = expf(hl[k] - m); sum += hl[k]; }
#pragma unroll
                        for (int k = 0; k < HK; ++k) head.probs[p * HK + k] = hl[k] / sum;
                    }
                }
            }
            tx += dtx;
            if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
            ty += dty;
            if (ty >= tiles_y) { ty -= tiles_y; ++n; }
            n += dn;
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty_bar[buf]);
            XC_ACC(d_work);
        }
#ifdef SQ_XC_PHASE_DIAG
        if (phase_dbg && warp == 2 && lane == 0) { phase_dbg[blockIdx.x * 8 + 5] = d_wait0; phase_dbg[blockIdx.x * 8 + 6] = d_work; }
#endif
    }
#undef XC_T0
#undef XC_ACC
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------- fused PAIR of x-combined 3x3 convs (conv1 -> conv2 on chip)
// conv_block of the reference (networks/unet.py:265-277) as ONE launch for Cout <= 32 planar layers: the
// intermediate activation (conv1's output) never goes to HBM.  Per tile:
//     TMA: (8S+2) x 16 input patch(es) -> stage 1 (x-combined MMAs, as conv_xc_kernel) -> TMEM
//     epilogue 1: TMEM -> combine / scale / shift / ReLU / bf16 -> shared-memory patch P1 (zeros outside the
//                 image = conv2's SAME padding), laid out exactly like a TMA-fetched patch
//     stage 2: x-combined MMAs reading P1 -> the SAME TMEM columns -> epilogue 2 (store / pool / fused head)
// Geometry: 16 input columns -> 14 columns of P1 -> 12 output columns; 8S+2 input rows -> 8S rows of P1 ->
// 8S-2 output rows (S = 2: a 14 x 12 output tile from an 18 x 16 input patch).  The MMA warp issues
// stage 1 of tile i+1 BEFORE stage 2 of tile i and the epilogue warps mirror that order, so the tensor pipe
// works on the next tile while P1 of the current one is being written; two accumulator buffers, two CTAs
// per SM.  Weights of both convs stay resident in shared memory.
template <int COUT, int S>
struct PCfg {
    static constexpr int N = 3 * COUT;
    static constexpr int TW = 12, TH = 8 * S - 2;        // output tile
    static constexpr int PW = 16, PH = 8 * S + 2;        // input patch of stage 1 == allocated P1 patch
    static constexpr int A_BYTES = 2 * PH * PW * 16;     // 16 channels of a patch
    static constexpr int P1_BYTES = (COUT / 16) * A_BYTES;
    static constexpr int B_BYTES = 3 * 2 * N * 16;       // one k-step of x-combined weights
    static constexpr int ACC_STRIDE = (N <= 64) ? 64 : 128;
    static constexpr int ACC_COLS = S * ACC_STRIDE;
    static constexpr uint32_t LBO_A = PH * PW * 16, SBO_A = 128, LBO_B = N * 16, SBO_B = 128;
};

template <int COUT, int S, int EPI, int HK>
__global__ void __launch_bounds__(TC_THREADS, 2)
conv_xc_pair_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                    int ks0, int ks1, const bf16 *__restrict__ w1, const float *__restrict__ scale1,
                    const float *__restrict__ shift1, const bf16 *__restrict__ w2,
                    const float *__restrict__ scale2, const float *__restrict__ shift2,
                    bf16 *__restrict__ out, bf16 *__restrict__ out_pool, HeadArgs head, int nimg, int H, int W,
                    int nstages)
{
    using C = PCfg<COUT, S>;
    constexpr int K2 = COUT / 16, NC16 = COUT / 16, CBo = COUT / 8;
    constexpr int TMEM_COLS = (2 * C::ACC_COLS <= 64) ? 64 : (2 * C::ACC_COLS <= 128) ? 128
                            : (2 * C::ACC_COLS <= 256) ? 256 : 512;
    static_assert(2 * 2 * C::ACC_COLS <= 512, "two CTAs per SM must fit in TMEM");
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full_bar[XC_MAX_STAGES], empty_bar[XC_MAX_STAGES], w_bar;
    __shared__ uint64_t acc1_full[2], p1_full[2], acc2_full[2], acc_free[2];
    __shared__ uint32_t tmem_base_sh;
    __shared__ __align__(16) float s_scale1[COUT], s_shift1[COUT], s_scale2[COUT], s_shift2[COUT];
    __shared__ __align__(16) float s_head[EPI == EPI_HEAD ? COUT * HK + HK : 4];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (W + C::TW - 1) / C::TW, tiles_y = (H + C::TH - 1) / C::TH;
    const int tpi = tiles_x * tiles_y, ntiles = nimg * tpi;
    const int ksteps1 = ks0 + ks1;
    const int w1_bytes = ksteps1 * C::B_BYTES, w2_bytes = K2 * C::B_BYTES;
    const int stage_bytes = ksteps1 * C::A_BYTES;
    uint8_t *sW1 = smem, *sW2 = smem + w1_bytes;
    uint8_t *sP1 = smem + ((w1_bytes + w2_bytes + 127) & ~127);
    uint8_t *ring = sP1 + 2 * C::P1_BYTES;

    if (threadIdx.x == 0) {
        for (int i = 0; i < nstages; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&acc1_full[i], 1);
            tc::mbar_init(&p1_full[i], 4 * EPI_GROUPS);
            tc::mbar_init(&acc2_full[i], 1);
            tc::mbar_init(&acc_free[i], 4 * EPI_GROUPS);
        }
        tc::mbar_init(&w_bar, 1);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&mapA0);
        tc::tma_prefetch_desc(&mapA1);
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_base_sh, TMEM_COLS); tc::tmem_relinquish(); }
    for (int i = threadIdx.x; i < COUT; i += TC_THREADS) {
        s_scale1[i] = scale1[i]; s_shift1[i] = shift1[i]; s_scale2[i] = scale2[i]; s_shift2[i] = shift2[i];
    }
    if constexpr (EPI == EPI_HEAD) {
        for (int i = threadIdx.x; i < COUT * HK; i += TC_THREADS) s_head[(i % HK) * COUT + i / HK] = head.w[i];
        for (int i = threadIdx.x; i < HK; i += TC_THREADS) s_head[COUT * HK + i] = head.w[COUT * HK + i];
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(&w_bar, (uint32_t)(w1_bytes + w2_bytes));
            for (int q = 0; q < ksteps1; ++q)
                tc::bulk_load(sW1 + (size_t)q * C::B_BYTES, w1 + (size_t)q * (C::B_BYTES / 2), C::B_BYTES, &w_bar);
            for (int q = 0; q < K2; ++q)
                tc::bulk_load(sW2 + (size_t)q * C::B_BYTES, w2 + (size_t)q * (C::B_BYTES / 2), C::B_BYTES, &w_bar);
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, n = t / tpi;
                const int x0 = tx * C::TW - 2, y0 = ty * C::TH - 2;
                tc::mbar_wait(&empty_bar[stage], phase ^ 1);
                tc::mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
                uint8_t *sA = ring + (size_t)stage * stage_bytes;
                tc::tma_load_5d(sA, &mapA0, &full_bar[stage], x0 * 8, y0, 0, 0, n);        // all ks0 k-steps, one box
                if (ks1) tc::tma_load_5d(sA + (size_t)ks0 * C::A_BYTES, &mapA1, &full_bar[stage], x0 * 8, y0, 0, 0, n);
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ======================================================= MMA issuer
        const uint32_t idesc = tc::instr_desc_bf16(128, C::N);
        const uint32_t a_hi = ((C::SBO_A >> 4) & 0x3FFFu) | (1u << 14);
        const uint32_t b_hi = ((C::SBO_B >> 4) & 0x3FFFu) | (1u << 14);
        const uint32_t lbo_a = ((C::LBO_A >> 4) & 0x3FFFu) << 16, lbo_b = ((C::LBO_B >> 4) & 0x3FFFu) << 16;
        int stage = 0;
        uint32_t phase = 0;
        tc::mbar_wait(&w_bar, 0);
        // stage 2 of this CTA's tile number i: P1[i & 1] x W2 -> accumulator buffer i & 1
        auto stage2 = [&](int i) {
            const int b = i & 1;
            tc::mbar_wait(&p1_full[b], (i >> 1) & 1);
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint32_t a_lo = ((tc::smem_u32(sP1 + (size_t)b * C::P1_BYTES) >> 4) & 0x3FFFu) | lbo_a;
                const uint32_t b_lo = ((tc::smem_u32(sW2) >> 4) & 0x3FFFu) | lbo_b;
                const uint32_t d0 = tmem_base + b * C::ACC_COLS;
#pragma unroll
                for (int k = 0; k < K2; ++k)
#pragma unroll
                    for (int j = 0; j < S; ++j)
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
                            tc::umma_bf16_parts(d0 + j * C::ACC_STRIDE,
                                                a_lo + (uint32_t)(k * (C::A_BYTES >> 4) + (j * 8 + ky) * C::PW), a_hi,
                                                b_lo + (uint32_t)(k * (C::B_BYTES >> 4) + ky * 2 * C::N), b_hi, idesc,
                                                (k == 0 && ky == 0) ? 0u : 1u);
                tc::umma_commit(&acc2_full[b]);
            }
            __syncwarp();
        };
        for (int it = 0; it < my_tiles; ++it) {
            const int b = it & 1;
            tc::mbar_wait(&acc_free[b], ((it >> 1) & 1) ^ 1);
            tc::mbar_wait(&full_bar[stage], phase);
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint32_t a_lo = ((tc::smem_u32(ring + (size_t)stage * stage_bytes) >> 4) & 0x3FFFu) | lbo_a;
                const uint32_t b_lo = ((tc::smem_u32(sW1) >> 4) & 0x3FFFu) | lbo_b;
                const uint32_t d0 = tmem_base + b * C::ACC_COLS;
                for (int k = 0; k < ksteps1; ++k)
#pragma unroll
                    for (int j = 0; j < S; ++j)
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
                            tc::umma_bf16_parts(d0 + j * C::ACC_STRIDE,
                                                a_lo + (uint32_t)(k * (C::A_BYTES >> 4) + (j * 8 + ky) * C::PW), a_hi,
                                                b_lo + (uint32_t)(k * (C::B_BYTES >> 4) + ky * 2 * C::N), b_hi, idesc,
                                                (k == 0 && ky == 0) ? 0u : 1u);
                tc::umma_commit(&empty_bar[stage]);
                tc::umma_commit(&acc1_full[b]);
            }
            __syncwarp();
            if (++stage == nstages) { stage = 0; phase ^= 1; }
            if (it > 0) stage2(it - 1);
        }
        stage2(my_tiles - 1);
    } else {
        // ========================================================= epilogue
        const int q4 = warp & 3;
        const int half = (warp - 2) >> 2;
        const int ph = q4 * 2 + (lane >> 4), pw = lane & 15;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.0f, 0.0f);
        // x-combination of one 16-channel chunk of sub-tile j: v[e] = D'[x-1][0] + D'[x][1] + D'[x+1][2]
        auto combine = [&](uint32_t col, float *v) {
            uint32_t a0[16], a1[16], a2[16];
            tc::tmem_ld16(col, a0);
            tc::tmem_ld16(col + COUT, a1);
            tc::tmem_ld16(col + 2 * COUT, a2);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e)
                v[e] = __shfl_up_sync(0xffffffffu, __uint_as_float(a0[e]), 1) + __uint_as_float(a1[e]) +
                       __shfl_down_sync(0xffffffffu, __uint_as_float(a2[e]), 1);
        };
        // epilogue 1 of this CTA's tile number i: accumulators -> P1[i & 1]
        auto epi1 = [&](int i) {
            const int b = i & 1;
            const int t = (int)blockIdx.x + i * (int)gridDim.x;
            const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y;
            tc::mbar_wait(&acc1_full[b], (i >> 1) & 1);
            tc::tc_fence_after();
            const uint32_t tbase = tmem_base + lane_addr + b * C::ACC_COLS;
            uint8_t *p1 = sP1 + (size_t)b * C::P1_BYTES;
#pragma unroll
            for (int ii = 0; ii < (S * NC16 + EPI_GROUPS - 1) / EPI_GROUPS; ++ii) {
                const int item = ii * EPI_GROUPS + half;
                if (item >= S * NC16) break;
                const int j = item / NC16, c16 = item % NC16;
                const int r = j * 8 + ph;                                   // P1 row
                const int Y = ty * C::TH - 1 + r, X = tx * C::TW - 2 + pw;   // image position of this lane's value
                const bool inside = (Y >= 0) && (Y < H) && (X >= 0) && (X < W);
                float v[16];
                combine(tbase + j * C::ACC_STRIDE + c16 * 16, v);
                uint32_t o[8];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 sc = *reinterpret_cast<const float4 *>(s_scale1 + c16 * 16 + 4 * g);
                    const float4 sh = *reinterpret_cast<const float4 *>(s_shift1 + c16 * 16 + 4 * g);
                    __nv_bfloat162 p0 = __hmax2(__floats2bfloat162_rn(fmaf(v[4 * g], sc.x, sh.x), fmaf(v[4 * g + 1], sc.y, sh.y)), zero2);
                    __nv_bfloat162 p1v = __hmax2(__floats2bfloat162_rn(fmaf(v[4 * g + 2], sc.z, sh.z), fmaf(v[4 * g + 3], sc.w, sh.w)), zero2);
                    o[2 * g] = inside ? *reinterpret_cast<uint32_t *>(&p0) : 0u;      // conv2's SAME padding
                    o[2 * g + 1] = inside ? *reinterpret_cast<uint32_t *>(&p1v) : 0u;
                }
                if (pw >= 1 && pw <= 14) {
                    uint8_t *d = p1 + ((size_t)(c16 * 2) * C::PH + r) * (C::PW * 16) + (pw - 1) * 16;
                    *reinterpret_cast<uint4 *>(d) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4 *>(d + C::PH * C::PW * 16) = make_uint4(o[4], o[5], o[6], o[7]);
                }
            }
            tc::tc_fence_before();
            tc::fence_proxy_async();                  // generic-proxy writes of P1 -> visible to the UMMA reads
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&p1_full[b]);
        };
        // epilogue 2 of tile number i: accumulators -> output
        auto epi2 = [&](int i) {
            const int b = i & 1;
            const int t = (int)blockIdx.x + i * (int)gridDim.x;
            const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, n = t / tpi;
            tc::mbar_wait(&acc2_full[b], (i >> 1) & 1);
            tc::tc_fence_after();
            const uint32_t tbase = tmem_base + lane_addr + b * C::ACC_COLS;
            const int x = tx * C::TW + pw - 1;
            const bool xok = (pw >= 1) && (pw <= C::TW) && (x < W);
            if constexpr (EPI != EPI_HEAD) {
                const size_t plane = (size_t)H * W * 8;
#pragma unroll
                for (int ii = 0; ii < (S * NC16 + EPI_GROUPS - 1) / EPI_GROUPS; ++ii) {
                    const int item = ii * EPI_GROUPS + half;
                    if (item >= S * NC16) break;
                    const int j = item / NC16, c16 = item % NC16;
                    const int r = j * 8 + ph, y = ty * C::TH + r;
                    const bool valid = xok && (r < C::TH) && (y < H);
                    float v[16];
                    combine(tbase + j * C::ACC_STRIDE + c16 * 16, v);
                    uint32_t o[8];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 sc = *reinterpret_cast<const float4 *>(s_scale2 + c16 * 16 + 4 * g);
                        const float4 sh = *reinterpret_cast<const float4 *>(s_shift2 + c16 * 16 + 4 * g);
                        __nv_bfloat162 p0 = __hmax2(__floats2bfloat162_rn(fmaf(v[4 * g], sc.x, sh.x), fmaf(v[4 * g + 1], sc.y, sh.y)), zero2);
                        __nv_bfloat162 p1v = __hmax2(__floats2bfloat162_rn(fmaf(v[4 * g + 2], sc.z, sh.z), fmaf(v[4 * g + 3], sc.w, sh.w)), zero2);
                        o[2 * g] = *reinterpret_cast<uint32_t *>(&p0);
                        o[2 * g + 1] = *reinterpret_cast<uint32_t *>(&p1v);
                    }
                    if (valid) {
                        bf16 *p = out + ((((size_t)n * CBo + c16 * 2) * H + y) * W + x) * 8;
                        *reinterpret_cast<uint4 *>(p) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4 *>(p + plane) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                    if (EPI == EPI_POOL) {
                        // rows (y, y+1) sit 16 lanes apart (TH and the tile origin are even), columns (x, x+1) in
                        // lanes (pw, pw+1) with pw odd
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const uint32_t m = bf162_max(o[e], __shfl_xor_sync(0xffffffffu, o[e], 16));
                            o[e] = bf162_max(m, __shfl_down_sync(0xffffffffu, m, 1));
                        }
                        if (valid && lane < 16 && (pw & 1)) {
                            const int Hp = H >> 1, Wp = W >> 1;
                            bf16 *p = out_pool + ((((size_t)n * CBo + c16 * 2) * Hp + (y >> 1)) * Wp + (x >> 1)) * 8;
                            *reinterpret_cast<uint4 *>(p) = make_uint4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<uint4 *>(p + (size_t)Hp * Wp * 8) = make_uint4(o[4], o[5], o[6], o[7]);
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int j = half; j < S; j += EPI_GROUPS) {
                    const int r = j * 8 + ph, y = ty * C::TH + r;
                    const bool valid = xok && (r < C::TH) && (y < H);
                    float hl[HK > 0 ? HK : 1];
#pragma unroll
                    for (int k = 0; k < HK; ++k) hl[k] = 0.0f;
#pragma unroll 1
                    for (int c16 = 0; c16 < NC16; ++c16) {
                        float v[16], f[16];
                        combine(tbase + j * C::ACC_STRIDE + c16 * 16, v);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 sc = *reinterpret_cast<const float4 *>(s_scale2 + c16 * 16 + 4 * g);
                            const float4 sh = *reinterpret_cast<const float4 *>(s_shift2 + c16 * 16 + 4 * g);
                            // the head consumes the activation as it would have been stored (bf16)
                            const float2 t0 = __bfloat1622float2(__floats2bfloat162_rn(
                                fmaxf(fmaf(v[4 * g], sc.x, sh.x), 0.0f), fmaxf(fmaf(v[4 * g + 1], sc.y, sh.y), 0.0f)));
                            const float2 t1 = __bfloat1622float2(__floats2bfloat162_rn(
                                fmaxf(fmaf(v[4 * g + 2], sc.z, sh.z), 0.0f), fmaxf(fmaf(v[4 * g + 3], sc.w, sh.w), 0.0f)));
                            f[4 * g] = t0.x; f[4 * g + 1] = t0.y; f[4 * g + 2] = t1.x; f[4 * g + 3] = t1.y;
                        }
#pragma unroll
                        for (int k = 0; k < HK; ++k) {
                            const float4 *wk = reinterpret_cast<const float4 *>(s_head + k * COUT + c16 * 16);
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const float4 w4 = wk[g];
                                hl[k] = fmaf(f[4 * g], w4.x, hl[k]);
                                hl[k] = fmaf(f[4 * g + 1], w4.y, hl[k]);
                                hl[k] = fmaf(f[4 * g + 2], w4.z, hl[k]);
                                hl[k] = fmaf(f[4 * g + 3], w4.w, hl[k]);
                            }
                        }
                    }
                    const size_t p = ((size_t)n * H + y) * W + x;
                    int best = 0;
                    float m = -INFINITY;
#pragma unroll
                    for (int k = 0; k < HK; ++k) {
                        hl[k] += s_head[COUT * HK + k];
                        if (hl[k] > m) { m = hl[k]; best = k; }
                    }
                    if (valid && head.mask) head.mask[p] = (uint8_t)best;
                    if (valid && head.logits) {
#pragma unroll
                        for (int k = 0; k < HK; ++k) head.logits[p * HK + k] = hl[k];
                    }
                    if (valid && head.probs) {
                        float sum = 0.0f;
#pragma unroll
                        for (int k = 0; k < HK; ++k) { hl[k] = expf(hl[k] - m); sum += hl[k]; }
#pragma unroll
                        for (int k = 0; k < HK; ++k) head.probs[p * HK + k] = hl[k] / sum;
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_free[b]);
        };
        // same order as the MMA warp: stage 1 of tile i+1 is drained before stage 2 of tile i
        for (int it = 0; it < my_tiles; ++it) {
            epi1(it);
            if (it > 0) epi2(it - 1);
        }
        epi2(my_tiles - 1);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------ level-0 layers on the QUAD (space-to-depth) layout
// At Cout = 16 a 128 x 16 x 16 MMA is bound by its A-operand fetch (4 KB of shared memory, ~39 clk) and does
// 1/4 of the math a 64-wide instruction does in ~48 clk.  The level-0 tensors (16 channels at full
// resolution) are therefore kept as QUAD tensors: a 2x2 block of pixels becomes ONE pixel of a half-resolution
// image with 4 x 16 = 64 channels, channel = q*16 + c with q = 2*(y & 1) + (x & 1):
//     quad[n][(q*16 + c) / 8][y / 2][x / 2][c % 8]
// On that layout
//   * a 3x3 conv 16 -> 16 is a 3x3 conv 64 -> 64 at half resolution whose weight blocks are zero for most
//     (tap, input-parity) pairs: input parity (iy, ix) only reaches the taps ty in {iy ? -1,0 : 0,+1},
//     tx likewise -- FOUR 128 x 64 x 16 MMAs per 16 input channels (one k-step = one input parity of one
//     source) instead of 4 x 9 = 36 MMAs 128 x 16 x 16 for the same 512 output pixels;
//   * the 2x2 stride-2 up-conv 32 -> 16 is a plain 1x1 conv 32 -> 64 (output parity = kernel tap), no
//     scattered store;
//   * the fused 2x2 max-pool is a max over the four 16-column groups of ONE accumulator row (no shuffles);
//   * the fused head runs once per 16-column group.
// Same persistent warp-specialised structure as conv_tc_kernel (TMA producer / MMA issuer / 8 epilogue
// warps, smem ring + double-buffered TMEM); the weights of all k-steps stay resident in shared memory and a
// ring stage holds KPS k-steps fetched by one TMA box.  NTAP = 4: quad 3x3 conv; NTAP = 1: 1x1 conv.
template <int S, int NTAP>
struct QCfg {
    static constexpr int COUT = 64;
    static constexpr int TH = 16 * S;
    static constexpr int PW = NTAP == 4 ? 10 : 8, PH = NTAP == 4 ? TH + 2 : TH;
    static constexpr int A_BYTES = 2 * PH * PW * 16;            // one k-step (16 channels) of the patch
    static constexpr int W_BYTES = NTAP * 2 * COUT * 16;        // one k-step of the weights
    static constexpr int ACC_COLS = S * COUT;
    static constexpr uint32_t LBO_A = PH * PW * 16, SBO_A = PW * 16;
    static constexpr uint32_t LBO_B = COUT * 16, SBO_B = 128;
    static_assert(A_BYTES % 128 == 0, "TMA destination alignment");
};

constexpr int QD_MAX_STAGES = 8;

// The layer's folded epilogue (16 channels, the same for the four parities) and the fused head's weights, passed BY
// VALUE: the epilogue warps read them as constant-bank operands of their FFMAs.  ncu on the head variant with these
// tables in shared memory: 20 % short-scoreboard + 10 % MIO-throttle stalls (16 LDS.128 per 16-column work item).
struct QdEpi {
    float sc[16], sh[16];
    float hw[4][16], hb[4];        // head: [class][channel], bias[class]
};

template <int S, int NTAP, int KPS, int EPI, int HK>
__global__ void __launch_bounds__(TC_THREADS, 2)
conv_qd_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               int ks0, int ks1, const bf16 *__restrict__ wts, const QdEpi ep, bf16 *__restrict__ out,
               bf16 *__restrict__ out_pool,
               HeadArgs head, int nimg, int H, int W, int relu, int nstages, int wres, long long *phase_dbg)
{
#ifdef SQ_XC_PHASE_DIAG
#define QD_T0() const long long t_ = phase_dbg ? clock64() : 0
#define QD_ACC(var) if (phase_dbg) var += clock64() - t_
    long long d_wait0 = 0, d_wait1 = 0, d_work = 0;
#else
#define QD_T0() do { } while (0)
#define QD_ACC(var) do { } while (0)
#endif
    // H, W: the quad image (half the level-0 frame).  scale / shift: 64 entries (the layer's 16, four times).
    using C = QCfg<S, NTAP>;
    constexpr int COUT = C::COUT, NBUF = 2, A_STAGE = KPS * C::A_BYTES;
    constexpr int TMEM_COLS = (NBUF * C::ACC_COLS <= 128) ? 128 : (NBUF * C::ACC_COLS <= 256) ? 256 : 512;
    static_assert(2 * NBUF * C::ACC_COLS <= 512, "two co-resident CTAs must fit in TMEM");
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full_bar[QD_MAX_STAGES], empty_bar[QD_MAX_STAGES], tfull_bar[2], tempty_bar[2], w_bar;
    __shared__ uint32_t tmem_base_sh;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (W + 7) >> 3, tiles_y = (H + C::TH - 1) / C::TH;
    const int ntiles = nimg * tiles_x * tiles_y;
    const int ksteps = ks0 + ks1;
    uint8_t *ring = smem + ((wres + 127) & ~127);

    if (threadIdx.x == 0) {
        for (int i = 0; i < nstages; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull_bar[i], 1); tc::mbar_init(&tempty_bar[i], 4 * EPI_GROUPS); }
        tc::mbar_init(&w_bar, 1);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&mapA0);
        tc::tma_prefetch_desc(&mapA1);
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_base_sh, TMEM_COLS); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(&w_bar, (uint32_t)wres);
            for (int q = 0; q < ksteps; ++q)
                tc::bulk_load(smem + (size_t)q * C::W_BYTES, wts + (size_t)q * (C::W_BYTES / 2), C::W_BYTES, &w_bar);
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, n = t / (tiles_x * tiles_y);
                const int x0 = tx * 8 - (NTAP == 4 ? 1 : 0), y0 = ty * C::TH - (NTAP == 4 ? 1 : 0);
                for (int ks = 0; ks < ksteps; ks += KPS) {
                    { QD_T0(); tc::mbar_wait(&empty_bar[stage], phase ^ 1); QD_ACC(d_wait0); }
                    QD_T0();
                    tc::mbar_arrive_expect_tx(&full_bar[stage], A_STAGE);
                    uint8_t *sA = ring + (size_t)stage * A_STAGE;
                    if (ks < ks0) tc::tma_load_5d(sA, &mapA0, &full_bar[stage], x0 * 8, y0, ks * 2, 0, n);
                    else          tc::tma_load_5d(sA, &mapA1, &full_bar[stage], x0 * 8, y0, (ks - ks0) * 2, 0, n);
                    QD_ACC(d_work);
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
            }
#ifdef SQ_XC_PHASE_DIAG
            if (phase_dbg) { phase_dbg[blockIdx.x * 8 + 0] = d_wait0; phase_dbg[blockIdx.x * 8 + 1] = d_work; }
#endif
        }
    } else if (warp == 1) {
        // ======================================================= MMA issuer
        const uint32_t idesc = tc::instr_desc_bf16(128, COUT);
        const uint32_t a_hi = ((C::SBO_A >> 4) & 0x3FFFu) | (1u << 14);
        const uint32_t b_hi = ((C::SBO_B >> 4) & 0x3FFFu) | (1u << 14);
        const uint32_t w_lo = ((tc::smem_u32(smem) >> 4) & 0x3FFFu) | (((C::LBO_B >> 4) & 0x3FFFu) << 16);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        tc::mbar_wait(&w_bar, 0);
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int buf = it % NBUF;
            { QD_T0(); tc::mbar_wait(&tempty_bar[buf], ((it / NBUF) & 1) ^ 1); QD_ACC(d_wait0); }
            tc::tc_fence_after();
            for (int ks = 0; ks < ksteps; ks += KPS) {
                { QD_T0(); tc::mbar_wait(&full_bar[stage], phase); QD_ACC(d_wait1); }
                tc::tc_fence_after();
                QD_T0();
                if (tc::elect_one()) {
                    const uint32_t a_base = tc::smem_u32(ring + (size_t)stage * A_STAGE);
                    const uint32_t a_lo = ((a_base >> 4) & 0x3FFFu) | (((C::LBO_A >> 4) & 0x3FFFu) << 16);
                    const uint32_t d0 = tmem_base + buf * C::ACC_COLS;
#pragma unroll
                    for (int kk = 0; kk < KPS; ++kk) {
                        const int k = ks + kk;
                        // input parity of this k-step: the first patch row / column its two taps start from
                        const int iy = (k >> 1) & 1, ix = k & 1;
                        const uint32_t b_k = w_lo + (uint32_t)(k * (C::W_BYTES >> 4));
#pragma unroll
                        for (int j = 0; j < S; ++j) {
                            if (NTAP == 4) {
#pragma unroll
                                for (int tp = 0; tp < 4; ++tp) {
                                    const uint32_t a_off = (uint32_t)(kk * (C::A_BYTES >> 4) +
                                                                      (j * 16 + (tp >> 1) + 1 - iy) * C::PW + (tp & 1) + 1 - ix);
                                    tc::umma_bf16_parts(d0 + j * COUT, a_lo + a_off, a_hi, b_k + tp * 2 * COUT, b_hi, idesc,
                                                        (k > 0 || tp > 0) ? 1u : 0u);
                                }
                            } else {
                                tc::umma_bf16_parts(d0 + j * COUT, a_lo + (uint32_t)(kk * (C::A_BYTES >> 4) + j * 16 * C::PW), a_hi,
                                                    b_k, b_hi, idesc, k > 0 ? 1u : 0u);
                            }
                        }
                    }
                    tc::umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                QD_ACC(d_work);
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
            if (tc::elect_one()) tc::umma_commit(&tfull_bar[buf]);
            __syncwarp();
        }
#ifdef SQ_XC_PHASE_DIAG
        if (phase_dbg && lane == 0) {
            phase_dbg[blockIdx.x * 8 + 2] = d_wait0; phase_dbg[blockIdx.x * 8 + 3] = d_wait1; phase_dbg[blockIdx.x * 8 + 4] = d_work;
        }
#endif
    } else {
        // ========================================================= epilogue
        // Work items are (sub-tile j, 16-column group c16), taken two at a time: both TMEM loads in flight before the
        // single wait; ReLU on the packed bf16 pairs; tile coordinates advance incrementally (no division).
        const int q4 = warp & 3, half = (warp - 2) >> 2;
        const int r = q4 * 32 + lane, ph = r >> 3, pw = r & 7;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        const size_t plane = (size_t)H * W * 8;
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.0f, 0.0f);
        const int tpi = tiles_x * tiles_y, gstep = (int)gridDim.x;
        int tx = (int)blockIdx.x % tiles_x, ty = ((int)blockIdx.x / tiles_x) % tiles_y, n = (int)blockIdx.x / tpi;
        const int dtx = gstep % tiles_x, dty = (gstep / tiles_x) % tiles_y, dn = gstep / tpi;
        // convert one 16-column group: scale / shift (+ ReLU) -> 8 packed bf16 pairs
        auto convert = [&](const uint32_t *v, uint32_t *o) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                __nv_bfloat162 p = __floats2bfloat162_rn(fmaf(__uint_as_float(v[2 * e]), ep.sc[2 * e], ep.sh[2 * e]),
                                                         fmaf(__uint_as_float(v[2 * e + 1]), ep.sc[2 * e + 1], ep.sh[2 * e + 1]));
                if (relu) p = __hmax2(p, zero2);
                o[e] = *reinterpret_cast<uint32_t *>(&p);
            }
        };
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int buf = it % NBUF;
            const int x = tx * 8 + pw;
            { QD_T0(); tc::mbar_wait(&tfull_bar[buf], (it / NBUF) & 1); QD_ACC(d_wait0); }
            tc::tc_fence_after();
            QD_T0();
            const uint32_t tbase = tmem_base + lane_addr + buf * C::ACC_COLS;
            uint32_t mx[8];                          // POOL: running max over the four parities
            // this group's pairs of items: STORE (j, c16) with (4j + c16) % 2 == half in order; POOL all four groups
            // of sub-tile j (j % 2 == half); HEAD the two column parities of row parity `half`, every sub-tile
#pragma unroll 1
            for (int i = 0; i < S; ++i) {
                int j, ca, cb;
                if (EPI == EPI_STORE) { j = i; ca = half; cb = half + 2; }
                else if (EPI == EPI_POOL) { j = 2 * (i >> 1) + half; ca = 2 * (i & 1); cb = ca + 1; }
                else { j = i; ca = 2 * half; cb = ca + 1; }
                const int y = ty * C::TH + j * 16 + ph;
                const bool valid = (y < H) && (x < W);
                uint32_t va[16], vb[16];
                tc::tmem_ld16(tbase + j * COUT + ca * 16, va);
                tc::tmem_ld16(tbase + j * COUT + cb * 16, vb);
                tc::tmem_ld_wait();
                uint32_t oa[8], ob[8];
                convert(va, oa);
                convert(vb, ob);
                if constexpr (EPI == EPI_HEAD) {
                    // the head consumes the activation as it would have been stored (bf16); 16-column group c16 is
                    // level-0 pixel (2y + oy, 2x + ox), (oy, ox) = (c16 >> 1, c16 & 1): ca / cb are ox = 0 / 1 of row `half`
                    int best[2];
                    const size_t p0 = ((size_t)n * (2 * H) + 2 * y + half) * (size_t)(2 * W) + 2 * x;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const uint32_t *o = u ? ob : oa;
                        float hl[HK > 0 ? HK : 1];
#pragma unroll
                        for (int k = 0; k < HK; ++k) hl[k] = 0.0f;
                        float f[16];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float2 t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&o[e]));
                            f[2 * e] = t2.x;
                            f[2 * e + 1] = t2.y;
                        }
#pragma unroll
                        for (int k = 0; k < HK; ++k) {
#pragma unroll
                            for (int c = 0; c < 16; ++c) hl[k] = fmaf(f[c], ep.hw[k][c], hl[k]);
                        }
                        int bk = 0;
                        float m = -INFINITY;
#pragma unroll
                        for (int k = 0; k < HK; ++k) {
                            hl[k] += ep.hb[k];
                            if (hl[k] > m) { m = hl[k]; bk = k; }
                        }
                        best[u] = bk;
                        if (valid && head.logits) {
#pragma unroll
                            for (int k = 0; k < HK; ++k) head.logits[(p0 + u) * HK + k] = hl[k];
                        }
                        if (valid && head.probs) {
                            float sum = 0.0f;
#pragma unroll
                            for (int k = 0; k < HK; ++k) { hl[k] = expf(hl[k] - m); sum += hl[k]; }
#pragma unroll
                            for (int k = 0; k < HK; ++k) head.probs[(p0 + u) * HK + k] = hl[k] / sum;
                        }
                    }
                    // both pixels of this lane's row (ox = 0, 1) leave in one 2-byte store
                    if (valid && head.mask) *reinterpret_cast<uint16_t *>(head.mask + p0) = (uint16_t)(best[0] | (best[1] << 8));
                } else {
                    if (valid) {
                        bf16 *p = out + ((((size_t)n * 8) * H + y) * W + x) * 8;
                        *reinterpret_cast<uint4 *>(p + (2 * ca) * plane) = make_uint4(oa[0], oa[1], oa[2], oa[3]);
                        *reinterpret_cast<uint4 *>(p + (2 * ca + 1) * plane) = make_uint4(oa[4], oa[5], oa[6], oa[7]);
                        *reinterpret_cast<uint4 *>(p + (2 * cb) * plane) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
                        *reinterpret_cast<uint4 *>(p + (2 * cb + 1) * plane) = make_uint4(ob[4], ob[5], ob[6], ob[7]);
                    }
                    if constexpr (EPI == EPI_POOL) {
                        // the 2x2 max-pooled tensor of the level-0 activation (16 channels at quad resolution): the max
                        // over the four 16-column groups of the accumulator row, two per pass
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const uint32_t m2 = bf162_max(oa[e], ob[e]);
                            mx[e] = (i & 1) ? bf162_max(mx[e], m2) : m2;
                        }
                        if ((i & 1) && valid) {
                            bf16 *pp = out_pool + ((((size_t)n * 2) * H + y) * W + x) * 8;
                            *reinterpret_cast<uint4 *>(pp) = make_uint4(mx[0], mx[1], mx[2], mx[3]);
                            *reinterpret_cast<uint4 *>(pp + plane) = make_uint4(mx[4], mx[5], mx[6], mx[7]);
                        }
                    }
                }
            }
            tx += dtx;
            if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
            ty += dty;
            if (ty >= tiles_y) { ty -= tiles_y; ++n; }
            n += dn;
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty_bar[buf]);
            QD_ACC(d_work);
        }
#ifdef SQ_XC_PHASE_DIAG
        if (phase_dbg && warp == 2 && lane == 0) { phase_dbg[blockIdx.x * 8 + 5] = d_wait0; phase_dbg[blockIdx.x * 8 + 6] = d_work; }
#endif
    }
#undef QD_T0
#undef QD_ACC
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// --------------------- down0/conv1 -> down0/conv2 (+ 2x2 max-pool) as ONE launch on the quad layout
// The 16-channel output of the first conv (64 of the 108 B/px the two layers move) never goes to HBM.
// Per tile of 8 x 32 quad pixels (16 x 64 frame pixels), one CTA per SM, 22 warps:
//   warp 0       TMA: the fp32 halo of the tile (24 x 70 frame pixels, zero fill outside the frame)
//   warps 2-5    builders: im2col of the first conv for the 10 x 34 quad pixels of conv2's halo patch --
//                row p = the 4 x 4 frame-pixel window around quad pixel p, 16 bf16 = ONE k-step (A1)
//   warp 1       MMA: conv1 = 3 MMAs 128 x 64 x 16 (A1 x B1, B1 = the 3x3 kernel scattered over
//                (output parity, window position)) -> acc1; conv2 = 32 quad MMAs on the patch P -> acc2
//   warps 6-21   epilogue 1: acc1 -> scale/shift/ReLU -> bf16 -> P in shared memory, laid out exactly like a
//                TMA-fetched quad patch (zeros outside the frame = conv2's SAME padding);
//                epilogue 2: acc2 -> scale/shift/ReLU -> bf16 -> quad tensor + in-thread 2x2 max-pool
// The MMA warp issues conv1 of tile i+1 before conv2 of tile i and the epilogue warps mirror that order, so
// the tensor pipe runs conv2 of a tile while its successor's patch is being written (A1, the fp32 halo, P and
// acc2 are double-buffered).  conv2's output columns are ordered (co / 8, parity, co % 8): one 32-column
// TMEM load holds all four parities of 8 channels, i.e. four 16-byte stores and the pooled vector.
struct QF {
    static constexpr int TH = 32, PW = 10, PH = 34, PROWS = PW * PH;      // conv2's halo patch (quad pixels)
    static constexpr int MT1 = 3, A1_ROWS = 128 * MT1;
    static constexpr int A1_BYTES = A1_ROWS * 32;                         // [2][384][8] bf16
    static constexpr int RAW_W = 32, RAW_H = 2 * PH + 2;       // 128-byte rows; the box starts one column early (16-byte aligned)
    static constexpr int RAW_BYTES = (RAW_W * RAW_H * 4 + 127) / 128 * 128;
    static constexpr int P_BYTES = 8 * PROWS * 16;
    static constexpr int W2_BYTES = 4 * 4 * 2 * 64 * 16, B1_BYTES = 2 * 64 * 16;
    static constexpr int OFF_B1 = 0, OFF_W2 = B1_BYTES, OFF_A1 = OFF_W2 + W2_BYTES, OFF_RAW = OFF_A1 + 2 * A1_BYTES,
                         OFF_P = OFF_RAW + 2 * RAW_BYTES, SMEM = OFF_P + 2 * P_BYTES;
    // two MMA-issuing warps (one per conv2 sub-tile; a single thread issues one 128x64x16 MMA per ~70 clk, the pipe takes 48)
    static constexpr int MMA_WARPS = 2, EPI_WARPS = 16, BLD_WARPS = 4, FIRST_BLD = 1 + MMA_WARPS,
                         FIRST_EPI = FIRST_BLD + BLD_WARPS, THREADS = 32 * (FIRST_EPI + EPI_WARPS);
    static_assert(OFF_A1 % 128 == 0 && OFF_RAW % 128 == 0 && OFF_P % 128 == 0 && P_BYTES % 128 == 0, "alignment");
};

// the two layers' folded epilogues, passed BY VALUE: the epilogue warps read them as constant-bank operands of
// their FFMAs (shared-memory loads queue behind the tensor core's operand fetches while MMAs are running)
struct QfEpi {
    float sc1[16], sh1[16], sc2[16], sh2[16];
};

// RAW16: the frames are raw uint16 camera values and ImageNorm (pipeline.py:338-356: (x - mean) / std per frame, in
// float32 exactly as sq_image_norm does) is applied by the builders while they pack the im2col -- the float32 copy
// of the frames never exists.  stats: (mean, std) per frame.
template <bool RAW16>
__global__ void __launch_bounds__(QF::THREADS, 1)
conv_qf_kernel(const __grid_constant__ CUtensorMap mapIn, const bf16 *__restrict__ wts, const QfEpi ep,
               const float2 *__restrict__ stats,
               bf16 *__restrict__ out, bf16 *__restrict__ out_pool, int nimg, int H, int W, long long *phase_dbg)
{
#ifdef SQ_XC_PHASE_DIAG
#define QF_W(i, stmt) do { const long long t_ = phase_dbg ? clock64() : 0; stmt; if (phase_dbg) dacc[i] += clock64() - t_; } while (0)
#define QF_T0() const long long t0_ = phase_dbg ? clock64() : 0
#define QF_T1(i) if (phase_dbg) dacc[i] += clock64() - t0_
    long long dacc[6] = {0, 0, 0, 0, 0, 0};
    const long long t_kernel0 = clock64();
#else
#define QF_W(i, stmt) do { stmt; } while (0)
#define QF_T0() do { } while (0)
#define QF_T1(i) do { } while (0)
#endif
    // H, W: the quad image.  wts: B1 then W2 (QF::B1_BYTES + QF::W2_BYTES), scale / shift: the layers' own 16.
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t w_bar, raw_full[2], raw_empty[2], a1_full[2], a1_empty[2], acc1_full, acc1_empty,
        p_full[2], p_empty[2], acc2_full[2], acc2_empty[2];
    __shared__ uint32_t tmem_base_sh;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (W + 7) >> 3, tiles_y = (H + QF::TH - 1) / QF::TH;
    const int ntiles = nimg * tiles_x * tiles_y;
    const int nt = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (threadIdx.x == 0) {
        tc::mbar_init(&w_bar, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&raw_full[i], 1); tc::mbar_init(&raw_empty[i], QF::BLD_WARPS);
            tc::mbar_init(&a1_full[i], QF::BLD_WARPS); tc::mbar_init(&a1_empty[i], QF::MMA_WARPS);
            tc::mbar_init(&p_full[i], QF::EPI_WARPS); tc::mbar_init(&p_empty[i], QF::MMA_WARPS);
            tc::mbar_init(&acc2_full[i], QF::MMA_WARPS); tc::mbar_init(&acc2_empty[i], QF::EPI_WARPS);
        }
        tc::mbar_init(&acc1_full, QF::MMA_WARPS); tc::mbar_init(&acc1_empty, QF::EPI_WARPS);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&mapIn);
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_base_sh, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;
    constexpr uint32_t ACC1 = 0, ACC2 = 256;                 // TMEM columns: acc1 3 x 64, acc2 2 x (2 x 64)

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(&w_bar, QF::B1_BYTES + QF::W2_BYTES);
            tc::bulk_load(smem + QF::OFF_B1, wts, QF::B1_BYTES, &w_bar);
            for (int q = 0; q < 4; ++q)
                tc::bulk_load(smem + QF::OFF_W2 + q * (QF::W2_BYTES / 4), wts + (QF::B1_BYTES + q * (QF::W2_BYTES / 4)) / 2,
                              QF::W2_BYTES / 4, &w_bar);
            for (int it = 0; it < nt; ++it) {
                const int t = blockIdx.x + it * gridDim.x, b = it & 1;
                const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, n = t / (tiles_x * tiles_y);
                QF_W(0, tc::mbar_wait(&raw_empty[b], ((it >> 1) & 1) ^ 1));
                tc::mbar_arrive_expect_tx(&raw_full[b], QF::RAW_W * QF::RAW_H * (RAW16 ? 2 : 4));
                // frame pixels (16 tx - 3 ..., 64 ty - 3 ...): the 4 x 4 windows of quad pixels (8 tx - 1 ..., 32 ty - 1 ...);
                // the box starts at column 16 tx - 4 (uint16: 16 tx - 8) so that every row of it is 16-byte aligned
                // in global memory
                tc::tma_load_3d(smem + QF::OFF_RAW + b * QF::RAW_BYTES, &mapIn, &raw_full[b], 16 * tx - (RAW16 ? 8 : 4), 64 * ty - 3, n);
            }
        }
    } else if (warp <= QF::MMA_WARPS) {
        // ======================================================= MMA issuers: warp 1 owns conv2 sub-tile 0, warp 2 sub-tile 1
        const int jw = warp - 1;
        const uint32_t idesc = tc::instr_desc_bf16(128, 64);
        const uint32_t sbase = tc::smem_u32(smem);
        const uint32_t hi128 = ((128u >> 4) & 0x3FFFu) | (1u << 14);                 // SBO = 128 (A1, B1, W2)
        const uint32_t hiP = (((uint32_t)(QF::PW * 16) >> 4) & 0x3FFFu) | (1u << 14);   // SBO of the patch = one row
        const uint32_t b1_lo = (((sbase + QF::OFF_B1) >> 4) & 0x3FFFu) | (((64u * 16u) >> 4) << 16);
        const uint32_t w2_lo = (((sbase + QF::OFF_W2) >> 4) & 0x3FFFu) | (((64u * 16u) >> 4) << 16);
        tc::mbar_wait(&w_bar, 0);
        for (int it = 0; it <= nt; ++it) {
            if (it < nt) {
                const int b = it & 1;
                QF_W(0, tc::mbar_wait(&a1_full[b], (it >> 1) & 1));
                QF_W(1, tc::mbar_wait(&acc1_empty, (it & 1) ^ 1));
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t a1_lo = (((sbase + QF::OFF_A1 + b * QF::A1_BYTES) >> 4) & 0x3FFFu) |
                                           ((((uint32_t)QF::A1_ROWS * 16u) >> 4) << 16);
                    for (int m = jw; m < QF::MT1; m += QF::MMA_WARPS)           // M-tiles 0, 2 / 1
                        tc::umma_bf16_parts(tmem_base + ACC1 + m * 64, a1_lo + m * (2048 >> 4), hi128, b1_lo, hi128, idesc, 0u);
                    tc::umma_commit(&a1_empty[b]);
                    tc::umma_commit(&acc1_full);
                }
                __syncwarp();
            }
            if (it >= 1) {
                const int j2 = it - 1, b = j2 & 1;
                QF_W(2, tc::mbar_wait(&p_full[b], (j2 >> 1) & 1));
                QF_W(3, tc::mbar_wait(&acc2_empty[b], ((j2 >> 1) & 1) ^ 1));
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t p_lo = (((sbase + QF::OFF_P + b * QF::P_BYTES) >> 4) & 0x3FFFu) |
                                          ((((uint32_t)QF::PROWS * 16u) >> 4) << 16);
                    const uint32_t d0 = tmem_base + ACC2 + b * 128;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int iy = k >> 1, ix = k & 1;
#pragma unroll
                        for (int tp = 0; tp < 4; ++tp) {
                            const uint32_t a_off = (uint32_t)(k * (2 * QF::PROWS) + (jw * 16 + (tp >> 1) + 1 - iy) * QF::PW +
                                                              (tp & 1) + 1 - ix);
                            tc::umma_bf16_parts(d0 + jw * 64, p_lo + a_off, hiP, w2_lo + (uint32_t)((k * 4 + tp) * (2048 >> 4)),
                                                hi128, idesc, (k > 0 || tp > 0) ? 1u : 0u);
                        }
                    }
                    tc::umma_commit(&p_empty[b]);
                    tc::umma_commit(&acc2_full[b]);
                }
                __syncwarp();
            }
        }
    } else if (warp < QF::FIRST_EPI) {
        // ===================================================== im2col builders
        const int bt = threadIdx.x - 32 * QF::FIRST_BLD;     // 0 .. 127
        int roff[3];                                         // this thread's rows p = bt + 128 k: window origin in the halo
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int p = bt + 128 * k, py = p / QF::PW, px = p - py * QF::PW;
            roff[k] = (2 * py) * QF::RAW_W + 2 * px;
        }
        for (int it = 0; it < nt; ++it) {
            const int b = it & 1;
            QF_W(0, tc::mbar_wait(&raw_full[b], (it >> 1) & 1));
            QF_W(1, tc::mbar_wait(&a1_empty[b], ((it >> 1) & 1) ^ 1));
            uint8_t *a1 = smem + QF::OFF_A1 + b * QF::A1_BYTES;
            if constexpr (!RAW16) {
                const float *raw = reinterpret_cast<const float *>(smem + QF::OFF_RAW + b * QF::RAW_BYTES);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int p = bt + 128 * k;
                    if (p < QF::PROWS) {
                        uint32_t o[8];
#pragma unroll
                        for (int wy = 0; wy < 4; ++wy) {
                            // window columns 2 px + 1 .. 2 px + 4 of the staged row (the box starts one column early)
                            const float2 c0 = *reinterpret_cast<const float2 *>(raw + roff[k] + wy * QF::RAW_W);
                            const float2 c1 = *reinterpret_cast<const float2 *>(raw + roff[k] + wy * QF::RAW_W + 2);
                            const float2 c2 = *reinterpret_cast<const float2 *>(raw + roff[k] + wy * QF::RAW_W + 4);
                            o[2 * wy] = pack_bf16(c0.y, c1.x);
                            o[2 * wy + 1] = pack_bf16(c1.y, c2.x);
                        }
                        *reinterpret_cast<uint4 *>(a1 + p * 16) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4 *>(a1 + QF::A1_ROWS * 16 + p * 16) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
            } else {
                // raw uint16 values, window columns 2 px + 5 .. 2 px + 8 of the staged row (the box starts five columns
                // early); normalised in float32; positions outside the frame are the conv's zero padding, not (0 - mean) / std
                const uint32_t *raw = reinterpret_cast<const uint32_t *>(smem + QF::OFF_RAW + b * QF::RAW_BYTES);
                const int t = blockIdx.x + it * gridDim.x;
                const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, n = t / (tiles_x * tiles_y);
                if (stats == nullptr) {
                    // the 16-bit values ARE bf16 (normalised frames rounded once by sq_image_norm_u16_to_bf16; TMA's
                    // zero fill outside the frame is the conv's zero padding): the im2col is a byte shuffle
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int p = bt + 128 * k;
                        if (p < QF::PROWS) {
                            const int py = p / QF::PW, px = p - py * QF::PW;
                            uint32_t o[8];
#pragma unroll
                            for (int wy = 0; wy < 4; ++wy) {
                                const uint32_t *rw = raw + (2 * py + wy) * (QF::RAW_W / 2) + px + 2;
                                const uint32_t w0 = rw[0], w1 = rw[1], w2 = rw[2];
                                o[2 * wy] = __byte_perm(w0, w1, 0x5432);          // columns 2px+5, 2px+6
                                o[2 * wy + 1] = __byte_perm(w1, w2, 0x5432);      // columns 2px+7, 2px+8
                            }
                            *reinterpret_cast<uint4 *>(a1 + p * 16) = make_uint4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<uint4 *>(a1 + QF::A1_ROWS * 16 + p * 16) = make_uint4(o[4], o[5], o[6], o[7]);
                        }
                    }
                    goto built;
                }
                const float2 ms = __ldg(stats + n);
                const int X0 = 16 * tx - 3, Y0 = 64 * ty - 3, H0 = 2 * H, W0 = 2 * W;
                const bool border = X0 < 0 || Y0 < 0 || X0 + 2 * QF::PW + 2 > W0 || Y0 + 2 * QF::PH + 2 > H0;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int p = bt + 128 * k;
                    if (p < QF::PROWS) {
                        const int py = p / QF::PW, px = p - py * QF::PW;
                        uint32_t o[8];
#pragma unroll
                        for (int wy = 0; wy < 4; ++wy) {
                            const uint32_t *rw = raw + (2 * py + wy) * (QF::RAW_W / 2) + px + 2;
                            const uint32_t w0 = rw[0], w1 = rw[1], w2 = rw[2];
                            float f[4] = {(float)(w0 >> 16), (float)(w1 & 0xffffu), (float)(w1 >> 16), (float)(w2 & 0xffffu)};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                f[e] = __fdiv_rn(__fsub_rn(f[e], ms.x), ms.y);
                                if (border) {
                                    const int Y = Y0 + 2 * py + wy, X = X0 + 2 * px + e;
                                    if (Y < 0 || Y >= H0 || X < 0 || X >= W0) f[e] = 0.0f;
                                }
                            }
                            o[2 * wy] = pack_bf16(f[0], f[1]);
                            o[2 * wy + 1] = pack_bf16(f[2], f[3]);
                        }
                        *reinterpret_cast<uint4 *>(a1 + p * 16) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4 *>(a1 + QF::A1_ROWS * 16 + p * 16) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
        built:
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tc::mbar_arrive(&a1_full[b]); tc::mbar_arrive(&raw_empty[b]); }
        }
    } else {
        // ========================================================= epilogue
        const int g = (warp - QF::FIRST_EPI) >> 2, q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        int ppy[3], ppx[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) { const int p = m * 128 + r; ppy[m] = p / QF::PW; ppx[m] = p - ppy[m] * QF::PW; }
        // epilogue 2: this group owns sub-tile j2 and channels c8*8 .. c8*8+7 (all four parities)
        const int j2 = g >> 1, c8 = g & 1;
        float sc2[8], sh2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { sc2[e] = c8 ? ep.sc2[8 + e] : ep.sc2[e]; sh2[e] = c8 ? ep.sh2[8 + e] : ep.sh2[e]; }
        const size_t plane = (size_t)H * W * 8;
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.0f, 0.0f);
        // tile coordinates advance incrementally (no division in the loop); (tx, ty, n): the tile of epilogue 1,
        // (tx2, ty2, n2): the tile of epilogue 2 (one iteration behind)
        const int tpi = tiles_x * tiles_y, gstep = (int)gridDim.x;
        int tx = (int)blockIdx.x % tiles_x, ty = ((int)blockIdx.x / tiles_x) % tiles_y, n = (int)blockIdx.x / tpi;
        const int dtx = gstep % tiles_x, dty = (gstep / tiles_x) % tiles_y, dn = gstep / tpi;
        int tx2 = 0, ty2 = 0, n2 = 0;
        for (int it = 0; it <= nt; ++it) {
            if (it < nt) {
                // ---- epilogue 1: conv1's accumulators -> the bf16 halo patch of conv2
                const int b = it & 1;
                const int x0 = tx * 8 - 1, y0 = ty * QF::TH - 1;
                QF_W(0, tc::mbar_wait(&acc1_full, it & 1));
                QF_W(1, tc::mbar_wait(&p_empty[b], ((it >> 1) & 1) ^ 1));
                tc::tc_fence_after();
                QF_T0();
                uint8_t *P = smem + QF::OFF_P + b * QF::P_BYTES + (2 * g) * (QF::PROWS * 16) + r * 16;
                // tiles that touch the frame border zero the patch pixels outside the frame (warp-uniform test)
                const bool border = x0 < 0 || y0 < 0 || x0 + QF::PW > W || y0 + QF::PH > H;
                uint32_t v[3][16];
                tc::tmem_ld16(tmem_base + lane_addr + ACC1 + g * 16, v[0]);
                tc::tmem_ld16(tmem_base + lane_addr + ACC1 + 64 + g * 16, v[1]);
                if (q4 != 3) tc::tmem_ld16(tmem_base + lane_addr + ACC1 + 128 + g * 16, v[2]);   // rows 352.. lie beyond the patch
                tc::tmem_ld_wait();
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    if (m == 2 && q4 == 3) continue;
                    uint32_t o[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {        // scale / shift: constant-bank operands (kernel parameters), no LDS
                        __nv_bfloat162 h = __hmax2(__floats2bfloat162_rn(fmaf(__uint_as_float(v[m][2 * q]), ep.sc1[2 * q], ep.sh1[2 * q]),
                                                                         fmaf(__uint_as_float(v[m][2 * q + 1]), ep.sc1[2 * q + 1], ep.sh1[2 * q + 1])),
                                                   zero2);      // ReLU after the rounding (max commutes with it)
                        o[q] = *reinterpret_cast<uint32_t *>(&h);
                    }
                    if (border) {
                        const int Y = y0 + ppy[m], X = x0 + ppx[m];
                        if (!(Y >= 0 && Y < H && X >= 0 && X < W)) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) o[e] = 0u;
                        }
                    }
                    if (m < 2 || r < QF::PROWS - 256) {
                        *reinterpret_cast<uint4 *>(P + m * 2048) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4 *>(P + QF::PROWS * 16 + m * 2048) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
                tc::fence_proxy_async();                     // generic-proxy writes -> visible to the UMMA reads
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) { tc::mbar_arrive(&p_full[b]); tc::mbar_arrive(&acc1_empty); }
                QF_T1(3);
            }
            if (it >= 1) {
                // ---- epilogue 2: conv2's accumulators -> quad tensor + pooled tensor
                const int i2 = it - 1, b = i2 & 1;
                const int y = ty2 * QF::TH + j2 * 16 + (r >> 3), x = tx2 * 8 + (r & 7);
#ifdef SQ_XC_PHASE_DIAG
                const bool valid = (y < H) && (x < W) && !(phase_dbg && phase_dbg[gridDim.x * 32] == 2);   // experiment: no stores
#else
                const bool valid = (y < H) && (x < W);
#endif
                QF_W(2, tc::mbar_wait(&acc2_full[b], (i2 >> 1) & 1));
                tc::tc_fence_after();
                QF_T0();
                uint32_t v[32];
                tc::tmem_ld32(tmem_base + lane_addr + ACC2 + b * 128 + j2 * 64 + c8 * 32, v);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&acc2_empty[b]);          // the accumulators are in registers
                if (valid) {
                    // ONE address per work item; the four parity planes are 2 * plane apart (a 64-bit add each -- written
                    // as `po + 2 * q * plane` inside per-store `if (valid)` blocks the compiler rebuilt the whole 64-bit
                    // product chain, ~30 integer instructions, in front of every 16-byte store)
                    char *pq = reinterpret_cast<char *>(out + ((((size_t)n2 * 8 + c8) * H + y) * W + x) * 8);
                    const size_t step2 = 2 * plane * sizeof(bf16);
                    uint32_t mx[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            __nv_bfloat162 h = __hmax2(__floats2bfloat162_rn(fmaf(__uint_as_float(v[q * 8 + 2 * e]), sc2[2 * e], sh2[2 * e]),
                                                                             fmaf(__uint_as_float(v[q * 8 + 2 * e + 1]), sc2[2 * e + 1], sh2[2 * e + 1])),
                                                       zero2);
                            o[e] = *reinterpret_cast<uint32_t *>(&h);
                        }
                        *reinterpret_cast<uint4 *>(pq) = make_uint4(o[0], o[1], o[2], o[3]);
                        pq += step2;
#pragma unroll
                        for (int e = 0; e < 4; ++e) mx[e] = (q == 0) ? o[e] : bf162_max(mx[e], o[e]);
                    }
                    *reinterpret_cast<uint4 *>(out_pool + ((((size_t)n2 * 2 + c8) * H + y) * W + x) * 8) =
                        make_uint4(mx[0], mx[1], mx[2], mx[3]);
                }
                QF_T1(4);
            }
            tx2 = tx; ty2 = ty; n2 = n;
            tx += dtx;
            if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
            ty += dty;
            if (ty >= tiles_y) { ty -= tiles_y; ++n; }
            n += dn;
        }
    }
#ifdef SQ_XC_PHASE_DIAG
    // per role (producer, MMA, first builder warp, first epilogue warp): clocks spent in each wait
    if (phase_dbg && lane == 0 && (warp <= 1 || warp == QF::FIRST_BLD || warp == QF::FIRST_EPI)) {
        const int role = warp <= 1 ? warp : (warp == QF::FIRST_BLD ? 2 : 3);
        dacc[5] = clock64() - t_kernel0;                 // the role's whole lifetime, in clock64 ticks
        for (int i = 0; i < 6; ++i) phase_dbg[(blockIdx.x * 4 + role) * 8 + i] = dacc[i];
    }
#endif
#undef QF_W
#undef QF_T0
#undef QF_T1
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// ------------- up0/upscale -> up0/conv1 (concat bridge) as ONE launch on the quad layout
// The up-sampled tensor (32 of the 80 + 64 B/px the two layers move) never goes to HBM, and the result is
// bit-identical to the two-launch quad path (same MMAs, same bf16 rounding of the intermediate).  Structure of
// conv_qf_kernel with two differences: "conv1" is the 2x2 stride-2 up-conv = a 1x1 conv (cin -> 4 x 16) over the
// 10 x 34 half-resolution pixels of conv's halo patch, whose A operand is the TMA-fetched patch of the level-1
// tensor itself (no builders); and the conv reads a second source, the skip tensor, through an 8-stage ring of
// one-k-step TMA boxes (two tiles of prefetch).  One CTA per SM:
//   warp 0       TMA: level-1 patch (A1) + 4 skip k-steps per tile
//   warp 1       MMA: up-conv 3 M-tiles x ksu k-steps -> acc1; conv: 4 skip k-steps + 4 P k-steps, 64 MMAs -> acc2
//   warps 2-17   epilogue 1: acc1 + bias -> bf16 -> P (zeros outside the frame); epilogue 2: acc2 -> ReLU -> quad tensor
// Schedule (two balanced phases per tile): the tensor pipe runs  [P half of tile i-1] [up-conv + skip half of tile i]
// while the epilogue warps run  [epilogue 1 of tile i, under the skip half of tile i] [epilogue 2 of tile i-1, under
// the P half of tile i]; P is single-buffered (free once the P half of the previous tile has retired), the
// conv accumulators are double-buffered.
struct QU {
    static constexpr int TH = 32, PW = 10, PH = 34, PROWS = PW * PH;
    static constexpr int KSTEP_BYTES = 2 * PROWS * 16;                    // 16 channels of a patch
    static constexpr int A1_BYTES = (2 * KSTEP_BYTES + 1024 + 127) / 128 * 128;   // up to 32 input channels (+ M-tile over-read)
    static constexpr int P_BYTES = 4 * KSTEP_BYTES;
    static constexpr int WU_BYTES = 2 * 2 * 64 * 16, W_BYTES = 8 * 4 * 2 * 64 * 16;
    static constexpr int NRING = 8;
    static constexpr int OFF_WU = 0, OFF_W = WU_BYTES, OFF_A1 = OFF_W + W_BYTES, OFF_P = OFF_A1 + A1_BYTES,
                         OFF_RING = OFF_P + P_BYTES, SMEM = OFF_RING + NRING * KSTEP_BYTES;
    // two MMA-issuing warps (one per sub-tile): a single thread sustains one 128x64x16 MMA per ~70 clk here, the pipe 48
    // warp 0: weights + level-1 patches; warps 1-2: MMA; warp 3: the skip ring's own TMA producer (decoupled from the
    // single-buffered level-1 patch, so the ring really runs two tiles ahead); warps 4-19: epilogue
    static constexpr int MMA_WARPS = 2, RING_WARP = 1 + MMA_WARPS, FIRST_EPI = RING_WARP + 1, EPI_WARPS = 16,
                         THREADS = 32 * (FIRST_EPI + EPI_WARPS);
    static_assert(OFF_A1 % 128 == 0 && OFF_P % 128 == 0 && OFF_RING % 128 == 0 && KSTEP_BYTES % 128 == 0, "alignment");
    static_assert(SMEM + 1024 <= 232448, "shared memory");
};

__global__ void __launch_bounds__(QU::THREADS, 1)
conv_qu_kernel(const __grid_constant__ CUtensorMap mapCur, const __grid_constant__ CUtensorMap mapSkip,
               const bf16 *__restrict__ wup, const bf16 *__restrict__ wts, const QfEpi ep, int ksu,
               bf16 *__restrict__ out, int nimg, int H, int W, long long *phase_dbg)
{
#ifdef SQ_XC_PHASE_DIAG
#define QU_W(i, stmt) do { const long long t_ = phase_dbg ? clock64() : 0; stmt; if (phase_dbg) dacc[i] += clock64() - t_; } while (0)
    long long dacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_kernel0 = clock64();
#else
#define QU_W(i, stmt) do { stmt; } while (0)
#endif
    // H, W: the quad image.  wup: the up-conv as a 1x1 conv (ksu k-steps x 2 KB), wts: conv1's quad weights with
    // permuted output columns (8 k-steps: up-sampled source parities 0-3, skip parities 0-3).  ep.sc1 / sh1: the
    // up-conv's epilogue, ep.sc2 / sh2: the conv's.
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t w_bar, a1_full, a1_empty, acc1_full, acc1_empty, p_full, p_empty,
        ring_full[QU::NRING], ring_empty[QU::NRING], acc2_full[2], acc2_empty[2];
    __shared__ uint32_t tmem_base_sh;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (W + 7) >> 3, tiles_y = (H + QU::TH - 1) / QU::TH;
    const int ntiles = nimg * tiles_x * tiles_y;
    const int nt = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (threadIdx.x == 0) {
        tc::mbar_init(&w_bar, 1);
        tc::mbar_init(&a1_full, 1); tc::mbar_init(&a1_empty, QU::MMA_WARPS);
        tc::mbar_init(&acc1_full, QU::MMA_WARPS); tc::mbar_init(&acc1_empty, QU::EPI_WARPS);
        tc::mbar_init(&p_full, QU::EPI_WARPS); tc::mbar_init(&p_empty, QU::MMA_WARPS);
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc2_full[i], QU::MMA_WARPS); tc::mbar_init(&acc2_empty[i], QU::EPI_WARPS); }
        for (int i = 0; i < QU::NRING; ++i) { tc::mbar_init(&ring_full[i], 1); tc::mbar_init(&ring_empty[i], QU::MMA_WARPS); }
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&mapCur);
        tc::tma_prefetch_desc(&mapSkip);
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_base_sh, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;
    constexpr uint32_t ACC1 = 0, ACC2 = 256;                 // TMEM columns: acc1 3 x 64, acc2 2 x (2 x 64)

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(&w_bar, (uint32_t)(ksu * 2048 + QU::W_BYTES));
            tc::bulk_load(smem + QU::OFF_WU, wup, (uint32_t)(ksu * 2048), &w_bar);
            for (int q = 0; q < 8; ++q)
                tc::bulk_load(smem + QU::OFF_W + q * 8192, wts + q * 4096, 8192, &w_bar);
            for (int it = 0; it < nt; ++it) {
                const int t = blockIdx.x + it * gridDim.x;
                const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, n = t / (tiles_x * tiles_y);
                const int x0 = tx * 8 - 1, y0 = ty * QU::TH - 1;
                QU_W(0, tc::mbar_wait(&a1_empty, (it & 1) ^ 1));
                tc::mbar_arrive_expect_tx(&a1_full, (uint32_t)(ksu * QU::KSTEP_BYTES));
                tc::tma_load_5d(smem + QU::OFF_A1, &mapCur, &a1_full, x0 * 8, y0, 0, 0, n);
            }
        }
    } else if (warp == QU::RING_WARP) {
        // ===================================================== TMA producer of the skip ring
        if (lane == 0) {
            int rs = 0;
            uint32_t rphase = 0;
            for (int it = 0; it < nt; ++it) {
                const int t = blockIdx.x + it * gridDim.x;
                const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, n = t / (tiles_x * tiles_y);
                const int x0 = tx * 8 - 1, y0 = ty * QU::TH - 1;
                for (int ks = 0; ks < 4; ++ks) {
                    QU_W(1, tc::mbar_wait(&ring_empty[rs], rphase ^ 1));
                    tc::mbar_arrive_expect_tx(&ring_full[rs], QU::KSTEP_BYTES);
                    tc::tma_load_5d(smem + QU::OFF_RING + rs * QU::KSTEP_BYTES, &mapSkip, &ring_full[rs], x0 * 8, y0, ks * 2, 0, n);
                    if (++rs == QU::NRING) { rs = 0; rphase ^= 1; }
                }
            }
        }
    } else if (warp <= QU::MMA_WARPS) {
        // ======================================================= MMA issuers: warp 1 owns sub-tile 0, warp 2 sub-tile 1
        const int jw = warp - 1;
        const uint32_t idesc = tc::instr_desc_bf16(128, 64);
        const uint32_t sbase = tc::smem_u32(smem);
        const uint32_t hi128 = ((128u >> 4) & 0x3FFFu) | (1u << 14);
        const uint32_t hiP = (((uint32_t)(QU::PW * 16) >> 4) & 0x3FFFu) | (1u << 14);
        const uint32_t lboP = (((uint32_t)QU::PROWS * 16u) >> 4) << 16;
        const uint32_t wu_lo = (((sbase + QU::OFF_WU) >> 4) & 0x3FFFu) | (((64u * 16u) >> 4) << 16);
        const uint32_t w_lo = (((sbase + QU::OFF_W) >> 4) & 0x3FFFu) | (((64u * 16u) >> 4) << 16);
        const uint32_t a1_lo = (((sbase + QU::OFF_A1) >> 4) & 0x3FFFu) | lboP;
        int rs = 0;
        uint32_t rphase = 0;
        tc::mbar_wait(&w_bar, 0);
        for (int it = 0; it <= nt; ++it) {
            if (it >= 1) {
                // ---- P half of tile it-1: the up-sampled patch written by epilogue 1
                const int i2 = it - 1, b = i2 & 1;
                const uint32_t d0 = tmem_base + ACC2 + b * 128;
                QU_W(0, tc::mbar_wait(&p_full, i2 & 1));
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t p_lo = (((sbase + QU::OFF_P) >> 4) & 0x3FFFu) | lboP;
                    const uint32_t pj = p_lo + (uint32_t)(jw * 16 * QU::PW), dj = d0 + jw * 64;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int iy = k >> 1, ix = k & 1;
#pragma unroll
                        for (int tp = 0; tp < 4; ++tp)
                            tc::umma_bf16_parts(dj, pj + (uint32_t)(k * (QU::KSTEP_BYTES >> 4) + ((tp >> 1) + 1 - iy) * QU::PW + (tp & 1) + 1 - ix),
                                                hiP, w_lo + (uint32_t)((k * 4 + tp) * 128), hi128, idesc, 1u);
                    }
                    tc::umma_commit(&p_empty);
                    tc::umma_commit(&acc2_full[b]);
                }
                __syncwarp();
            }
            if (it < nt) {
                // ---- up-conv of tile it, then the skip half of its conv (prefetched through the ring)
                const int b = it & 1;
                QU_W(1, tc::mbar_wait(&a1_full, it & 1));
                QU_W(2, tc::mbar_wait(&acc1_empty, (it & 1) ^ 1));
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    for (int m = jw; m < 3; m += QU::MMA_WARPS)              // M-tiles 0, 2 / 1
                        for (int k = 0; k < ksu; ++k)
                            tc::umma_bf16_parts(tmem_base + ACC1 + m * 64, a1_lo + (uint32_t)(k * (QU::KSTEP_BYTES >> 4) + m * 128), hi128,
                                                wu_lo + (uint32_t)(k * 128), hi128, idesc, k > 0 ? 1u : 0u);
                    tc::umma_commit(&a1_empty);
                    tc::umma_commit(&acc1_full);
                }
                __syncwarp();
                QU_W(3, tc::mbar_wait(&acc2_empty[b], ((it >> 1) & 1) ^ 1));
                tc::tc_fence_after();
                const uint32_t d0 = tmem_base + ACC2 + b * 128;
                for (int ks = 0; ks < 4; ++ks) {
                    QU_W(4, tc::mbar_wait(&ring_full[rs], rphase));
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
                        const uint32_t s_lo = (((sbase + QU::OFF_RING + rs * QU::KSTEP_BYTES) >> 4) & 0x3FFFu) | lboP;
                        const int iy = ks >> 1, ix = ks & 1;
#pragma unroll
                        for (int tp = 0; tp < 4; ++tp)
                            tc::umma_bf16_parts(d0 + jw * 64, s_lo + (uint32_t)((jw * 16 + (tp >> 1) + 1 - iy) * QU::PW + (tp & 1) + 1 - ix), hiP,
                                                w_lo + (uint32_t)(((4 + ks) * 4 + tp) * 128), hi128, idesc, (ks > 0 || tp > 0) ? 1u : 0u);
                        tc::umma_commit(&ring_empty[rs]);
                    }
                    __syncwarp();
                    if (++rs == QU::NRING) { rs = 0; rphase ^= 1; }
                }
            }
        }
    } else {
        // ========================================================= epilogue
        const int g = (warp - QU::FIRST_EPI) >> 2, q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        int ppy[3], ppx[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) { const int p = m * 128 + r; ppy[m] = p / QU::PW; ppx[m] = p - ppy[m] * QU::PW; }
        const int j2 = g >> 1, c8 = g & 1;
        float sc2[8], sh2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { sc2[e] = c8 ? ep.sc2[8 + e] : ep.sc2[e]; sh2[e] = c8 ? ep.sh2[8 + e] : ep.sh2[e]; }
        const size_t plane = (size_t)H * W * 8;
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.0f, 0.0f);
        const int tpi = tiles_x * tiles_y, gstep = (int)gridDim.x;
        int tx = (int)blockIdx.x % tiles_x, ty = ((int)blockIdx.x / tiles_x) % tiles_y, n = (int)blockIdx.x / tpi;
        const int dtx = gstep % tiles_x, dty = (gstep / tiles_x) % tiles_y, dn = gstep / tpi;
        int tx2 = 0, ty2 = 0, n2 = 0;
        for (int it = 0; it <= nt; ++it) {
            if (it < nt) {
                // ---- epilogue 1: the up-conv's accumulators (+ bias, no ReLU) -> the bf16 halo patch of the conv
                const int x0 = tx * 8 - 1, y0 = ty * QU::TH - 1;
                QU_W(0, tc::mbar_wait(&acc1_full, it & 1));
                QU_W(1, tc::mbar_wait(&p_empty, (it & 1) ^ 1));
                tc::tc_fence_after();
                const long long te1_ = clock64();
                uint8_t *P = smem + QU::OFF_P + (2 * g) * (QU::PROWS * 16) + r * 16;
                const bool border = x0 < 0 || y0 < 0 || x0 + QU::PW > W || y0 + QU::PH > H;
                uint32_t v[3][16];
                tc::tmem_ld16(tmem_base + lane_addr + ACC1 + g * 16, v[0]);
                tc::tmem_ld16(tmem_base + lane_addr + ACC1 + 64 + g * 16, v[1]);
                if (q4 != 3) tc::tmem_ld16(tmem_base + lane_addr + ACC1 + 128 + g * 16, v[2]);   // rows 352.. lie beyond the patch
                tc::tmem_ld_wait();
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    if (m == 2 && q4 == 3) continue;
                    uint32_t o[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        o[q] = pack_bf16(fmaf(__uint_as_float(v[m][2 * q]), ep.sc1[2 * q], ep.sh1[2 * q]),
                                         fmaf(__uint_as_float(v[m][2 * q + 1]), ep.sc1[2 * q + 1], ep.sh1[2 * q + 1]));
                    if (border) {
                        const int Y = y0 + ppy[m], X = x0 + ppx[m];
                        if (!(Y >= 0 && Y < H && X >= 0 && X < W)) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) o[e] = 0u;
                        }
                    }
                    if (m < 2 || r < QU::PROWS - 256) {
                        *reinterpret_cast<uint4 *>(P + m * 2048) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4 *>(P + QU::PROWS * 16 + m * 2048) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
                tc::fence_proxy_async();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) { tc::mbar_arrive(&p_full); tc::mbar_arrive(&acc1_empty); }
#ifdef SQ_XC_PHASE_DIAG
                dacc[3] += clock64() - te1_;
#endif
            }
            if (it >= 1) {
                // ---- epilogue 2: the conv's accumulators -> ReLU -> quad tensor
                const int i2 = it - 1, b = i2 & 1;
                const int y = ty2 * QU::TH + j2 * 16 + (r >> 3), x = tx2 * 8 + (r & 7);
                const bool valid = (y < H) && (x < W);
                QU_W(2, tc::mbar_wait(&acc2_full[b], (i2 >> 1) & 1));
                tc::tc_fence_after();
                const long long te2_ = clock64();
                uint32_t v[32];
                tc::tmem_ld32(tmem_base + lane_addr + ACC2 + b * 128 + j2 * 64 + c8 * 32, v);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&acc2_empty[b]);
                if (valid) {
                    // one address per work item, the parity planes a 64-bit add apart (see conv_qf_kernel's epilogue 2)
                    char *pq = reinterpret_cast<char *>(out + ((((size_t)n2 * 8 + c8) * H + y) * W + x) * 8);
                    const size_t step2 = 2 * plane * sizeof(bf16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            __nv_bfloat162 h = __hmax2(__floats2bfloat162_rn(fmaf(__uint_as_float(v[q * 8 + 2 * e]), sc2[2 * e], sh2[2 * e]),
                                                                             fmaf(__uint_as_float(v[q * 8 + 2 * e + 1]), sc2[2 * e + 1], sh2[2 * e + 1])),
                                                       zero2);
                            o[e] = *reinterpret_cast<uint32_t *>(&h);
                        }
                        *reinterpret_cast<uint4 *>(pq) = make_uint4(o[0], o[1], o[2], o[3]);
                        pq += step2;
                    }
                }
#ifdef SQ_XC_PHASE_DIAG
                dacc[4] += clock64() - te2_;
#endif
            }
            tx2 = tx; ty2 = ty; n2 = n;
            tx += dtx;
            if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
            ty += dty;
            if (ty >= tiles_y) { ty -= tiles_y; ++n; }
            n += dn;
        }
    }
#ifdef SQ_XC_PHASE_DIAG
    if (phase_dbg && lane == 0 && (warp <= 1 || warp == QU::FIRST_EPI)) {           // producer, first MMA warp, first epilogue warp
        const int role = warp <= 1 ? warp : 2;
        dacc[7] = clock64() - t_kernel0;
        for (int i = 0; i < 8; ++i) phase_dbg[(blockIdx.x * 3 + role) * 8 + i] = dacc[i];
    }
#endif
#undef QU_W
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------- bandwidth-bound kernels
// First conv: fp32 NHWC input with few channels -> bf16 blocked.  K = 9*CIN (9 or 27) is far
// too small for a tcgen05 tile, and on CUDA cores the layer is instruction-bound (144 FMA per
// pixel at Cout=16) instead of HBM-bound.  A warp-level mma.sync (m16n8k16, bf16 x bf16 -> fp32)
// does the 16-pixel x 16-tap x 8-channel product in one instruction from registers: the A
// fragment is the im2col of 16 consecutive pixels (k = tap*CIN + c, zero beyond 9*CIN), built
// straight from the (L1-cached) input; B fragments (weights) live in registers.
// wf: [9*CIN][COUT] fp32 holding bf16-rounded weights.  HBM-bound: 4*CIN B in, 2*COUT B out per px.
__device__ __forceinline__ void mma_m16n8k16_bf16(float *c, const uint32_t *a, uint32_t b0, uint32_t b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int FIRST_TW = 128;              // pixels per block row
// rows per block: one warp per row, several passes (a multiple of 8).  Tall blocks amortise the
// per-block set-up (weight fragments, tap offsets: ~100 instructions per thread) and shrink the halo
// share of the staging; the cap keeps the staged tile below ~40 KB of static shared memory.
__host__ __device__ constexpr int first_rows(int cin, int kz)
{
    const int r = (10240 / (kz * ((FIRST_TW + 2) * cin + 2)) - 2) / 8 * 8;
    return r > 32 ? 32 : (r < 8 ? 8 : r);
}

// QOUT: the output is written as a quad tensor (conv_qd_kernel's layout; planar frames with even H and W).
template <int CIN, int COUT, int KZ, bool QOUT = false>
__global__ void __launch_bounds__(256)
first_conv_kernel(const float *__restrict__ in, const float *__restrict__ wf,
                  const float *__restrict__ scale, const float *__restrict__ shift,
                  bf16 *__restrict__ out, int nimg, int D, int H, int W)
{
    // KZ = 1: planar 3x3 (D = 1); KZ = 3: 3x3x3 on slice z of a volume, taps ordered (kz, ky, kx)
    constexpr int KTOT = 9 * KZ * CIN, KS = (KTOT + 15) / 16, NT = COUT / 8;
    constexpr int FIRST_ROWS = first_rows(CIN, KZ);
    constexpr int SROW = (FIRST_TW + 2) * CIN + 2;            // staged row pitch (floats)
    constexpr int SSLICE = (FIRST_ROWS + 2) * SROW;
    __shared__ float tile[KZ * SSLICE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int np = blockIdx.z, x0 = blockIdx.x * FIRST_TW, y0 = blockIdx.y * FIRST_ROWS;
    const int z = np % D;

    // stage the KZ x (ROWS+2) x (TW+2) input halo tile once (zero outside the image = SAME padding):
    // one warp per staged row, lanes along the row
    constexpr int ROWF = (FIRST_TW + 2) * CIN;
#pragma unroll 2
    for (int ri = warp; ri < KZ * (FIRST_ROWS + 2); ri += 8) {
        const int kz = ri / (FIRST_ROWS + 2), ry = ri % (FIRST_ROWS + 2);
        const int zz = z + kz - (KZ >> 1), yy = y0 + ry - 1, e0 = (x0 - 1) * CIN;
        const bool row_ok = zz >= 0 && zz < D && yy >= 0 && yy < H;
        const float *srow = in + ((size_t)(np + zz - z) * H + (row_ok ? yy : 0)) * W * CIN + e0;
        float *trow = tile + kz * SSLICE + ry * SROW;
        // every load of the row in flight before the first store (ncu: this kernel's top stall was the
        // L2 latency of the staging loads, issued one per iteration)
        constexpr int NLD = (ROWF + 31) / 32;
        float v[NLD];
#pragma unroll
        for (int k = 0; k < NLD; ++k) {
            const int rx = lane + 32 * k, e = e0 + rx;
            v[k] = (row_ok && rx < ROWF && e >= 0 && e < W * CIN) ? __ldg(srow + rx) : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < NLD; ++k) {
            const int rx = lane + 32 * k;
            if (rx < ROWF) trow[rx] = v[k];
        }
    }
    // B fragments: b0 = (k = 2t, 2t+1 ; n = g), b1 = (k = 2t+8, 2t+9 ; n = g); zero beyond KTOT
    uint32_t bw[KS][NT][2];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k0 = ks * 16 + 2 * t + 8 * h;
                const float w0 = k0 < KTOT ? wf[k0 * COUT + nt * 8 + g] : 0.0f;
                const float w1 = k0 + 1 < KTOT ? wf[(k0 + 1) * COUT + nt * 8 + g] : 0.0f;
                bw[ks][nt][h] = pack_bf16(w0, w1);
            }
    float sc[NT][2], sh[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) { sc[nt][e] = scale[nt * 8 + 2 * t + e]; sh[nt][e] = shift[nt * 8 + 2 * t + e]; }
    // this thread's A-fragment sources: k = ks*16 + 2t + (j&1) + 8*(j>>1) -> staged-tile address of
    // tap (dz,dy,dx), channel c for pixel g of this warp's row.  Taps beyond KTOT alias tap 0: their
    // weights are zero, so the (finite) value read there does not matter.
    int src[KS][4];                                      // offsets into `tile` (keeps the loads LDS)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = ks * 16 + 2 * t + (j & 1) + 8 * (j >> 1);
            const int tap = (k < KTOT) ? k / CIN : 0, c = (k < KTOT) ? k % CIN : 0;
            const int tz = tap / 9, tr = tap % 9;
            src[ks][j] = tz * SSLICE + (warp + tr / 3) * SROW + (g + tr % 3) * CIN + c;
        }
    __syncthreads();
    // quad output: pixel (y, x) -> channel blocks (2*(y&1) + (x&1))*NT + nt of the half-resolution image; x0, 16q
    // and 8r are even, so the column parity is g's and one fragment row spans 8 pixels = 4 quad columns
    const size_t plane = QOUT ? (size_t)(H >> 1) * (W >> 1) * 8 : (size_t)H * W * 8;
#pragma unroll 1
    for (int rr = 0; rr < FIRST_ROWS; rr += 8) {
        const int y = y0 + warp + rr;
        if (y >= H) return;
        const int roff = rr * SROW;                      // this row's offset in the staged tile
        bf16 *orow = QOUT ? out + ((((size_t)np * 4 + 2 * (y & 1) + (g & 1)) * NT * (H >> 1) + (y >> 1)) * (W >> 1) +
                                   ((x0 + g) >> 1)) * 8 + 2 * t
                          : out + (((size_t)np * NT * H + y) * W + x0 + g) * 8 + 2 * t;
#pragma unroll 2
        for (int q = 0; q < FIRST_TW / 16; ++q) {
            if (x0 + 16 * q >= W) break;
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[nt][e] = 0.0f;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                // rows g / g+8 of the fragment = pixels 16q+g and 16q+g+8
                uint32_t a[4];
                a[0] = pack_bf16(tile[src[ks][0] + roff + (16 * q) * CIN], tile[src[ks][1] + roff + (16 * q) * CIN]);
                a[1] = pack_bf16(tile[src[ks][0] + roff + (16 * q + 8) * CIN], tile[src[ks][1] + roff + (16 * q + 8) * CIN]);
                a[2] = pack_bf16(tile[src[ks][2] + roff + (16 * q) * CIN], tile[src[ks][3] + roff + (16 * q) * CIN]);
                a[3] = pack_bf16(tile[src[ks][2] + roff + (16 * q + 8) * CIN], tile[src[ks][3] + roff + (16 * q + 8) * CIN]);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma_m16n8k16_bf16(acc[nt], a, bw[ks][nt][0], bw[ks][nt][1]);
            }
            // C fragment: (row g, cols 2t,2t+1), (row g+8, cols 2t,2t+1) -> channel block nt, channels 2t,2t+1
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const float v0 = fmaxf(fmaf(acc[nt][2 * r], sc[nt][0], sh[nt][0]), 0.0f);
                    const float v1 = fmaxf(fmaf(acc[nt][2 * r + 1], sc[nt][1], sh[nt][1]), 0.0f);
                    if (x0 + 16 * q + g + 8 * r < W)
                        *reinterpret_cast<uint32_t *>(orow + nt * plane + (size_t)((16 * q + 8 * r) >> (QOUT ? 1 : 0)) * 8) = pack_bf16(v0, v1);
                }
        }
    }
}

// 2x2 max pool on the blocked layout: one thread per output 16-byte vector.
__global__ void maxpool_bf16_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out,
                                    long long nvec, int Ho, int Wo)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvec) return;
    const int x = (int)(i % Wo);
    const int y = (int)((i / Wo) % Ho);
    const long long plane = i / ((long long)Wo * Ho);          // n*CB + cb
    const uint4 *p = in + (plane * (2 * Ho) + 2 * y) * (2 * Wo) + 2 * x;
    const uint4 a = p[0], b = p[1], c = p[2 * Wo], d = p[2 * Wo + 1];
    uint4 m;
    m.x = bf162_max(bf162_max(a.x, b.x), bf162_max(c.x, d.x));
    m.y = bf162_max(bf162_max(a.y, b.y), bf162_max(c.y, d.y));
    m.z = bf162_max(bf162_max(a.z, b.z), bf162_max(c.z, d.z));
    m.w = bf162_max(bf162_max(a.w, b.w), bf162_max(c.w, d.w));
    out[i] = m;
}

// Volumes: first 3x3x3 conv, fp32 NDHWC input with few channels -> bf16 blocked [n*D+z][c/8][y][x][8].
// One thread per voxel; weights ([27*CIN][COUT] fp32, bf16-rounded) are broadcast from shared memory.
template <int COUT>
__global__ void __launch_bounds__(128)
first_conv3d_kernel(const float *__restrict__ in, const float *__restrict__ wf,
                    const float *__restrict__ scale, const float *__restrict__ shift,
                    bf16 *__restrict__ out, int nimg, int D, int H, int W, int CIN)
{
    extern __shared__ float sw3[];                      // [27*CIN][COUT] + scale[COUT] + shift[COUT]
    const int KT = 27 * CIN;
    for (int i = threadIdx.x; i < KT * COUT + 2 * COUT; i += blockDim.x)
        sw3[i] = i < KT * COUT ? wf[i] : (i < KT * COUT + COUT ? scale[i - KT * COUT] : shift[i - KT * COUT - COUT]);
    __syncthreads();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)nimg * D * H * W) return;
    const int x = (int)(p % W), y = (int)((p / W) % H), z = (int)((p / ((long long)W * H)) % D);
    const long long n = p / ((long long)W * H * D);
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.0f;
    for (int kz = 0; kz < 3; ++kz) {
        const int zz = z + kz - 1;
        if (zz < 0 || zz >= D) continue;
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            if (yy < 0 || yy >= H) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = x + kx - 1;
                if (xx < 0 || xx >= W) continue;
                const float *src = in + ((((size_t)n * D + zz) * H + yy) * W + xx) * CIN;
                for (int c = 0; c < CIN; ++c) {
                    const float v = __bfloat162float(__float2bfloat16_rn(__ldg(src + c)));
                    const float4 *w4 = reinterpret_cast<const float4 *>(sw3 + (size_t)(((kz * 3 + ky) * 3 + kx) * CIN + c) * COUT);
#pragma unroll
                    for (int q = 0; q < COUT / 4; ++q) {
                        const float4 w = w4[q];
                        acc[4 * q] = fmaf(v, w.x, acc[4 * q]);
                        acc[4 * q + 1] = fmaf(v, w.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(v, w.z, acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(v, w.w, acc[4 * q + 3]);
                    }
                }
            }
        }
    }
    const float *sc = sw3 + KT * COUT, *sh = sc + COUT;
    const long long np = n * D + z;
#pragma unroll
    for (int cb = 0; cb < COUT / 8; ++cb) {
        uint4 o;
        uint32_t *ow = reinterpret_cast<uint32_t *>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = cb * 8 + 2 * e;
            ow[e] = pack_bf16(fmaxf(fmaf(acc[c], sc[c], sh[c]), 0.0f), fmaxf(fmaf(acc[c + 1], sc[c + 1], sh[c + 1]), 0.0f));
        }
        *reinterpret_cast<uint4 *>(out + ((((size_t)np * (COUT / 8) + cb) * H + y) * W + x) * 8) = o;
    }
}

// Volumes: depth half of the 2x2x2 max pool.  in: [n][2*Do][plane] (the xy-pooled slices), out:
// [n][Do][plane]; plane = (c/8)*h*w 16-byte vectors.
__global__ void zpool_bf16_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, long long nvec,
                                  long long plane)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvec) return;
    const long long s = i / plane, r = i - s * plane;          // s = n*Do + zo
    const uint4 a = in[(2 * s) * plane + r], b = in[(2 * s + 1) * plane + r];
    uint4 m;
    m.x = bf162_max(a.x, b.x);
    m.y = bf162_max(a.y, b.y);
    m.z = bf162_max(a.z, b.z);
    m.w = bf162_max(a.w, b.w);
    out[i] = m;
}

__global__ void eltwise_bf16_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                                    long long n2, int op, uint32_t *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const uint32_t ua = a[i], ub = b[i];
    const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&ua));
    const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&ub));
    float r0, r1;
    if (op == SQ_BRIDGE_ADD) { r0 = x.x + y.x; r1 = x.y + y.y; }
    else if (op == SQ_BRIDGE_MUL) { r0 = x.x * y.x; r1 = x.y * y.y; }
    else { r0 = x.x - y.x; r1 = x.y - y.y; }
    out[i] = pack_bf16(r0, r1);
}

// 1x1 conv head + softmax + argmax.  wf: [C][K] fp32 (bf16-rounded values).
__global__ void head_bf16_kernel(const bf16 *__restrict__ in, const float *__restrict__ wf,
                                 const float *__restrict__ bias, int C, int K, long long npix_img,
                                 int nimg, float *__restrict__ logits, float *__restrict__ probs,
                                 uint8_t *__restrict__ mask)
{
    extern __shared__ float sw[];                       // [C][K] + bias[K]
    for (int i = threadIdx.x; i < C * K + K; i += blockDim.x) sw[i] = i < C * K ? wf[i] : bias[i - C * K];
    __syncthreads();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix_img * nimg) return;
    const long long n = p / npix_img, q = p % npix_img;
    float l[16];
    for (int k = 0; k < K; ++k) l[k] = 0.0f;
    for (int cb = 0; cb < C / 8; ++cb) {
        const uint4 u = *reinterpret_cast<const uint4 *>(in + (((size_t)n * (C / 8) + cb) * npix_img + q) * 8);
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w4[e]));
            const float *w0 = sw + (cb * 8 + 2 * e) * K;
            for (int k = 0; k < K; ++k) l[k] = fmaf(f.y, w0[K + k], fmaf(f.x, w0[k], l[k]));
        }
    }
    int best = 0;
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) {
        l[k] += sw[C * K + k];
        if (l[k] > m) { m = l[k]; best = k; }
    }
    if (logits) for (int k = 0; k < K; ++k) logits[p * K + k] = l[k];
    if (mask) mask[p] = (uint8_t)best;
    if (probs) {
        float s = 0.0f;
        for (int k = 0; k < K; ++k) { l[k] = expf(l[k] - m); s += l[k]; }
        for (int k = 0; k < K; ++k) probs[p * K + k] = l[k] / s;
    }
}

// ------------------------------------------------------------------ host side
uint16_t host_bf16(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
float host_bf16_round(float f)
{
    uint32_t u = (uint32_t)host_bf16(f) << 16;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

struct TcState {
    int sm_count = 148;
};

// SQ_GRID_DIV=d (experiments only): persistent grids of the tcgen05 kernels shrink to 1/d of the resident
// CTA slots, so that launches of two streams can be co-resident on every SM.
int grid_div()
{
    static int d = 0;
    if (d == 0) { const char *e = getenv("SQ_GRID_DIV"); d = e ? std::max(1, atoi(e)) : 1; }
    return d;
}

int dev_upload(sq_unet_s *u, const void *src, size_t bytes, void **dst)
{
    SQ_CUDA(cudaMalloc(dst, std::max<size_t>(bytes, 16)));
    u->dev_allocs.push_back(*dst);
    SQ_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return SQ_OK;
}

int make_map(CUtensorMap *m, const bf16 *ptr, int nimg, int D, int CB, int H, int W, int PW, int PH,
             int box_cb = 2, int box_z = 1)
{
    sq_encode_tiled_fn enc = sq_get_encode_tiled();
    SQ_REQUIRE(enc, SQ_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    // act[n][z][c/8][y][x][8] as a 5-D tensor (x*8, y, c/8, z, n); planar stacks have D = 1
    cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)CB, (cuuint64_t)D, (cuuint64_t)nimg};
    cuuint64_t strides[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)CB * H * W * 16,
                             (cuuint64_t)D * CB * H * W * 16};
    cuuint32_t box[5] = {(cuuint32_t)PW * 8, (cuuint32_t)PH, (cuuint32_t)box_cb, (cuuint32_t)box_z, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void *)ptr, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SQ_REQUIRE(r == CUDA_SUCCESS, SQ_ECUDA, "cuTensorMapEncodeTiled failed (%d) for (%d,%d,%d,%d,%d)",
               (int)r, nimg, D, CB, H, W);
    return SQ_OK;
}

// geometry of one launch: nimg volumes of D slices (planar: D = 1); KZ depth taps; up-conv output
// image index = in_image * out_mul + out_off
struct TcGeo {
    int nimg, D, H, W, KZ, out_mul, out_off;
    size_t w_off;                           // element offset into the layer's tensor-core weights
};

template <int COUT, int S, bool UP, int NBUF, int MINB, int EPI, int HK = 0, int NMMA = 1, bool CL = false>
int launch_tc(sq_unet_s *u, const SqLayer &L, const bf16 *in0, int cb0, const bf16 *in1, int cb1,
              bf16 *out, bf16 *out_pool, const HeadArgs &head, const TcGeo &g, int relu,
              cudaStream_t st)
{
    using C = Cfg<COUT, S, UP>;
    const int nimg = g.nimg, H = g.H, W = g.W;
    CUtensorMap m0, m1;
    SQ_TRY(make_map(&m0, in0, nimg, g.D, cb0, H, W, C::PW, C::PH));
    if (in1) SQ_TRY(make_map(&m1, in1, nimg, g.D, cb1, H, W, C::PW, C::PH));
    else m1 = m0;
    static_assert(MINB * NBUF * C::ACC_COLS <= 512, "co-resident CTAs must fit in TMEM");
    static_assert(NMMA == 1 || S % NMMA == 0, "sub-tiles must split evenly over the MMA warps");
    static_assert(!CL || (MINB == 1 && !UP), "cluster variant: one CTA per SM, 3x3 convs");
    int nstages = std::min(MAX_STAGES, ((MINB == 1 ? 200 : 216 / MINB) * 1024 - 2048) / C::STAGE_BYTES);
    nstages = std::max(nstages, 2);
    const size_t smem = (size_t)nstages * C::STAGE_BYTES + 1024;
    auto kern = conv_tc_kernel<COUT, S, UP, NBUF, MINB, EPI, HK, NMMA, CL>;
    static size_t attr_smem[64] = {0};          // the attribute is per device: one slot per device id
    size_t &have = attr_smem[u->h->device & 63];
    if (smem > have) {
        SQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
    const int threads = TC_THREADS + 32 * (NMMA - 1);
    const int tiles = nimg * g.D * ((W + 7) / 8) * ((H + C::TH - 1) / C::TH);
    TcEpi epc;                                   // read by the plain-store epilogue of the 32..128-channel convs only
    {
        const std::vector<float> &s0 = u->host[L.scope + "/_scale"].data, &t0 = u->host[L.scope + "/_shift"].data;
        for (int i = 0; i < 128; ++i) { epc.sc[i] = i < (int)s0.size() ? s0[i] : 1.0f; epc.sh[i] = i < (int)t0.size() ? t0[i] : 0.0f; }
    }
    if (CL) {
        // persistent clusters: as many CTA pairs as can be co-resident (GPCs with an odd SM count leave one out)
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3((unsigned)threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        static int max_clusters[64] = {0};
        int &mc = max_clusters[u->h->device & 63];
        if (mc == 0) {
            cfg.gridDim = dim3((unsigned)(2 * (u->h->sm_count / 2)));
            SQ_CUDA(cudaOccupancyMaxActiveClusters(&mc, kern, &cfg));
            SQ_REQUIRE(mc >= 1, SQ_ECUDA, "conv_tc: no 2-CTA cluster fits on this device");
        }
        const int grid = 2 * std::min(mc, (tiles + 1) / 2);
        cfg.gridDim = dim3((unsigned)grid);
        SQ_CUDA(cudaLaunchKernelEx(&cfg, kern, m0, m1, cb0 / 2, in1 ? cb1 / 2 : 0, (const bf16 *)L.w_tc + g.w_off,
                                   (const float *)L.scale, (const float *)L.shift, out, out_pool, head, nimg, H, W,
                                   relu, nstages, g.D, g.KZ, g.out_mul, g.out_off, FirstArgs{}, epc));
    } else {
        const int grid = std::min(tiles, MINB * u->h->sm_count / grid_div());
        kern<<<grid, threads, smem, st>>>(m0, m1, cb0 / 2, in1 ? cb1 / 2 : 0, (const bf16 *)L.w_tc + g.w_off, L.scale,
                                         L.shift, out, out_pool, head, nimg, H, W, relu, nstages, g.D, g.KZ,
                                         g.out_mul, g.out_off, FirstArgs{}, epc);
    }
    ++u->last_launches;
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

// Level-0 fusion: down0/conv1 (Cin = 1 -> 16) computed inside down0/conv2's producer (NBLD builder warps), the
// 16-channel intermediate never touches HBM.  Planar stacks, 16 -> 16 channels, fused 2x2 max-pool epilogue.
template <int NBLD>
int launch_tc_first_n(sq_unet_s *u, const SqLayer &L1, const SqLayer &L2, const float *in, bf16 *out, bf16 *out_pool,
                      const TcGeo &g, cudaStream_t st)
{
    constexpr int COUT = 16, S = 4, NBUF = 2, MINB = 2;
    using C = Cfg<COUT, S, false>;
    const int nimg = g.nimg, H = g.H, W = g.W;
    CUtensorMap m0;
    SQ_TRY(make_map(&m0, out, nimg, 1, 2, H, W, C::PW, C::PH));      // unused by the kernel (no TMA patch loads)
    int nstages = std::min(MAX_STAGES, ((216 / MINB) * 1024 - 2048) / C::STAGE_BYTES);
    nstages = std::max(nstages, 2);
    const size_t smem = (size_t)nstages * C::STAGE_BYTES + 1024;
    auto kern = conv_tc_kernel<COUT, S, false, NBUF, MINB, EPI_POOL, 0, 1, false, NBLD>;
    static size_t attr_smem[64] = {0};
    size_t &have = attr_smem[u->h->device & 63];
    const int threads = TC_THREADS + 32 * NBLD;
    if (smem > have) {
        SQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
    const int tiles = nimg * ((W + 7) / 8) * ((H + C::TH - 1) / C::TH);
    const int grid = std::min(tiles, MINB * u->h->sm_count);
    HeadArgs none = {};
    const FirstArgs fa = {in, (const float *)L1.w_tc, (const float *)L1.scale, (const float *)L1.shift};
    kern<<<grid, threads, smem, st>>>(m0, m0, 1, 0, (const bf16 *)L2.w_tc + g.w_off, L2.scale, L2.shift, out, out_pool,
                                     none, nimg, H, W, 1, nstages, 1, 1, g.out_mul, g.out_off, fa, TcEpi{});
    ++u->last_launches;
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

int launch_tc_first(sq_unet_s *u, const SqLayer &L1, const SqLayer &L2, const float *in, bf16 *out, bf16 *out_pool,
                    const TcGeo &g, cudaStream_t st)
{
    return launch_tc_first_n<4>(u, L1, L2, in, out, out_pool, g, st);
}

template <int COUT, int S, int MINB, int NMMA = 1, bool CL = false>
int conv3x3_epi(sq_unet_s *u, const SqLayer &L, const bf16 *in0, int c0, const bf16 *in1, int c1,
                bf16 *out, bf16 *out_pool, const HeadArgs *head, const TcGeo &g, cudaStream_t st)
{
    const HeadArgs none = {nullptr, 0, nullptr, nullptr, nullptr};
    if (head) {
        if constexpr (COUT <= 32) {
            switch (head->K) {
            case 2: return launch_tc<COUT, S, false, 2, MINB, EPI_HEAD, 2>(u, L, in0, c0 / 8, in1, c1 / 8, out, nullptr, *head, g, 1, st);
            case 3: return launch_tc<COUT, S, false, 2, MINB, EPI_HEAD, 3>(u, L, in0, c0 / 8, in1, c1 / 8, out, nullptr, *head, g, 1, st);
            case 4: return launch_tc<COUT, S, false, 2, MINB, EPI_HEAD, 4>(u, L, in0, c0 / 8, in1, c1 / 8, out, nullptr, *head, g, 1, st);
            }
        }
        SQ_REQUIRE(false, SQ_EUNSUPPORTED, "fused head supports filters[0] <= 32 and 2..4 classes");
    }
    if (out_pool) {
        if constexpr (COUT <= 128)
            return launch_tc<COUT, S, false, 2, MINB, EPI_POOL, 0, NMMA, CL>(u, L, in0, c0 / 8, in1, c1 / 8, out, out_pool,
                                                                         none, g, 1, st);
        else
            SQ_REQUIRE(false, SQ_EUNSUPPORTED, "fused pool needs filters <= 128");
    }
    return launch_tc<COUT, S, false, 2, MINB, EPI_STORE, 0, NMMA, CL>(u, L, in0, c0 / 8, in1, c1 / 8, out, nullptr, none,
                                                                  g, 1, st);
}

constexpr int SQ_NOT_APPLICABLE = 1;        // launch_xc: configuration does not fit, use the 9-tap kernel

template <int COUT, int S, int NBUF, int MINB, int EPI, int HK, bool PADACC, int KPS, int ZPS>
int launch_xc_k(sq_unet_s *u, const SqLayer &L, const bf16 *in0, int cb0, const bf16 *in1, int cb1,
                bf16 *out, bf16 *out_pool, const HeadArgs &head, const TcGeo &g, cudaStream_t st)
{
    using C = XCfg<COUT, S, PADACC>;
    const int nimg = g.nimg, H = g.H, W = g.W;
    CUtensorMap m0, m1;
    SQ_TRY(make_map(&m0, in0, nimg, g.D, cb0, H, W, C::PW, C::PH, 2 * KPS, ZPS));
    if (in1) SQ_TRY(make_map(&m1, in1, nimg, g.D, cb1, H, W, C::PW, C::PH, 2 * KPS, ZPS));
    else m1 = m0;
    static_assert(MINB * NBUF * C::ACC_COLS <= 512, "co-resident CTAs must fit in TMEM");
    const int budget = (MINB == 1 ? 200 : 216 / MINB) * 1024 - 2048;
    const int qsteps = g.KZ * (cb0 + (in1 ? cb1 : 0)) / 2;
    const int wres = qsteps * C::B_BYTES, stage_bytes = ZPS * KPS * C::A_BYTES;
    const int nstages = std::min(XC_MAX_STAGES, (budget - wres) / stage_bytes);
    SQ_REQUIRE(nstages >= 2 && g.KZ % ZPS == 0, SQ_ESTATE, "conv_xc: configuration does not fit");
    const size_t smem = (size_t)wres + (size_t)nstages * stage_bytes + 1024;
    auto kern = conv_xc_kernel<COUT, S, NBUF, MINB, EPI, HK, PADACC, KPS, ZPS>;
    static size_t attr_smem[64] = {0};          // the attribute is per device: one slot per device id
    size_t &have = attr_smem[u->h->device & 63];
    if (smem > have) {
        SQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
    const int tiles = nimg * g.D * ((W + C::TW - 1) / C::TW) * ((H + C::TH - 1) / C::TH);
    const int grid = std::min(tiles, MINB * u->h->sm_count / grid_div());
    long long *phase_dbg = nullptr;
#ifdef SQ_XC_PHASE_DIAG
    if (getenv("SQ_XC_PHASE")) {
        SQ_CUDA(cudaMalloc(&phase_dbg, (size_t)grid * 8 * sizeof(long long)));
        SQ_CUDA(cudaMemset(phase_dbg, 0, (size_t)grid * 8 * sizeof(long long)));
    }
#endif
    kern<<<grid, TC_THREADS, smem, st>>>(m0, m1, cb0 / 2, in1 ? cb1 / 2 : 0, (const bf16 *)L.w_xc, L.scale,
                                         L.shift, out, out_pool, head, nimg, H, W, nstages, g.D, g.KZ, wres,
                                         phase_dbg);
    ++u->last_launches;
    SQ_CHECK_LAUNCH();
    if (phase_dbg) {
        // diagnostics: average clocks per tile each role spent waiting / working
        SQ_CUDA(cudaStreamSynchronize(st));
        std::vector<long long> h((size_t)grid * 8);
        SQ_CUDA(cudaMemcpy(h.data(), phase_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        SQ_CUDA(cudaFree(phase_dbg));
        double a[8] = {0};
        for (int b2 = 0; b2 < grid; ++b2) for (int i = 0; i < 8; ++i) a[i] += (double)h[(size_t)b2 * 8 + i] / grid;
        const double tpc = (double)tiles / grid;
        fprintf(stderr, "xc_phase %s Cout=%d S=%d KPS=%d ZPS=%d stages=%d tiles/CTA=%.0f | per tile clk: producer wait_empty %.0f issue %.0f | "
                "mma wait_tempty %.0f wait_full %.0f issue %.0f | epilogue wait_tfull %.0f work %.0f\n", L.scope.c_str(), COUT, S, KPS, ZPS,
                nstages, tpc, a[0] / tpc, a[1] / tpc, a[2] / tpc, a[3] / tpc, a[4] / tpc, a[5] / tpc, a[6] / tpc);
    }
    return SQ_OK;
}

// Pick the largest box that fits: all three depth taps (volumes) and two channel-block pairs per
// stage if at least two stages remain next to the resident weights; SQ_NOT_APPLICABLE if even
// single-k-step stages do not fit.
template <int COUT, int S, int NBUF, int MINB, int EPI, int HK, bool PADACC>
int launch_xc(sq_unet_s *u, const SqLayer &L, const bf16 *in0, int cb0, const bf16 *in1, int cb1,
              bf16 *out, bf16 *out_pool, const HeadArgs &head, const TcGeo &g, cudaStream_t st)
{
    using C = XCfg<COUT, S, PADACC>;
    const int budget = (MINB == 1 ? 200 : 216 / MINB) * 1024 - 2048;
    const int qsteps = g.KZ * (cb0 + (in1 ? cb1 : 0)) / 2;
    const int room = budget - qsteps * C::B_BYTES;
    const bool pairable = (cb0 % 4 == 0) && (!in1 || cb1 % 4 == 0);
    auto fits = [&](int zps, int kps) { return room >= 2 * zps * kps * C::A_BYTES; };
#define SQ_XC_GO(KPS, ZPS) \
    return launch_xc_k<COUT, S, NBUF, MINB, EPI, HK, PADACC, KPS, ZPS>(u, L, in0, cb0, in1, cb1, out, out_pool, head, g, st)
    if (g.KZ == 3) {
        if (pairable && fits(3, 2)) SQ_XC_GO(2, 3);
        if (fits(3, 1)) SQ_XC_GO(1, 3);
    }
    if (pairable && fits(1, 2)) SQ_XC_GO(2, 1);
    if (fits(1, 1)) SQ_XC_GO(1, 1);
#undef SQ_XC_GO
    return SQ_NOT_APPLICABLE;
}

template <int COUT, int S, int NBUF, int MINB, bool PADACC = true>
int conv3x3_xc(sq_unet_s *u, const SqLayer &L, const bf16 *in0, int c0, const bf16 *in1, int c1,
               bf16 *out, bf16 *out_pool, const HeadArgs *head, const TcGeo &g, cudaStream_t st)
{
    const HeadArgs none = {nullptr, 0, nullptr, nullptr, nullptr};
    if (head) {
        switch (head->K) {
        case 2: return launch_xc<COUT, S, NBUF, MINB, EPI_HEAD, 2, PADACC>(u, L, in0, c0 / 8, in1, c1 / 8, out, nullptr, *head, g, st);
        case 3: return launch_xc<COUT, S, NBUF, MINB, EPI_HEAD, 3, PADACC>(u, L, in0, c0 / 8, in1, c1 / 8, out, nullptr, *head, g, st);
        case 4: return launch_xc<COUT, S, NBUF, MINB, EPI_HEAD, 4, PADACC>(u, L, in0, c0 / 8, in1, c1 / 8, out, nullptr, *head, g, st);
        }
        SQ_REQUIRE(false, SQ_EUNSUPPORTED, "fused head supports 2..4 classes");
    }
    if (out_pool)
        return launch_xc<COUT, S, NBUF, MINB, EPI_POOL, 0, PADACC>(u, L, in0, c0 / 8, in1, c1 / 8, out, out_pool, none, g, st);
    return launch_xc<COUT, S, NBUF, MINB, EPI_STORE, 0, PADACC>(u, L, in0, c0 / 8, in1, c1 / 8, out, nullptr, none, g, st);
}

// Fused conv_block (conv1 -> conv2 in one launch, the intermediate stays on chip): planar layers with
// Cout = 16 whose stage-1 input fits one ring stage.  Opt-in (SQ_PAIR=1): parity-green but SLOWER than the two
// launches on B200 (up0: 0.92 ms against 0.296 + 0.230 ms per 4 frames of 2048^2) -- the x-combined epilogue
// costs ~260 clk per 128-lane x 16-channel item per SM (half of it the 32 shuffles at one warp-shuffle per
// clock), a tile needs four of them, and with two accumulator buffers the chain MMA1 -> epilogue 1 -> MMA2 ->
// epilogue 2 of a tile serialises.
bool pair_enabled()
{
    const char *e = getenv("SQ_PAIR");
    return e && atoi(e) != 0;
}

template <int COUT, int S, int EPI, int HK>
int launch_pair_k(sq_unet_s *u, const SqLayer &L1, const SqLayer &L2, const bf16 *in0, int cb0, const bf16 *in1,
                  int cb1, bf16 *out, bf16 *out_pool, const HeadArgs &head, const TcGeo &g, cudaStream_t st)
{
    using C = PCfg<COUT, S>;
    const int nimg = g.nimg, H = g.H, W = g.W;
    const int ks0 = cb0 / 2, ks1 = in1 ? cb1 / 2 : 0, ksteps1 = ks0 + ks1;
    CUtensorMap m0, m1;
    SQ_TRY(make_map(&m0, in0, nimg, 1, cb0, H, W, C::PW, C::PH, cb0, 1));
    if (in1) SQ_TRY(make_map(&m1, in1, nimg, 1, cb1, H, W, C::PW, C::PH, cb1, 1));
    else m1 = m0;
    const int budget = (216 / 2) * 1024 - 2048;
    const int fixed = ((ksteps1 + COUT / 16) * C::B_BYTES + 127) / 128 * 128 + 2 * C::P1_BYTES;
    const int stage_bytes = ksteps1 * C::A_BYTES;
    const int nstages = std::min(XC_MAX_STAGES, (budget - fixed) / stage_bytes);
    if (nstages < 2) return SQ_NOT_APPLICABLE;
    const size_t smem = (size_t)fixed + (size_t)nstages * stage_bytes + 1024;
    auto kern = conv_xc_pair_kernel<COUT, S, EPI, HK>;
    static size_t attr_smem[64] = {0};
    size_t &have = attr_smem[u->h->device & 63];
    if (smem > have) {
        SQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
    const int tiles = nimg * ((W + C::TW - 1) / C::TW) * ((H + C::TH - 1) / C::TH);
    const int grid = std::min(tiles, 2 * u->h->sm_count / grid_div());
    kern<<<grid, TC_THREADS, smem, st>>>(m0, m1, ks0, ks1, (const bf16 *)L1.w_xc, L1.scale, L1.shift,
                                         (const bf16 *)L2.w_xc, L2.scale, L2.shift, out, out_pool, head, nimg, H, W,
                                         nstages);
    ++u->last_launches;
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

// SQ_NOT_APPLICABLE: the caller runs the two layers separately
int conv_pair_tc(sq_unet_s *u, const SqLayer &L1, const SqLayer &L2, const bf16 *in0, int c0, const bf16 *in1, int c1,
                 bf16 *out, bf16 *out_pool, const HeadArgs *head, const TcGeo &g, cudaStream_t st)
{
    if (!pair_enabled() || g.D != 1 || g.KZ != 1 || !L1.w_xc || !L2.w_xc || L1.cout != L2.cout ||
        L2.cin0 != L1.cout || L2.cin1 != 0 || c0 % 16 || c1 % 16 || (g.H & 1) || (g.W & 1))
        return SQ_NOT_APPLICABLE;
    const HeadArgs none = {nullptr, 0, nullptr, nullptr, nullptr};
    if (L1.cout == 16) {
        if (head) {
            switch (head->K) {
            case 2: return launch_pair_k<16, 2, EPI_HEAD, 2>(u, L1, L2, in0, c0 / 8, in1, c1 / 8, nullptr, nullptr, *head, g, st);
            case 3: return launch_pair_k<16, 2, EPI_HEAD, 3>(u, L1, L2, in0, c0 / 8, in1, c1 / 8, nullptr, nullptr, *head, g, st);
            case 4: return launch_pair_k<16, 2, EPI_HEAD, 4>(u, L1, L2, in0, c0 / 8, in1, c1 / 8, nullptr, nullptr, *head, g, st);
            }
            return SQ_NOT_APPLICABLE;
        }
        if (out_pool) return launch_pair_k<16, 2, EPI_POOL, 0>(u, L1, L2, in0, c0 / 8, in1, c1 / 8, out, out_pool, none, g, st);
        return launch_pair_k<16, 2, EPI_STORE, 0>(u, L1, L2, in0, c0 / 8, in1, c1 / 8, out, nullptr, none, g, st);
    }
    return SQ_NOT_APPLICABLE;
}

int xc_variant()
{
    // SQ_XC=0 disables the x-combined kernels, SQ_XC=2 forces them (A/B measurements, tests)
    const char *e = getenv("SQ_XC");
    return e ? atoi(e) : 1;
}

bool first_fusion_enabled()
{
    // SQ_FUSE_FIRST=1 opts in.  Bit-identical to the two-launch path but slower on B200 (0.60 vs 0.147 + 0.225 ms
    // per 4 frames of 2048^2): at N = 16 the tensor core's A-operand reads already use most of the 128 B/clk of
    // shared-memory bandwidth, and the builders' reads and writes queue behind them.
    const char *e = getenv("SQ_FUSE_FIRST");
    return e && atoi(e) != 0;
}

bool use_clusters()
{
    // SQ_CLUSTER=1 opts in: measured neutral on B200 (profiles/README.md), so plain launches stay the default
    const char *e = getenv("SQ_CLUSTER");
    return e && atoi(e) != 0;
}

// out_pool != NULL: also write the 2x2 max-pooled tensor; head != NULL: fused 1x1 head, no `out`.
int conv3x3_tc(sq_unet_s *u, const SqLayer &L, const bf16 *in0, int c0, const bf16 *in1, int c1,
               bf16 *out, bf16 *out_pool, const HeadArgs *head, const TcGeo &g, cudaStream_t st)
{
    // x-combined kernel: its MMA phase costs ~0.4x the 9-tap kernel's per k-step, but its epilogue reads
    // 3x the accumulator columns and tcgen05.ld moves only 64 B/clk/SM (3.4 clk/px at Cout 16, 6.9 at
    // Cout 32, against 2.7-2.8 clk/px PER K-STEP of 9-tap MMAs): it wins from 2 k-steps per tile at
    // Cout 16 and from 3 at Cout 32 (profiles/r1_xc_ablation.log).  SQ_XC=0 disables it, 2 forces it.
    const int v = xc_variant();
    const int qsteps = g.KZ * ((c0 + c1) / 16);
    if (L.w_xc && v > 0 && (qsteps >= (L.cout == 16 ? 2 : 3) || v == 2)) {
        int r = SQ_NOT_APPLICABLE;
        if (L.cout == 16) r = conv3x3_xc<16, 2, 2, 2>(u, L, in0, c0, in1, c1, out, out_pool, head, g, st);
        if (L.cout == 32) r = conv3x3_xc<32, 1, 2, 2>(u, L, in0, c0, in1, c1, out, out_pool, head, g, st);
        if (r != SQ_NOT_APPLICABLE) return r;
    }
    switch (L.cout) {
    case 16:  return conv3x3_epi<16, 4, 2>(u, L, in0, c0, in1, c1, out, out_pool, head, g, st);
    case 32:  return conv3x3_epi<32, 4, 2>(u, L, in0, c0, in1, c1, out, out_pool, head, g, st);
    case 64:  return conv3x3_epi<64, 2, 2>(u, L, in0, c0, in1, c1, out, out_pool, head, g, st);
    case 128:   // 2 MMA warps; SQ_CLUSTER=1: weight multicast across CTA pairs
        if (use_clusters()) return conv3x3_epi<128, 2, 1, 2, true>(u, L, in0, c0, in1, c1, out, out_pool, head, g, st);
        return conv3x3_epi<128, 2, 1, 2>(u, L, in0, c0, in1, c1, out, out_pool, head, g, st);
    case 256:
        if (use_clusters()) return conv3x3_epi<256, 1, 1, 1, true>(u, L, in0, c0, in1, c1, out, out_pool, head, g, st);
        return conv3x3_epi<256, 1, 1>(u, L, in0, c0, in1, c1, out, out_pool, head, g, st);
    }
    SQ_REQUIRE(false, SQ_EUNSUPPORTED, "bf16 mode: unsupported filter count %d", L.cout);
}

int upconv_tc(sq_unet_s *u, const SqLayer &L, const bf16 *in, bf16 *out, const TcGeo &g, cudaStream_t st)
{
    const HeadArgs none = {nullptr, 0, nullptr, nullptr, nullptr};
    switch (L.cout) {
    case 16:  return launch_tc<16, 2, true, 2, 2, EPI_STORE>(u, L, in, L.cin0 / 8, nullptr, 0, out, nullptr, none, g, 0, st);
    case 32:  return launch_tc<32, 1, true, 2, 2, EPI_STORE>(u, L, in, L.cin0 / 8, nullptr, 0, out, nullptr, none, g, 0, st);
    case 64:  return launch_tc<64, 1, true, 2, 1, EPI_STORE>(u, L, in, L.cin0 / 8, nullptr, 0, out, nullptr, none, g, 0, st);
    case 128: return launch_tc<128, 1, true, 1, 1, EPI_STORE>(u, L, in, L.cin0 / 8, nullptr, 0, out, nullptr, none, g, 0, st);
    }
    SQ_REQUIRE(false, SQ_EUNSUPPORTED, "bf16 mode: unsupported up-conv filter count %d", L.cout);
}

// ---- quad (space-to-depth) level-0 launches.  g.H, g.W: the quad image (half the frame).
bool quad_enabled()
{
    // SQ_QUAD=0 keeps level 0 on the full-resolution kernels (A/B measurements, tests)
    const char *e = getenv("SQ_QUAD");
    return !e || atoi(e) != 0;
}

template <int NTAP, int KPS, int EPI, int HK>
int launch_qd(sq_unet_s *u, const SqLayer &L, const bf16 *in0, int cb0, const bf16 *in1, int cb1, bf16 *out,
              bf16 *out_pool, const HeadArgs &head, const TcGeo &g, int relu, cudaStream_t st)
{
    constexpr int S = 2;
    using C = QCfg<S, NTAP>;
    const int nimg = g.nimg, H = g.H, W = g.W;
    CUtensorMap m0, m1;
    SQ_TRY(make_map(&m0, in0, nimg, 1, cb0, H, W, C::PW, C::PH, 2 * KPS, 1));
    if (in1) SQ_TRY(make_map(&m1, in1, nimg, 1, cb1, H, W, C::PW, C::PH, 2 * KPS, 1));
    else m1 = m0;
    const int ks0 = cb0 / 2, ks1 = in1 ? cb1 / 2 : 0, ksteps = ks0 + ks1;
    SQ_REQUIRE(ks0 % KPS == 0 && ks1 % KPS == 0 && (NTAP == 1 || (ks0 % 4 == 0 && ks1 % 4 == 0)), SQ_ESTATE,
               "conv_qd: %d + %d k-steps do not fit the stage layout", ks0, ks1);
    const int budget = (216 / 2) * 1024 - 2048;
    const int wres = ksteps * C::W_BYTES, stage_bytes = KPS * C::A_BYTES;
    const int nstages = std::min(QD_MAX_STAGES, (budget - wres) / stage_bytes);
    SQ_REQUIRE(nstages >= 2, SQ_ESTATE, "conv_qd: configuration does not fit in shared memory");
    const size_t smem = (size_t)((wres + 127) & ~127) + (size_t)nstages * stage_bytes + 1024;
    auto kern = conv_qd_kernel<S, NTAP, KPS, EPI, HK>;
    static size_t attr_smem[64] = {0};
    size_t &have = attr_smem[u->h->device & 63];
    if (smem > have) {
        SQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
    const int tiles = nimg * ((W + 7) / 8) * ((H + C::TH - 1) / C::TH);
    const int grid = std::min(tiles, 2 * u->h->sm_count / grid_div());
    long long *phase_dbg = nullptr;
#ifdef SQ_XC_PHASE_DIAG
    if (getenv("SQ_XC_PHASE")) {
        SQ_CUDA(cudaMalloc(&phase_dbg, (size_t)grid * 8 * sizeof(long long)));
        SQ_CUDA(cudaMemset(phase_dbg, 0, (size_t)grid * 8 * sizeof(long long)));
    }
#endif
    QdEpi ep = {};
    {
        const std::vector<float> &s0 = u->host[L.scope + "/_scale"].data, &t0 = u->host[L.scope + "/_shift"].data;
        for (int i = 0; i < 16; ++i) { ep.sc[i] = s0[i]; ep.sh[i] = t0[i]; }
        if (EPI == EPI_HEAD) {
            // bf16-rounded head weights [channel][class] and the fp32 bias, as the stand-alone head and the other fused heads use them
            const std::vector<float> &hk = u->host["UNet/to_image/kernel"].data, &hb = u->host["UNet/to_image/bias"].data;
            for (int c = 0; c < 16; ++c)
                for (int k2 = 0; k2 < HK; ++k2) ep.hw[k2][c] = host_bf16_round(hk[(size_t)c * HK + k2]);
            for (int k2 = 0; k2 < HK; ++k2) ep.hb[k2] = hb[k2];
        }
    }
    kern<<<grid, TC_THREADS, smem, st>>>(m0, m1, ks0, ks1, (const bf16 *)L.w_qd, ep, out, out_pool, head, nimg, H, W, relu,
                                         nstages, wres, phase_dbg);
    ++u->last_launches;
    SQ_CHECK_LAUNCH();
    if (phase_dbg) {
        // diagnostics (-DSQ_XC_PHASE_DIAG builds only): average clocks per tile each role spent waiting / working
        SQ_CUDA(cudaStreamSynchronize(st));
        std::vector<long long> h((size_t)grid * 8);
        SQ_CUDA(cudaMemcpy(h.data(), phase_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        SQ_CUDA(cudaFree(phase_dbg));
        double a[8] = {0};
        for (int b2 = 0; b2 < grid; ++b2) for (int i = 0; i < 8; ++i) a[i] += (double)h[(size_t)b2 * 8 + i] / grid;
        const double tpc = (double)tiles / grid;
        fprintf(stderr, "qd_phase %s NTAP=%d KPS=%d EPI=%d stages=%d tiles/CTA=%.0f | per tile clk: producer wait_empty %.0f issue %.0f | "
                "mma wait_tempty %.0f wait_full %.0f issue %.0f | epilogue wait_tfull %.0f work %.0f\n", L.scope.c_str(), NTAP, KPS, EPI,
                nstages, tpc, a[0] / tpc, a[1] / tpc, a[2] / tpc, a[3] / tpc, a[4] / tpc, a[5] / tpc, a[6] / tpc);
    }
    return SQ_OK;
}

bool qfuse_enabled()
{
    // SQ_QFUSE=0: down0/conv1 and down0/conv2 as two launches (A/B measurements, tests)
    const char *e = getenv("SQ_QFUSE");
    return !e || atoi(e) != 0;
}

// down0/conv1 + down0/conv2 + pool in one launch (conv_qf_kernel); g.H, g.W: the quad image.  raw16 != NULL: the
// frames are raw uint16 values normalised in the loader with the per-frame (mean, std) in `stats`.
int launch_qf(sq_unet_s *u, const SqLayer &L1, const SqLayer &L2, const float *in, const uint16_t *raw16,
              const float2 *stats, bf16 *out, bf16 *out_pool, const TcGeo &g, cudaStream_t st)
{
    sq_encode_tiled_fn enc = sq_get_encode_tiled();
    SQ_REQUIRE(enc, SQ_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    const int H0 = 2 * g.H, W0 = 2 * g.W;
    const size_t esz = raw16 ? 2 : 4;
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)W0, (cuuint64_t)H0, (cuuint64_t)g.nimg};
    cuuint64_t strides[2] = {(cuuint64_t)W0 * esz, (cuuint64_t)H0 * W0 * esz};
    cuuint32_t box[3] = {(cuuint32_t)QF::RAW_W, (cuuint32_t)QF::RAW_H, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, raw16 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                     raw16 ? (void *)raw16 : (void *)in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SQ_REQUIRE(r == CUDA_SUCCESS, SQ_ECUDA, "cuTensorMapEncodeTiled failed (%d) for the frames (%d,%d,%d)", (int)r,
               g.nimg, H0, W0);
    const size_t smem = (size_t)QF::SMEM + 1024;
    static size_t attr_smem[64][2] = {{0}};
    size_t &have = attr_smem[u->h->device & 63][raw16 ? 1 : 0];
    if (smem > have) {
        if (raw16) SQ_CUDA(cudaFuncSetAttribute(conv_qf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else SQ_CUDA(cudaFuncSetAttribute(conv_qf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
    const int tiles = g.nimg * ((g.W + 7) / 8) * ((g.H + QF::TH - 1) / QF::TH);
    const int grid = std::min(tiles, u->h->sm_count / grid_div());
    long long *phase_dbg = nullptr;
#ifdef SQ_XC_PHASE_DIAG
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (getenv("SQ_XC_PHASE")) {
        SQ_CUDA(cudaMalloc(&phase_dbg, ((size_t)grid * 32 + 1) * sizeof(long long)));
        SQ_CUDA(cudaMemset(phase_dbg, 0, ((size_t)grid * 32 + 1) * sizeof(long long)));
        const long long flag = atoi(getenv("SQ_XC_PHASE"));            // 2: experiment without the epilogue-2 stores
        SQ_CUDA(cudaMemcpy(phase_dbg + (size_t)grid * 32, &flag, sizeof flag, cudaMemcpyHostToDevice));
        cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st);
    }
#endif
    QfEpi ep;
    {
        const std::vector<float> &s1 = u->host[L1.scope + "/_scale"].data, &t1 = u->host[L1.scope + "/_shift"].data;
        const std::vector<float> &s2 = u->host[L2.scope + "/_scale"].data, &t2 = u->host[L2.scope + "/_shift"].data;
        for (int i = 0; i < 16; ++i) { ep.sc1[i] = s1[i]; ep.sh1[i] = t1[i]; ep.sc2[i] = s2[i]; ep.sh2[i] = t2[i]; }
    }
    if (raw16)
        conv_qf_kernel<true><<<grid, QF::THREADS, smem, st>>>(m, (const bf16 *)L2.w_qf, ep, stats, out, out_pool, g.nimg, g.H,
                                                             g.W, phase_dbg);
    else
        conv_qf_kernel<false><<<grid, QF::THREADS, smem, st>>>(m, (const bf16 *)L2.w_qf, ep, nullptr, out, out_pool, g.nimg,
                                                              g.H, g.W, phase_dbg);
    ++u->last_launches;
    SQ_CHECK_LAUNCH();
#ifdef SQ_XC_PHASE_DIAG
    if (phase_dbg) {
        cudaEventRecord(e1, st);
        SQ_CUDA(cudaStreamSynchronize(st));
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h((size_t)grid * 32);
        SQ_CUDA(cudaMemcpy(h.data(), phase_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        SQ_CUDA(cudaFree(phase_dbg));
        double a[32] = {0};
        for (int b2 = 0; b2 < grid; ++b2) for (int i = 0; i < 32; ++i) a[i] += (double)h[(size_t)b2 * 32 + i] / grid;
        const double tpc = (double)tiles / grid;
        fprintf(stderr, "qf_phase %.3f ms, tiles/CTA=%.0f (%.0f clk/tile at 1.965 GHz) | per tile clk: producer wait raw_empty %.0f | mma wait a1_full %.0f "
                "acc1_empty %.0f p_full %.0f acc2_empty %.0f | builders wait raw_full %.0f a1_empty %.0f | epilogue wait acc1_full %.0f "
                "p_empty %.0f acc2_full %.0f; work: epilogue 1 %.0f, epilogue 2 %.0f; MMA warp lifetime %.0f ticks = %.3f GHz-equivalent\n", ms, tpc, ms * 1.965e6 / tpc, a[0] / tpc, a[8] / tpc, a[9] / tpc, a[10] / tpc, a[11] / tpc,
                a[16] / tpc, a[17] / tpc, a[24] / tpc, a[25] / tpc, a[26] / tpc, a[27] / tpc, a[28] / tpc, a[13], a[13] / (ms * 1e6));
    }
#endif
    return SQ_OK;
}

bool qup_enabled()
{
    // SQ_QUP=0: up0/upscale and up0/conv1 as two launches (A/B measurements, tests)
    const char *e = getenv("SQ_QUP");
    return !e || atoi(e) != 0;
}

// up0/upscale + up0/conv1 (concat bridge) in one launch (conv_qu_kernel); g.H, g.W: the quad image
int launch_qu(sq_unet_s *u, const SqLayer &Lu, const SqLayer &L1, const bf16 *cur, const bf16 *skip, bf16 *out,
              const TcGeo &g, cudaStream_t st)
{
    const int cbu = Lu.cin0 / 8, ksu = Lu.cin0 / 16;
    CUtensorMap mc, ms;
    SQ_TRY(make_map(&mc, cur, g.nimg, 1, cbu, g.H, g.W, QU::PW, QU::PH, cbu, 1));
    SQ_TRY(make_map(&ms, skip, g.nimg, 1, 8, g.H, g.W, QU::PW, QU::PH, 2, 1));
    const size_t smem = (size_t)QU::SMEM + 1024;
    static size_t attr_smem[64] = {0};
    size_t &have = attr_smem[u->h->device & 63];
    if (smem > have) {
        SQ_CUDA(cudaFuncSetAttribute(conv_qu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
    QfEpi ep;
    {
        const std::vector<float> &s1 = u->host[Lu.scope + "/_scale"].data, &t1 = u->host[Lu.scope + "/_shift"].data;
        const std::vector<float> &s2 = u->host[L1.scope + "/_scale"].data, &t2 = u->host[L1.scope + "/_shift"].data;
        for (int i = 0; i < 16; ++i) { ep.sc1[i] = s1[i]; ep.sh1[i] = t1[i]; ep.sc2[i] = s2[i]; ep.sh2[i] = t2[i]; }
    }
    const int tiles = g.nimg * ((g.W + 7) / 8) * ((g.H + QU::TH - 1) / QU::TH);
    const int grid = std::min(tiles, u->h->sm_count / grid_div());
    long long *phase_dbg = nullptr;
#ifdef SQ_XC_PHASE_DIAG
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (getenv("SQ_XC_PHASE")) {
        SQ_CUDA(cudaMalloc(&phase_dbg, (size_t)grid * 24 * sizeof(long long)));
        SQ_CUDA(cudaMemset(phase_dbg, 0, (size_t)grid * 24 * sizeof(long long)));
        cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st);
    }
#endif
    conv_qu_kernel<<<grid, QU::THREADS, smem, st>>>(mc, ms, (const bf16 *)Lu.w_qd, (const bf16 *)L1.w_qu, ep, ksu, out, g.nimg,
                                                   g.H, g.W, phase_dbg);
    ++u->last_launches;
    SQ_CHECK_LAUNCH();
#ifdef SQ_XC_PHASE_DIAG
    if (phase_dbg) {
        cudaEventRecord(e1, st);
        SQ_CUDA(cudaStreamSynchronize(st));
        float ms_ = 0; cudaEventElapsedTime(&ms_, e0, e1);
        std::vector<long long> h((size_t)grid * 24);
        SQ_CUDA(cudaMemcpy(h.data(), phase_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        SQ_CUDA(cudaFree(phase_dbg));
        double a[24] = {0};
        for (int b2 = 0; b2 < grid; ++b2) for (int i = 0; i < 24; ++i) a[i] += (double)h[(size_t)b2 * 24 + i] / grid;
        const double tpc = (double)tiles / grid;
        fprintf(stderr, "qu_phase %.3f ms, tiles/CTA=%.0f, %.0f ticks/tile (MMA warp lifetime) | per tile: producer wait a1_empty %.0f ring_empty %.0f | "
                "mma wait p_full %.0f a1_full %.0f acc1_empty %.0f acc2_empty %.0f ring_full %.0f | epilogue wait acc1_full %.0f p_empty %.0f "
                "acc2_full %.0f; work: epilogue 1 %.0f, epilogue 2 %.0f\n", ms_, tpc, a[15] / tpc, a[0] / tpc, a[1] / tpc, a[8] / tpc, a[9] / tpc,
                a[10] / tpc, a[11] / tpc, a[12] / tpc, a[16] / tpc, a[17] / tpc, a[18] / tpc, a[19] / tpc, a[20] / tpc);
    }
#endif
    return SQ_OK;
}

// quad 3x3 conv; c0 / c1: level-0 channels of the two sources (16 each) -> 4*c/8 channel blocks of the quad image
int conv3x3_qd(sq_unet_s *u, const SqLayer &L, const bf16 *in0, int c0, const bf16 *in1, int c1, bf16 *out,
               bf16 *out_pool, const HeadArgs *head, const TcGeo &g, cudaStream_t st)
{
    const HeadArgs none = {nullptr, 0, nullptr, nullptr, nullptr};
    const int cb0 = 4 * c0 / 8, cb1 = 4 * c1 / 8;
    if (head) {
        switch (head->K) {
        case 2: return launch_qd<4, 2, EPI_HEAD, 2>(u, L, in0, cb0, in1, cb1, nullptr, nullptr, *head, g, 1, st);
        case 3: return launch_qd<4, 2, EPI_HEAD, 3>(u, L, in0, cb0, in1, cb1, nullptr, nullptr, *head, g, 1, st);
        case 4: return launch_qd<4, 2, EPI_HEAD, 4>(u, L, in0, cb0, in1, cb1, nullptr, nullptr, *head, g, 1, st);
        }
        SQ_REQUIRE(false, SQ_EUNSUPPORTED, "fused head supports 2..4 classes");
    }
    if (out_pool) return launch_qd<4, 2, EPI_POOL, 0>(u, L, in0, cb0, in1, cb1, out, out_pool, none, g, 1, st);
    // two sources: 8 k-steps of resident weights (64 KB) leave room for single-k-step stages only
    if (in1) return launch_qd<4, 1, EPI_STORE, 0>(u, L, in0, cb0, in1, cb1, out, nullptr, none, g, 1, st);
    return launch_qd<4, 2, EPI_STORE, 0>(u, L, in0, cb0, in1, cb1, out, nullptr, none, g, 1, st);
}

SqLayer *layer_by_scope(sq_unet_s *u, const std::string &scope)
{
    for (SqLayer &l : u->layers)
        if (l.scope == scope) return &l;
    return nullptr;
}

}  // namespace

int sq_tc_destroy(sq_unet_s *u)
{
    delete (TcState *)u->tc_state;
    u->tc_state = nullptr;
    return SQ_OK;
}

// Re-lay-out the weights for the tensor-core kernels (called from sq_unet_finalize after the
// fp32 upload, so L.scale / L.shift already hold the folded epilogue).
int sq_tc_finalize(sq_unet_s *u)
{
    SQ_REQUIRE(u->cin <= 4, SQ_EUNSUPPORTED, "bf16 mode: num_inputs must be <= 4 (got %d)", u->cin);
    for (int f : u->filters)
        SQ_REQUIRE(f == 16 || f == 32 || f == 64 || f == 128 || f == 256, SQ_EUNSUPPORTED,
                   "bf16 mode: filters must be in {16,32,64,128,256} (got %d)", f);
    const int KZ = (u->ndim == 3) ? 3 : 1, UZ = (u->ndim == 3) ? 2 : 1;
    for (SqLayer &L : u->layers) {
        const std::vector<float> &k = u->host[L.scope + "/kernel"].data;
        const int C = L.cin0 + L.cin1, CO = L.cout;
        if (L.kind == SqLayer::CONV && C % 16 == 0) {
            // [kz][ks][tap9][kb][co][8]  <-  (D)HWIO kernel[kz*9 + tap9][ci][co]
            std::vector<uint16_t> w((size_t)KZ * 9 * C * CO);
            for (int kz = 0; kz < KZ; ++kz)
                for (int ks = 0; ks < C / 16; ++ks)
                    for (int tp = 0; tp < 9; ++tp)
                        for (int kb = 0; kb < 2; ++kb)
                            for (int co = 0; co < CO; ++co)
                                for (int e = 0; e < 8; ++e) {
                                    const int ci = ks * 16 + kb * 8 + e;
                                    w[(((((size_t)kz * (C / 16) + ks) * 9 + tp) * 2 + kb) * CO + co) * 8 + e] =
                                        host_bf16(k[((size_t)(kz * 9 + tp) * C + ci) * CO + co]);
                                }
            SQ_TRY(dev_upload(u, w.data(), w.size() * 2, &L.w_tc));
            if (CO <= 32) {
                // x-combined layout: [kz][ks][ky][kb][(kx, co)][8]
                std::vector<uint16_t> x((size_t)KZ * 9 * C * CO);
                const int N = 3 * CO;
                for (int kz = 0; kz < KZ; ++kz)
                    for (int ks = 0; ks < C / 16; ++ks)
                        for (int ky = 0; ky < 3; ++ky)
                            for (int kb = 0; kb < 2; ++kb)
                                for (int kx = 0; kx < 3; ++kx)
                                    for (int co = 0; co < CO; ++co)
                                        for (int e = 0; e < 8; ++e) {
                                            const int ci = ks * 16 + kb * 8 + e;
                                            x[((((((size_t)kz * (C / 16) + ks) * 3 + ky) * 2 + kb) * N) + kx * CO + co) * 8 + e] =
                                                host_bf16(k[((size_t)(kz * 9 + ky * 3 + kx) * C + ci) * CO + co]);
                                        }
                SQ_TRY(dev_upload(u, x.data(), x.size() * 2, &L.w_xc));
            }
        } else if (L.kind == SqLayer::UPCONV) {
            // [kz][ks][tap4][kb][co][8]  <-  TF conv_transpose kernel[kz*4 + tap4][co][ci]
            std::vector<uint16_t> w((size_t)UZ * 4 * C * CO);
            for (int kz = 0; kz < UZ; ++kz)
                for (int ks = 0; ks < C / 16; ++ks)
                    for (int tp = 0; tp < 4; ++tp)
                        for (int kb = 0; kb < 2; ++kb)
                            for (int co = 0; co < CO; ++co)
                                for (int e = 0; e < 8; ++e) {
                                    const int ci = ks * 16 + kb * 8 + e;
                                    w[(((((size_t)kz * (C / 16) + ks) * 4 + tp) * 2 + kb) * CO + co) * 8 + e] =
                                        host_bf16(k[((size_t)(kz * 4 + tp) * CO + co) * C + ci]);
                                }
            SQ_TRY(dev_upload(u, w.data(), w.size() * 2, &L.w_tc));
        } else {
            // first conv / head: CUDA-core kernels read fp32 copies of the bf16-rounded weights
            std::vector<float> w(k.size());
            for (size_t i = 0; i < k.size(); ++i) w[i] = host_bf16_round(k[i]);
            if (L.kind == SqLayer::HEAD) {              // fused head reads [C][K] weights then bias[K]
                const std::vector<float> &b = u->host[L.scope + "/bias"].data;
                w.insert(w.end(), b.begin(), b.end());
            }
            SQ_TRY(dev_upload(u, w.data(), w.size() * 4, &L.w_tc));
        }
    }
    // Level-0 layers of planar nets with filters[0] = 16 also get the quad (space-to-depth) form used by
    // conv_qd_kernel: [k-step = (source, input parity)][tap (a, b)][ci%16/8][(output parity, co)][8]
    if (u->ndim == 2 && u->nlev >= 2 && u->filters[0] == 16) {
        for (SqLayer &L : u->layers) {
            if (L.level != 0 || L.cout != 16) continue;
            const std::vector<float> &k = u->host[L.scope + "/kernel"].data;
            const int C = L.cin0 + L.cin1;
            std::vector<uint16_t> w;
            if (L.kind == SqLayer::CONV && C % 16 == 0) {
                w.assign((size_t)(C / 16) * 4 * 4 * 2 * 64 * 8, 0);
                for (int src = 0; src < C / 16; ++src)
                    for (int qi = 0; qi < 4; ++qi)                     // input parity (iy, ix)
                        for (int tp = 0; tp < 4; ++tp)                 // half-resolution tap (a, b): ty = a - iy, tx = b - ix
                            for (int kb = 0; kb < 2; ++kb)
                                for (int qo = 0; qo < 4; ++qo)         // output parity (oy, ox)
                                    for (int co = 0; co < 16; ++co)
                                        for (int e = 0; e < 8; ++e) {
                                            const int iy = qi >> 1, ix = qi & 1, oy = qo >> 1, ox = qo & 1;
                                            const int ky = 2 * ((tp >> 1) - iy) + iy - oy + 1, kx = 2 * ((tp & 1) - ix) + ix - ox + 1;
                                            if (ky < 0 || ky > 2 || kx < 0 || kx > 2) continue;
                                            const int ci = src * 16 + kb * 8 + e;
                                            w[((((((size_t)src * 4 + qi) * 4 + tp) * 2 + kb) * 64) + qo * 16 + co) * 8 + e] =
                                                host_bf16(k[((size_t)(ky * 3 + kx) * C + ci) * 16 + co]);
                                        }
            } else if (L.kind == SqLayer::UPCONV && C % 16 == 0) {
                // 1x1 conv C -> 64: [k-step][ci%16/8][(tap = output parity, co)][8]
                w.assign((size_t)C * 64, 0);
                for (int ks = 0; ks < C / 16; ++ks)
                    for (int kb = 0; kb < 2; ++kb)
                        for (int qo = 0; qo < 4; ++qo)
                            for (int co = 0; co < 16; ++co)
                                for (int e = 0; e < 8; ++e) {
                                    const int ci = ks * 16 + kb * 8 + e;
                                    w[((((size_t)ks * 2 + kb) * 64) + qo * 16 + co) * 8 + e] =
                                        host_bf16(k[((size_t)qo * 16 + co) * C + ci]);
                                }
            } else {
                continue;
            }
            SQ_TRY(dev_upload(u, w.data(), w.size() * 2, &L.w_qd));
            if (L.scope == "UNet/down0/conv2" && u->cin == 1 && C == 16) {
                // conv_qf_kernel: B1[(parity, co)][window position wy*4 + wx] of the first conv (one k-step), then this
                // layer's quad weights with the output columns reordered to (co / 8, parity, co % 8)
                const std::vector<float> &k1 = u->host["UNet/down0/conv1/kernel"].data;
                std::vector<uint16_t> f((size_t)(QF::B1_BYTES + QF::W2_BYTES) / 2, 0);
                for (int kb = 0; kb < 2; ++kb)
                    for (int qo = 0; qo < 4; ++qo)
                        for (int co = 0; co < 16; ++co)
                            for (int e = 0; e < 8; ++e) {
                                const int kk = kb * 8 + e, wy = kk >> 2, wx = kk & 3, ky = wy - (qo >> 1), kx = wx - (qo & 1);
                                if (ky < 0 || ky > 2 || kx < 0 || kx > 2) continue;
                                f[(((size_t)kb * 64) + qo * 16 + co) * 8 + e] = host_bf16(k1[(size_t)(ky * 3 + kx) * 16 + co]);
                            }
                const size_t base = QF::B1_BYTES / 2;
                for (int qi = 0; qi < 4; ++qi)
                    for (int tp = 0; tp < 4; ++tp)
                        for (int kb = 0; kb < 2; ++kb)
                            for (int qo = 0; qo < 4; ++qo)
                                for (int co = 0; co < 16; ++co)
                                    for (int e = 0; e < 8; ++e)
                                        f[base + (((((size_t)qi * 4 + tp) * 2 + kb) * 64) + (co / 8) * 32 + qo * 8 + co % 8) * 8 + e] =
                                            w[(((((size_t)qi * 4 + tp) * 2 + kb) * 64) + qo * 16 + co) * 8 + e];
                SQ_TRY(dev_upload(u, f.data(), f.size() * 2, &L.w_qf));
            }
            if (L.scope == "UNet/up0/conv1" && C == 32) {
                // conv_qu_kernel: the same weights with the output columns reordered to (co / 8, parity, co % 8)
                std::vector<uint16_t> f(w.size(), 0);
                for (size_t blk = 0; blk < w.size() / (64 * 8); ++blk)       // (k-step, tap, kb) blocks of [64][8]
                    for (int qo = 0; qo < 4; ++qo)
                        for (int co = 0; co < 16; ++co)
                            for (int e = 0; e < 8; ++e)
                                f[(blk * 64 + (co / 8) * 32 + qo * 8 + co % 8) * 8 + e] = w[(blk * 64 + qo * 16 + co) * 8 + e];
                SQ_TRY(dev_upload(u, f.data(), f.size() * 2, &L.w_qu));
            }
            std::vector<float> sc(64), sh(64);
            const std::vector<float> &s0 = u->host[L.scope + "/_scale"].data, &t0 = u->host[L.scope + "/_shift"].data;
            for (int i = 0; i < 64; ++i) { sc[i] = s0[i % 16]; sh[i] = t0[i % 16]; }
            SQ_TRY(dev_upload(u, sc.data(), 64 * 4, (void **)&L.scale_q));
            SQ_TRY(dev_upload(u, sh.data(), 64 * 4, (void **)&L.shift_q));
        }
    }
    TcState *s = new TcState();
    s->sm_count = u->h->sm_count;
    u->tc_state = s;
    return SQ_OK;
}

namespace {

int tc_run(sq_unet_s *u, bool dry, const float *in, int n, int dep, int hgt, int wid, float *probs,
           uint8_t *mask, float *logits, void *ws, size_t ws_bytes, cudaStream_t st, size_t *need,
           const uint16_t *raw16 = nullptr, const float2 *stats16 = nullptr)
{
    SqArena a(dry ? nullptr : ws, dry ? 0 : ws_bytes);
    const int nl = u->nlev;
    const bool vol = u->ndim == 3;
    const int KZ = vol ? 3 : 1;
    auto depth = [&](int l) { return vol ? (dep >> l) : 1; };
    std::vector<bf16 *> t1(nl), skip(nl), pooled(nl, nullptr), xyp(nl, nullptr), up(nl, nullptr),
        merged(nl, nullptr), ut(nl, nullptr), uo(nl, nullptr);
    for (int l = 0; l < nl; ++l) {
        const size_t px = (size_t)n * depth(l) * (hgt >> l) * (wid >> l);
        const size_t f = u->filters[l];
        t1[l] = a.take<bf16>(px * f);
        skip[l] = a.take<bf16>(px * f);
        if (l > 0) pooled[l] = a.take<bf16>(px * u->filters[l - 1]);
        if (l > 0 && vol) xyp[l] = a.take<bf16>(2 * px * u->filters[l - 1]);   // xy-pooled, full depth
        if (l < nl - 1) {
            up[l] = a.take<bf16>(px * f);
            if (u->bridge >= SQ_BRIDGE_ADD && u->bridge <= SQ_BRIDGE_SUB) merged[l] = a.take<bf16>(px * f);
            ut[l] = a.take<bf16>(px * f);
            uo[l] = a.take<bf16>(px * f);
        }
    }
    if (need) *need = a.off;
    if (dry) return SQ_OK;
    SQ_REQUIRE(a.ok(), SQ_ENOMEM, "unet(bf16): workspace %zu < %zu bytes", ws_bytes, a.off);

    u->last_launches = 0;
    sq_timer_mark(u, st, nullptr, 0);
    const int threads = 256;
    char scope[64];
    std::vector<char> pool_fused(nl + 1, 0);
    // level 0 on the quad (space-to-depth) layout: planar nets with 16 filters at level 0, even frame sizes and
    // a fused head; t1[0], skip[0], up[0], merged[0], ut[0] are then quad tensors (same sizes)
    bool quad = !vol && nl >= 2 && u->filters[0] == 16 && (hgt % 2 == 0) && (wid % 2 == 0) && u->cin <= 4 &&
                u->nout >= 2 && u->nout <= 4 && ((uintptr_t)mask % 2 == 0) && quad_enabled() &&
                !first_fusion_enabled() && !pair_enabled();
    for (const SqLayer &L : u->layers)
        if (L.level == 0 && L.kind != SqLayer::HEAD && !(L.kind == SqLayer::CONV && L.cin0 == u->cin && L.scope == "UNet/down0/conv1") &&
            !L.w_qd)
            quad = false;
    const TcGeo geoq = {n, 1, hgt / 2, wid / 2, 1, 1, 0, 0};
    for (int l = 0; l < nl; ++l) {
        const int D = depth(l), H = hgt >> l, W = wid >> l;
        const long long px = (long long)n * D * H * W;
        const TcGeo geo = {n, D, H, W, KZ, 1, 0, 0};
        snprintf(scope, sizeof scope, "UNet/down%d/conv1", l);
        SqLayer *c1 = layer_by_scope(u, scope);
        snprintf(scope, sizeof scope, "UNet/down%d/conv2", l);
        SqLayer *c2 = layer_by_scope(u, scope);
        // level-0 fusion: Cin = 1 -> 16 -> 16 channels on planar stacks with the pooled copy fused (SQ_FUSE_FIRST=0: off)
        const bool fuse_first = l == 0 && !vol && u->cin == 1 && c1->cout == 16 && c2->cout == 16 && nl > 1 &&
                                first_fusion_enabled();
        // quad layout: both convs of level 0 in one launch (the first conv's output stays on chip)
        const bool qfuse = l == 0 && quad && c2->w_qf && (wid % (raw16 ? 8 : 4) == 0) &&
                           ((uintptr_t)(raw16 ? (const void *)raw16 : (const void *)in) % 16 == 0) && qfuse_enabled();
        if (l == 0) SQ_REQUIRE(!raw16 || qfuse, SQ_ESTATE, "unet(bf16): raw uint16 frames need the fused first pair");
        if (l == 0 && vol) {
            const float *wf = (const float *)c1->w_tc;
            const int key = u->cin * 1000 + c1->cout;
            const int fr3 = first_rows(u->cin, 3);
            const dim3 g3((W + FIRST_TW - 1) / FIRST_TW, (H + fr3 - 1) / fr3, n * D);
            SQ_REQUIRE((long long)n * D <= 65535, SQ_EINVAL, "unet(bf16): more than 65535 slices per call");
#define SQ_FIRST3M(CI, CO) first_conv_kernel<CI, CO, 3><<<g3, 256, 0, st>>>(in, wf, c1->scale, c1->shift, t1[0], n, D, H, W)
            if (key == 1016) SQ_FIRST3M(1, 16);
            else if (key == 1032) SQ_FIRST3M(1, 32);
            else if (key == 2016) SQ_FIRST3M(2, 16);
            else {
                // other channel counts: one thread per voxel on CUDA cores
                const size_t sm = (size_t)(27 * u->cin + 2) * c1->cout * sizeof(float);
                const unsigned grid = (unsigned)((px + 127) / 128);
#define SQ_FIRST3(CO) first_conv3d_kernel<CO><<<grid, 128, sm, st>>>(in, wf, c1->scale, c1->shift, t1[0], n, D, H, W, u->cin)
                switch (c1->cout) {
                case 16: SQ_FIRST3(16); break;
                case 32: SQ_FIRST3(32); break;
                case 64: SQ_FIRST3(64); break;
                default:
                    SQ_REQUIRE(false, SQ_EUNSUPPORTED, "bf16 mode: first 3-D conv -> %d channels not instantiated",
                               c1->cout);
                }
#undef SQ_FIRST3
            }
#undef SQ_FIRST3M
            ++u->last_launches;
            SQ_CHECK_LAUNCH();
        } else if (l == 0 && (fuse_first || qfuse)) {
            // down0/conv1 runs inside down0/conv2's launch (launch_qf / launch_tc_first below)
        } else if (l == 0) {
            const int fr = first_rows(u->cin, 1);
            const dim3 grid((W + FIRST_TW - 1) / FIRST_TW, (H + fr - 1) / fr, n);
            const float *wf = (const float *)c1->w_tc;
#define SQ_FIRST(CI, CO) first_conv_kernel<CI, CO, 1><<<grid, 256, 0, st>>>(in, wf, c1->scale, c1->shift, t1[0], n, 1, H, W)
#define SQ_FIRSTQ(CI) first_conv_kernel<CI, 16, 1, true><<<grid, 256, 0, st>>>(in, wf, c1->scale, c1->shift, t1[0], n, 1, H, W)
            const int key = quad ? u->cin : u->cin * 1000 + c1->cout;
            switch (key) {
            case 1: SQ_FIRSTQ(1); break;
            case 2: SQ_FIRSTQ(2); break;
            case 3: SQ_FIRSTQ(3); break;
            case 4: SQ_FIRSTQ(4); break;
            case 1016: SQ_FIRST(1, 16); break;
            case 1032: SQ_FIRST(1, 32); break;
            case 1064: SQ_FIRST(1, 64); break;
            case 2016: SQ_FIRST(2, 16); break;
            case 2032: SQ_FIRST(2, 32); break;
            case 3016: SQ_FIRST(3, 16); break;
            case 3032: SQ_FIRST(3, 32); break;
            case 3064: SQ_FIRST(3, 64); break;
            case 4016: SQ_FIRST(4, 16); break;
            case 4032: SQ_FIRST(4, 32); break;
            default:
                SQ_REQUIRE(false, SQ_EUNSUPPORTED, "bf16 mode: first conv %d -> %d channels not instantiated",
                           u->cin, c1->cout);
            }
#undef SQ_FIRST
#undef SQ_FIRSTQ
            ++u->last_launches;
            SQ_CHECK_LAUNCH();
        } else {
            const int fp = u->filters[l - 1];
            if (!pool_fused[l]) {
                // xy pool over every slice of the level above (2*D of them in a volume)
                const long long nvec = px * (vol ? 2 : 1) * (fp / 8);
                maxpool_bf16_kernel<<<(unsigned)((nvec + threads - 1) / threads), threads, 0, st>>>(
                    (const uint4 *)skip[l - 1], (uint4 *)(vol ? xyp[l] : pooled[l]), nvec, H, W);
                ++u->last_launches;
                SQ_CHECK_LAUNCH();
            }
            if (vol) {
                const long long nvec = px * (fp / 8), plane = (long long)(fp / 8) * H * W;
                zpool_bf16_kernel<<<(unsigned)((nvec + threads - 1) / threads), threads, 0, st>>>(
                    (const uint4 *)xyp[l], (uint4 *)pooled[l], nvec, plane);
                ++u->last_launches;
                SQ_CHECK_LAUNCH();
            }
            if (vol || !pool_fused[l]) sq_timer_mark(u, st, "maxpool", 0);
            SQ_TRY(conv3x3_tc(u, *c1, pooled[l], c1->cin0, nullptr, 0, t1[l], nullptr, nullptr, geo, st));
        }
        if (!qfuse) sq_timer_mark(u, st, c1->scope.c_str(), c1->flops_per_px * px);
        // the level's second conv also emits the (xy-)pooled tensor the next level starts from
        const bool fuse_pool = (l < nl - 1) && c2->cout <= 128;
        if (l < nl - 1) pool_fused[l + 1] = fuse_pool;
        bf16 *pool_dst = !fuse_pool ? nullptr : (vol ? xyp[l + 1] : pooled[l + 1]);
        if (qfuse) {
            SQ_TRY(launch_qf(u, *c1, *c2, in, raw16, stats16, skip[0], pool_dst, geoq, st));
            const char *pname = u->aux_names.emplace(c1->scope, c1->scope + "+conv2").first->second.c_str();
            sq_timer_mark(u, st, pname, (c1->flops_per_px + c2->flops_per_px) * px);
            continue;
        }
        if (l == 0 && quad) SQ_TRY(conv3x3_qd(u, *c2, t1[0], c2->cin0, nullptr, 0, skip[0], pool_dst, nullptr, geoq, st));
        else if (fuse_first && pool_dst) SQ_TRY(launch_tc_first(u, *c1, *c2, in, skip[l], pool_dst, geo, st));
        else SQ_TRY(conv3x3_tc(u, *c2, t1[l], c2->cin0, nullptr, 0, skip[l], pool_dst, nullptr, geo, st));
        sq_timer_mark(u, st, c2->scope.c_str(), c2->flops_per_px * px);
    }
    SqLayer *head = layer_by_scope(u, "UNet/to_image");
    const bool head_fused = nl >= 2 && head->cout >= 2 && head->cout <= 4 && u->filters[0] <= 32;
    const bf16 *cur = skip[nl - 1];
    for (int l = nl - 2; l >= 0; --l) {
        const int D = depth(l), H = hgt >> l, W = wid >> l;
        const long long px = (long long)n * D * H * W;
        const TcGeo geo = {n, D, H, W, KZ, 1, 0, 0};
        snprintf(scope, sizeof scope, "UNet/up%d/upscale", l);
        SqLayer *us = layer_by_scope(u, scope);
        snprintf(scope, sizeof scope, "UNet/up%d/conv1", l);
        SqLayer *c1 = layer_by_scope(u, scope);
        snprintf(scope, sizeof scope, "UNet/up%d/conv2", l);
        SqLayer *c2 = layer_by_scope(u, scope);
        if (vol) {
            // 2x2x2 transposed conv = one planar 2x2 up-conv per output depth parity kz: input slice
            // s = n*D/2 + z feeds output slice 2*s + kz with the weights of depth tap kz
            const size_t wslice = (size_t)4 * us->cin0 * us->cout;
            for (int kz = 0; kz < 2; ++kz) {
                const TcGeo gu = {n * (D / 2), 1, H / 2, W / 2, 1, 2, kz, kz * wslice};
                SQ_TRY(upconv_tc(u, *us, cur, up[l], gu, st));
            }
        } else if (l == 0 && quad && u->bridge == SQ_BRIDGE_CONCAT && c1->w_qu && (us->cin0 == 16 || us->cin0 == 32) &&
                   qup_enabled()) {
            // up-conv + conv1 in one launch: the up-sampled tensor stays on chip
            const HeadArgs hq = {(const float *)head->w_tc, head->cout, logits, probs, mask};
            SQ_TRY(launch_qu(u, *us, *c1, cur, skip[0], ut[0], geoq, st));
            const char *pname = u->aux_names.emplace(us->scope, us->scope + "+conv1").first->second.c_str();
            sq_timer_mark(u, st, pname, (us->flops_per_px + c1->flops_per_px) * px);
            SQ_TRY(conv3x3_qd(u, *c2, ut[0], c2->cin0, nullptr, 0, nullptr, nullptr, &hq, geoq, st));
            sq_timer_mark(u, st, c2->scope.c_str(), (c2->flops_per_px + head->flops_per_px) * px);
            return SQ_OK;
        } else if (l == 0 && quad) {
            // 2x2 stride-2 up-conv onto the quad layout = 1x1 conv cin -> 4 x 16 at half resolution
            const HeadArgs none = {nullptr, 0, nullptr, nullptr, nullptr};
            if ((us->cin0 / 16) % 2 == 0)
                SQ_TRY((launch_qd<1, 2, EPI_STORE, 0>(u, *us, cur, us->cin0 / 8, nullptr, 0, up[0], nullptr, none, geoq, 0, st)));
            else
                SQ_TRY((launch_qd<1, 1, EPI_STORE, 0>(u, *us, cur, us->cin0 / 8, nullptr, 0, up[0], nullptr, none, geoq, 0, st)));
        } else {
            const TcGeo gu = {n, 1, H / 2, W / 2, 1, 1, 0, 0};
            SQ_TRY(upconv_tc(u, *us, cur, up[l], gu, st));
        }
        sq_timer_mark(u, st, us->scope.c_str(), us->flops_per_px * px);
        const bf16 *in0 = up[l], *in1 = nullptr;
        if (u->bridge == SQ_BRIDGE_CONCAT) {
            in1 = skip[l];
        } else if (u->bridge != SQ_BRIDGE_NONE) {
            const long long n2 = px * us->cout / 2;
            eltwise_bf16_kernel<<<(unsigned)((n2 + threads - 1) / threads), threads, 0, st>>>(
                (const uint32_t *)up[l], (const uint32_t *)skip[l], n2, u->bridge, (uint32_t *)merged[l]);
            ++u->last_launches;
            SQ_CHECK_LAUNCH();
            sq_timer_mark(u, st, "bridge", 0);
            in0 = merged[l];
        }
        const bool last = (l == 0 && head_fused);
        const HeadArgs ha = {(const float *)head->w_tc, head->cout, logits, probs, mask};
        // conv_block as one launch where the fused pair kernel applies (the intermediate `ut` stays on chip)
        if (l == 0 && quad) {
            SQ_TRY(conv3x3_qd(u, *c1, in0, c1->cin0, in1, c1->cin1, ut[0], nullptr, nullptr, geoq, st));
            sq_timer_mark(u, st, c1->scope.c_str(), c1->flops_per_px * px);
            SQ_TRY(conv3x3_qd(u, *c2, ut[0], c2->cin0, nullptr, 0, nullptr, nullptr, &ha, geoq, st));
            sq_timer_mark(u, st, c2->scope.c_str(), (c2->flops_per_px + head->flops_per_px) * px);
            return SQ_OK;
        }
        {
            const int r = vol ? SQ_NOT_APPLICABLE
                              : conv_pair_tc(u, *c1, *c2, in0, c1->cin0, in1, c1->cin1, last ? nullptr : uo[l], nullptr,
                                             last ? &ha : nullptr, geo, st);
            if (r != SQ_NOT_APPLICABLE) {
                SQ_TRY(r);
                const char *pname = u->aux_names.emplace(c1->scope, c1->scope + "+conv2").first->second.c_str();
                sq_timer_mark(u, st, pname,
                              (c1->flops_per_px + c2->flops_per_px + (last ? head->flops_per_px : 0.0)) * px);
                if (last) return SQ_OK;
                cur = uo[l];
                continue;
            }
        }
        SQ_TRY(conv3x3_tc(u, *c1, in0, c1->cin0, in1, c1->cin1, ut[l], nullptr, nullptr, geo, st));
        sq_timer_mark(u, st, c1->scope.c_str(), c1->flops_per_px * px);
        if (last) {
            // last conv of the net: 1x1 head + softmax + argmax run in its epilogue
            SQ_TRY(conv3x3_tc(u, *c2, ut[l], c2->cin0, nullptr, 0, nullptr, nullptr, &ha, geo, st));
            sq_timer_mark(u, st, c2->scope.c_str(), (c2->flops_per_px + head->flops_per_px) * px);
            return SQ_OK;
        }
        SQ_TRY(conv3x3_tc(u, *c2, ut[l], c2->cin0, nullptr, 0, uo[l], nullptr, nullptr, geo, st));
        sq_timer_mark(u, st, c2->scope.c_str(), c2->flops_per_px * px);
        cur = uo[l];
    }
    const long long px0 = (long long)hgt * wid, nslices = (long long)n * depth(0);
    const size_t sm = (size_t)(head->cin0 * head->cout + head->cout) * sizeof(float);
    head_bf16_kernel<<<(unsigned)((px0 * nslices + 127) / 128), 128, sm, st>>>(
        cur, (const float *)head->w_tc, head->shift, head->cin0, head->cout, px0, (int)nslices, logits, probs,
        mask);
    ++u->last_launches;
    SQ_CHECK_LAUNCH();
    sq_timer_mark(u, st, head->scope.c_str(), head->flops_per_px * px0 * nslices);
    return SQ_OK;
}

}  // namespace

int sq_tc_workspace_bytes(sq_unet_s *u, int n, int d, int hgt, int wid, size_t *bytes)
{
    return tc_run(u, true, nullptr, n, d, hgt, wid, nullptr, nullptr, nullptr, nullptr, 0, nullptr, bytes);
}

bool sq_tc_can_take_raw_u16(sq_unet_s *u, int hgt, int wid)
{
    // the conditions under which tc_run routes level 0 through conv_qf_kernel (SQ_QNORM=0: keep the float32 stage)
    const char *e = getenv("SQ_QNORM");
    if (e && atoi(e) == 0) return false;
    if (u->mode != SQ_MODE_BF16_TC || u->ndim != 2 || u->nlev < 2 || u->filters[0] != 16 || u->cin != 1 || u->nout < 2 ||
        u->nout > 4 || (hgt % 2) || (wid % 8) || !quad_enabled() || !qfuse_enabled() || first_fusion_enabled() || pair_enabled())
        return false;
    for (const SqLayer &L : u->layers) {
        if (L.scope == "UNet/down0/conv2" && !L.w_qf) return false;
        if (L.level == 0 && L.kind != SqLayer::HEAD && L.scope != "UNet/down0/conv1" && !L.w_qd) return false;
    }
    return true;
}

int sq_tc_forward_raw_u16(sq_unet_s *u, const uint16_t *raw, const float2 *stats, int n, int hgt, int wid,
                          uint8_t *mask, void *ws, size_t ws_bytes, cudaStream_t st)
{
    SQ_REQUIRE(((uintptr_t)mask % 2) == 0, SQ_EINVAL, "unet(bf16): mask pointer must be 2-byte aligned");
    return tc_run(u, false, nullptr, n, 1, hgt, wid, nullptr, mask, nullptr, ws, ws_bytes, st, nullptr, raw, stats);
}

int sq_tc_forward(sq_unet_s *u, const float *in, int n, int d, int hgt, int wid, float *probs,
                  uint8_t *mask, float *logits, void *ws, size_t ws_bytes, cudaStream_t st)
{
    return tc_run(u, false, in, n, d, hgt, wid, probs, mask, logits, ws, ws_bytes, st, nullptr);
}
