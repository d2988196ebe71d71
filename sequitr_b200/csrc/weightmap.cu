// Per-pixel loss-weight maps on the GPU.
//
//   W1  sq_weightmap_edt   <- pipeline.ImageWeightMap.pipe (reference pipeline.py:475-479)
//         d = EDT(1 - m);  w = w0*(1-m)*exp(-d*d/(2 sigma^2 + 1e-99)) + m + 1
//   W3  sq_weightmap_unet  <- north-star formula on instance labels
//         w = wc[m] + w0*(1-m)*exp(-(d1+d2)^2/(2 sigma^2 + 1e-99))
//
// Distance transform: EXACT separable Euclidean transform on integer squared
// distances (so it is bit-identical to SciPy's distance_transform_edt, which a
// jump-flooding pass is not):
//   phase 1 (rows)    horizontal distance to the nearest seed in the same row --
//                     a segmented prefix/suffix scan per row, uint16 per pixel;
//   phase 2 (columns) d2(x,y) = min_dy dy^2 + g(x,y+dy)^2, scanned outwards from
//                     the pixel and stopped as soon as dy^2 >= best (exact), or
//                     at a cut-off radius beyond which w0*exp(..) is below half
//                     an ulp of the class term -- the stored result is still the
//                     one the exact transform would give after rounding.
// W3 carries the two nearest DISTINCT instance labels through both phases.
// Both kernels are HBM/L2-bound integer work: 1 B (W1) or 4 B (W3) in and one
// float out per pixel; the uint16 / 12-byte intermediates stay L2-resident.
#include "sq_common.cuh"
#include <cstdlib>
#include <cmath>

namespace {

constexpr unsigned short INF16 = 0xFFFF;
constexpr unsigned INF32 = 0xFFFFFFFFu;
constexpr int ROW_THREADS = 256;

// ------------------------------------------------------------------ W1 phase 1
// grid (hgt, n), block 256, dyn smem: wid * 2 bytes
__global__ void edt_rows(const uint8_t *__restrict__ mask, unsigned short *__restrict__ g,
                         int *__restrict__ anyfg, int hgt, int wid)
{
    extern __shared__ unsigned short srow[];
    __shared__ int lastp[ROW_THREADS], firstp[ROW_THREADS];
    const int t = threadIdx.x;
    const long long ro = ((long long)blockIdx.y * hgt + blockIdx.x) * wid;
    const uint8_t *mk = mask + ro;
    const int seg = (wid + ROW_THREADS - 1) / ROW_THREADS;
    const int x0 = min(t * seg, wid), x1 = min(x0 + seg, wid);
    int lp = -1, fp = 0x3fffffff;
    for (int x = x0; x < x1; ++x)
        if (mk[x]) { lp = x; if (fp == 0x3fffffff) fp = x; }
    if (lp >= 0) anyfg[blockIdx.y] = 1;
    lastp[t] = lp;
    firstp[t] = fp;
    __syncthreads();
    // inclusive prefix-max of lastp, inclusive suffix-min of firstp (Hillis-Steele)
    for (int o = 1; o < ROW_THREADS; o <<= 1) {
        const int a = (t >= o) ? lastp[t - o] : -1;
        const int b = (t + o < ROW_THREADS) ? firstp[t + o] : 0x3fffffff;
        __syncthreads();
        lastp[t] = max(lastp[t], a);
        firstp[t] = min(firstp[t], b);
        __syncthreads();
    }
    int last = (t > 0) ? lastp[t - 1] : -1;
    for (int x = x0; x < x1; ++x) {
        if (mk[x]) last = x;
        srow[x] = (last >= 0) ? (unsigned short)min(x - last, 0xFFFE) : INF16;
    }
    int next = (t + 1 < ROW_THREADS) ? firstp[t + 1] : 0x3fffffff;
    for (int x = x1 - 1; x >= x0; --x) {
        if (mk[x]) next = x;
        if (next != 0x3fffffff) srow[x] = min(srow[x], (unsigned short)min(next - x, 0xFFFE));
    }
    __syncthreads();
    for (int x = t; x < wid; x += ROW_THREADS) g[ro + x] = srow[x];
}

// ------------------------------------------------------------------ W1 phase 2
// grid (ceil(W/32), ceil(H/8), n), block (32,8)
template <typename OutT>
__global__ void edt_cols_weight(const uint8_t *__restrict__ mask,
                                const unsigned short *__restrict__ g,
                                const int *__restrict__ anyfg, OutT *__restrict__ out,
                                int *__restrict__ d2_out, int hgt, int wid, int rmax,
                                double w0e, double denom)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= wid || y >= hgt) return;
    const long long fo = (long long)blockIdx.z * hgt * wid;
    const long long idx = fo + (long long)y * wid + x;
    if (mask[idx]) {
        out[idx] = (OutT)2.0;                       // w0*(1-1)*exp(0) + 1 + 1
        if (d2_out) d2_out[idx] = 0;
        return;
    }
    unsigned best;
    if (!anyfg[blockIdx.z]) {
        // SciPy's distance_transform_edt with no seed: virtual seed at (-1, 0)
        best = (unsigned)(y + 1) * (unsigned)(y + 1) + (unsigned)x * (unsigned)x;
    } else {
        const unsigned short g0 = g[idx];
        best = (g0 == INF16) ? INF32 : (unsigned)g0 * g0;
        const int lim = min(rmax, max(y, hgt - 1 - y));
        for (int dy = 1; dy <= lim; ++dy) {
            const unsigned dy2 = (unsigned)dy * dy;
            if (dy2 >= best) break;
            if (y - dy >= 0) {
                const unsigned short gv = g[idx - (long long)dy * wid];
                if (gv != INF16) best = min(best, dy2 + (unsigned)gv * gv);
            }
            if (y + dy < hgt) {
                const unsigned short gv = g[idx + (long long)dy * wid];
                if (gv != INF16) best = min(best, dy2 + (unsigned)gv * gv);
            }
        }
    }
    if (d2_out) d2_out[idx] = (best == INF32) ? 0x7fffffff : (int)best;
    double w = 1.0;
    if (best != INF32) {
        const double d = sqrt((double)best);        // SciPy returns sqrt(d2) ...
        w = w0e * exp(-(d * d) / denom) + 0.0 + 1.0;  // ... and pipeline.py:477 squares it again
    }
    out[idx] = (OutT)w;
}

// ------------------------------------------------------------------ W3 phase 1
struct Two {                 // two most recent distinct labels and their x positions
    int l1, p1, l2, p2;      // label 0 = none
};

__device__ __forceinline__ Two two_push(Two s, int l, int p)
{
    if (l == s.l1) { s.p1 = p; return s; }
    s.l2 = s.l1; s.p2 = s.p1;
    s.l1 = l;    s.p1 = p;
    return s;
}

// older A followed by newer B
__device__ __forceinline__ Two two_combine(const Two &A, const Two &B)
{
    if (B.l1 == 0) return A;
    Two r = B;
    if (r.l2 == 0) {
        if (A.l1 != 0 && A.l1 != r.l1) { r.l2 = A.l1; r.p2 = A.p1; }
        else if (A.l2 != 0 && A.l2 != r.l1) { r.l2 = A.l2; r.p2 = A.p2; }
    }
    return r;
}

struct Best2 {               // two nearest distinct labels with (squared or plain) distances
    int la, lb;
    unsigned da, db;         // da <= db
};

__device__ __forceinline__ void best2_insert(Best2 &b, int l, unsigned d)
{
    if (l == 0) return;
    if (l == b.la) { b.da = min(b.da, d); return; }
    if (l == b.lb) {
        b.db = min(b.db, d);
        if (b.db < b.da) { int tl = b.la; unsigned td = b.da; b.la = b.lb; b.da = b.db; b.lb = tl; b.db = td; }
        return;
    }
    if (d < b.da) { b.lb = b.la; b.db = b.da; b.la = l; b.da = d; }
    else if (d < b.db) { b.lb = l; b.db = d; }
}

// grid (hgt, n), block 256
__global__ void inst_rows(const int *__restrict__ labels, int *__restrict__ la,
                          int *__restrict__ lb, unsigned short *__restrict__ da,
                          unsigned short *__restrict__ db, int hgt, int wid)
{
    __shared__ Two sc[ROW_THREADS];
    const int t = threadIdx.x;
    const long long ro = ((long long)blockIdx.y * hgt + blockIdx.x) * wid;
    const int *lr = labels + ro;
    const int seg = (wid + ROW_THREADS - 1) / ROW_THREADS;
    const int x0 = min(t * seg, wid), x1 = min(x0 + seg, wid);
    const Two empty = {0, 0, 0, 0};

    // ---- left-to-right
    Two s = empty;
    for (int x = x0; x < x1; ++x) { const int l = lr[x]; if (l > 0) s = two_push(s, l, x); }
    sc[t] = s;
    __syncthreads();
    for (int o = 1; o < ROW_THREADS; o <<= 1) {
        Two a = empty;
        if (t >= o) a = sc[t - o];
        __syncthreads();
        if (t >= o) sc[t] = two_combine(a, sc[t]);
        __syncthreads();
    }
    s = (t > 0) ? sc[t - 1] : empty;
    for (int x = x0; x < x1; ++x) {
        const int l = lr[x];
        if (l > 0) s = two_push(s, l, x);
        Best2 b = {0, 0, INF32, INF32};
        if (s.l1) best2_insert(b, s.l1, (unsigned)(x - s.p1));
        if (s.l2) best2_insert(b, s.l2, (unsigned)(x - s.p2));
        la[ro + x] = b.la; lb[ro + x] = b.lb;
        da[ro + x] = (unsigned short)min(b.da, 0xFFFEu + (b.la == 0));
        db[ro + x] = (unsigned short)min(b.db, 0xFFFEu + (b.lb == 0));
    }
    __syncthreads();

    // ---- right-to-left ("newer" = smaller x)
    s = empty;
    for (int x = x1 - 1; x >= x0; --x) { const int l = lr[x]; if (l > 0) s = two_push(s, l, x); }
    sc[t] = s;
    __syncthreads();
    for (int o = 1; o < ROW_THREADS; o <<= 1) {
        Two a = empty;
        if (t + o < ROW_THREADS) a = sc[t + o];
        __syncthreads();
        if (t + o < ROW_THREADS) sc[t] = two_combine(a, sc[t]);
        __syncthreads();
    }
    s = (t + 1 < ROW_THREADS) ? sc[t + 1] : empty;
    for (int x = x1 - 1; x >= x0; --x) {
        const int l = lr[x];
        if (l > 0) s = two_push(s, l, x);
        Best2 b;
        b.la = la[ro + x]; b.lb = lb[ro + x];
        b.da = b.la ? da[ro + x] : INF32;
        b.db = b.lb ? db[ro + x] : INF32;
        if (s.l1) best2_insert(b, s.l1, (unsigned)(s.p1 - x));
        if (s.l2) best2_insert(b, s.l2, (unsigned)(s.p2 - x));
        la[ro + x] = b.la; lb[ro + x] = b.lb;
        da[ro + x] = b.la ? (unsigned short)min(b.da, 0xFFFEu) : INF16;
        db[ro + x] = b.lb ? (unsigned short)min(b.db, 0xFFFEu) : INF16;
    }
}

// ------------------------------------------------------------------ W3 phase 2
template <typename OutT>
__global__ void inst_cols_weight(const int *__restrict__ labels, const int *__restrict__ la,
                                 const int *__restrict__ lb, const unsigned short *__restrict__ da,
                                 const unsigned short *__restrict__ db, OutT *__restrict__ out,
                                 int hgt, int wid, int rmax, double w0, double denom,
                                 double wc0, double wc1)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= wid || y >= hgt) return;
    const long long fo = (long long)blockIdx.z * hgt * wid;
    const long long idx = fo + (long long)y * wid + x;
    if (labels[idx] > 0) { out[idx] = (OutT)wc1; return; }
    Best2 b = {0, 0, INF32, INF32};
    {
        const int l1 = la[idx], l2 = lb[idx];
        if (l1) { const unsigned d = da[idx]; best2_insert(b, l1, d * d); }
        if (l2) { const unsigned d = db[idx]; best2_insert(b, l2, d * d); }
    }
    const int lim = min(rmax, max(y, hgt - 1 - y));
    for (int dy = 1; dy <= lim; ++dy) {
        const unsigned dy2 = (unsigned)dy * dy;
        if (dy2 >= b.db) break;
#pragma unroll
        for (int sgn = -1; sgn <= 1; sgn += 2) {
            const int yy = y + sgn * dy;
            if (yy < 0 || yy >= hgt) continue;
            const long long j = idx + (long long)sgn * dy * wid;
            const int l1 = la[j];
            if (l1) {
                const unsigned d = da[j];
                best2_insert(b, l1, dy2 + d * d);
                const int l2 = lb[j];
                if (l2) { const unsigned e = db[j]; best2_insert(b, l2, dy2 + e * e); }
            }
        }
    }
    double w = wc0;
    if (b.lb != 0) {
        const double s = sqrt((double)b.da) + sqrt((double)b.db);
        w = wc0 + w0 * exp(-(s * s) / denom);
    }
    out[idx] = (OutT)w;
}

// ------------------------------------------------------------------ cut-off (bounded) kernels
// When the cut-off radius R (beyond which the exponential term cannot change the rounded
// result) is small -- R = 32 for the reference's w0 = 10, sigma = 5 -- distances only matter up
// to R, fit in a byte, and both phases become cheap:
//   row pass   every row is turned into bit masks in shared memory (one ballot per 32 pixels);
//              a pixel finds its nearest seed (W1) or its two nearest distinct labels (W3, via
//              run-start / run-end masks) with a handful of clz/ffs on <= 4 words -- O(1) per
//              pixel, fully parallel, 1 B/px (W1) or 10 B/px (W3) of L2-resident output;
//   col pass   a block stages the (64 + 2R) x 32 window of that output in shared memory and
//              scans +-dy with the exact early exit dy^2 >= best.  W1 weights depend on the
//              integer d2 only, so the fp64 sqrt/exp is evaluated once per distinct value
//              (<= R^2 + 1 of them, same expression as the general kernel) and looked up.
constexpr int WR_MAX = 64;             // largest cut-off radius handled by the bounded kernels
constexpr unsigned char INF8 = 255;
constexpr int CT_W = 32, CT_H = 64;    // column-pass tile

// highest set bit at index <= pos and >= lo in a word array (-1 if none)
__device__ __forceinline__ int prev_set(const unsigned *w, int pos, int lo)
{
    if (pos < lo) return -1;
    int wi = pos >> 5;
    unsigned m = w[wi] & (0xffffffffu >> (31 - (pos & 31)));
    const int wlo = lo >> 5;
    while (true) {
        if (m) {
            const int r = (wi << 5) + 31 - __clz(m);
            return r >= lo ? r : -1;
        }
        if (--wi < wlo) return -1;
        m = w[wi];
    }
}
// lowest set bit at index >= pos and <= hi (-1 if none)
__device__ __forceinline__ int next_set(const unsigned *w, int pos, int hi)
{
    if (pos > hi) return -1;
    int wi = pos >> 5;
    unsigned m = w[wi] & (0xffffffffu << (pos & 31));
    const int whi = hi >> 5;
    while (true) {
        if (m) {
            const int r = (wi << 5) + __ffs(m) - 1;
            return r <= hi ? r : -1;
        }
        if (++wi > whi) return -1;
        m = w[wi];
    }
}

// grid (hgt, n), block 256, dyn smem: (ceil(wid/32) + 4) words.  g8 = row distance if <= R else 255.
// Each pixel builds the 64-bit windows [x-63, x] and [x, x+63] of the row's foreground bit mask (two
// zero words pad each end) and takes one clz / ffs: O(1), no loops (R <= WR_MAX = 64).
__global__ void edt_rows_bits(const uint8_t *__restrict__ mask, unsigned char *__restrict__ g8,
                              int *__restrict__ anyfg, int hgt, int wid, int R)
{
    extern __shared__ unsigned bits_raw[];
    unsigned *bits = bits_raw + 2;                       // bits[-2..-1] and bits[nw..nw+1] are zero
    const long long ro = ((long long)blockIdx.y * hgt + blockIdx.x) * wid;
    const uint8_t *mk = mask + ro;
    const int nw = (wid + 31) >> 5;
    if (threadIdx.x < 2) { bits_raw[threadIdx.x] = 0u; bits[nw + threadIdx.x] = 0u; }
    bool any = false;
    for (int x = threadIdx.x; x < nw * 32; x += blockDim.x) {
        const unsigned b = __ballot_sync(0xffffffffu, x < wid && mk[x] != 0);
        if ((threadIdx.x & 31) == 0) bits[x >> 5] = b;
        any |= b != 0;
    }
    if (any && (threadIdx.x & 31) == 0) anyfg[blockIdx.y] = 1;
    __syncthreads();
    for (int x = threadIdx.x; x < wid; x += blockDim.x) {
        const int wi = x >> 5, sh = x & 31;
        const unsigned wm2 = bits[wi - 2], wm1 = bits[wi - 1], w0 = bits[wi], wp1 = bits[wi + 1], wp2 = bits[wi + 2];
        // left window: bit 63 = pixel x, bit 63-k = pixel x-k
        const unsigned lhi = __funnelshift_l(wm1, w0, 31 - sh);          // pixels x-31 .. x
        const unsigned llo = __funnelshift_l(wm2, wm1, 31 - sh);         // pixels x-63 .. x-32
        const unsigned long long L = ((unsigned long long)lhi << 32) | llo;
        // right window: bit k = pixel x+k
        const unsigned rlo = __funnelshift_r(w0, wp1, sh);               // pixels x .. x+31
        const unsigned rhi = __funnelshift_r(wp1, wp2, sh);              // pixels x+32 .. x+63
        const unsigned long long Rw = ((unsigned long long)rhi << 32) | rlo;
        int d = 1 << 20;
        if (L) d = __clzll((long long)L);
        if (Rw) d = min(d, __ffsll((long long)Rw) - 1);
        g8[ro + x] = (unsigned char)(d <= R ? d : INF8);
    }
}

// Same row pass, 16 pixels per thread (rows whose width is a multiple of 16 and 16-byte aligned): one
// 16-byte mask load -> 16 mask bits (SIMD byte compare + multiply gather), one 16-byte store of the
// 16 distances; the five bit-mask words around the chunk are read from shared memory once.
// grid (hgt, n), block = 32 * ceil(wid / 512) threads (<= 256), dyn smem: (ceil(wid/32) + 4) words
__global__ void edt_rows_bits16(const uint8_t *__restrict__ mask, unsigned char *__restrict__ g8,
                                int *__restrict__ anyfg, int hgt, int wid, int R)
{
    extern __shared__ unsigned bits_raw[];
    unsigned *bits = bits_raw + 2;
    unsigned short *bits16 = reinterpret_cast<unsigned short *>(bits);
    const long long ro = ((long long)blockIdx.y * hgt + blockIdx.x) * wid;
    const int nw = (wid + 31) >> 5, nchunk = wid >> 4;
    if (threadIdx.x < 2) { bits_raw[threadIdx.x] = 0u; bits[nw + threadIdx.x] = 0u; }
    if ((wid & 31) && threadIdx.x == 2) bits[nw - 1] = 0u;        // upper half of the last word (wid % 32 == 16)
    __syncthreads();
    unsigned any = 0;
    for (int c = threadIdx.x; c < nchunk; c += blockDim.x) {
        const uint4 m = __ldg(reinterpret_cast<const uint4 *>(mask + ro) + c);
        const unsigned w4[4] = {m.x, m.y, m.z, m.w};
        unsigned b16 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned nz = __vcmpne4(w4[k], 0u) & 0x01010101u;       // 1 per non-zero byte
            b16 |= ((nz * 0x10204080u) >> 28) << (4 * k);                 // bytes 0..3 -> bits 0..3
        }
        bits16[c] = (unsigned short)b16;
        any |= b16;
    }
    if (__any_sync(0xffffffffu, any != 0) && (threadIdx.x & 31) == 0) anyfg[blockIdx.y] = 1;
    __syncthreads();
    for (int c = threadIdx.x; c < nchunk; c += blockDim.x) {
        const int x0 = c << 4, wi = x0 >> 5, sh = x0 & 31;                 // sh = 0 or 16: a chunk never straddles a word
        const unsigned wm2 = bits[wi - 2], wm1 = bits[wi - 1], w0 = bits[wi], wp1 = bits[wi + 1], wp2 = bits[wi + 2];
        // distance of pixel x0-1 to the nearest seed at or left of it, of pixel x0+16 to the nearest at or right of
        // it (64-pixel windows; R <= WR_MAX = 64), then one sweep in each direction over the chunk's 16 bits:
        // ~10 instructions per pixel instead of two 64-bit window extractions + clz/ffs per pixel
        const unsigned lhi = sh ? __funnelshift_l(wm1, w0, 16) : wm1, llo = sh ? __funnelshift_l(wm2, wm1, 16) : wm2;
        const unsigned rlo = sh ? wp1 : __funnelshift_r(w0, wp1, 16), rhi = sh ? wp2 : __funnelshift_r(wp1, wp2, 16);
        const unsigned long long L = ((unsigned long long)lhi << 32) | llo;      // bit 63 = pixel x0-1
        const unsigned long long Rw = ((unsigned long long)rhi << 32) | rlo;     // bit 0  = pixel x0+16
        const unsigned b16 = (w0 >> sh) & 0xffffu;
        unsigned dl = L ? (unsigned)__clzll((long long)L) : 1000u;               // 0 if pixel x0-1 is a seed
        unsigned dlv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            dl = (b16 & (1u << j)) ? 0u : dl + 1u;
            dlv[j] = dl;
        }
        unsigned dr = Rw ? (unsigned)(__ffsll((long long)Rw) - 1) : 1000u;       // 0 if pixel x0+16 is a seed
        unsigned o[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 15; j >= 0; --j) {
            dr = (b16 & (1u << j)) ? 0u : dr + 1u;
            const unsigned d = min(dlv[j], dr);
            o[j >> 2] |= (d <= (unsigned)R ? d : (unsigned)INF8) << (8 * (j & 3));
        }
        reinterpret_cast<uint4 *>(g8 + ro)[c] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void w1_table_kernel(double *__restrict__ table, int n, double w0e, double denom)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = sqrt((double)i);
    table[i] = w0e * exp(-(d * d) / denom) + 0.0 + 1.0;
    if (i == n - 1) table[n] = 1.0;                      // one past the cut-off: the class weight alone
}

// W1 column pass: out(y, x) = min over dy of dy^2 + g(y + dy, x)^2, brute force over |dy| <= RMAX with the
// packed add-min of the DPX unit: ONE instruction (VIADDMNMX.U16x2: min(a + b, c) per 16-bit half) updates
// the running minimum of two adjacent columns for one dy, and with the loops fully unrolled dy^2 is an
// immediate operand.  A thread owns 2 columns x DP_JR consecutive rows; a window row loaded once from
// shared memory (one LDS.32 = 2 columns) feeds all DP_JR outputs.  (2 RMAX + 1) / 2 instructions per pixel
// (36.5 at RMAX = 36), no branches, no divergence -- the block-pruned scan it replaces spent ~360.
// grid (ceil(W/128), ceil(H/32), n), block 256 = 64 column pairs x 4 row groups; static smem:
// (32 + 2 RMAX) x 128 squared row distances (uint16; 0x7FFF = no seed within R, which cannot wrap:
// 0x7FFF + RMAX^2 < 65536).  Every candidate is a real pixel, so the minimum is exact whenever it is
// <= R^2 (rows up to RMAX >= R away are scanned; what they add is > R^2); beyond R the weight is exactly 1.
constexpr int DP_TW = 128, DP_TH = 32, DP_JR = 8;
constexpr unsigned DP_INF = 0x7FFFu;

// frame without any seed: SciPy's transform then measures to a virtual seed (reference behaviour kept by
// the general kernel too); cold path, kept out of line
__device__ __noinline__ double seedless_weight(const double *__restrict__ table, int y, int x, unsigned R2, double w0e,
                                               double denom)
{
    const unsigned b = (unsigned)(y + 1) * (unsigned)(y + 1) + (unsigned)x * (unsigned)x;
    if (b <= R2) return __ldg(table + b);
    const double d = sqrt((double)b);
    return w0e * exp(-(d * d) / denom) + 0.0 + 1.0;
}

template <typename OutT, int RMAX>
__global__ void __launch_bounds__(256)
edt_cols_dpx(const unsigned char *__restrict__ g8, const int *__restrict__ anyfg,
             const double *__restrict__ table, OutT *__restrict__ out, int hgt, int wid, int R,
             double w0e, double denom)
{
    constexpr int ROWS = DP_TH + 2 * RMAX;
    constexpr bool TAB = RMAX <= 36;                     // the look-up table fits next to the window
    __shared__ __align__(16) unsigned short s2[ROWS * DP_TW];
    __shared__ OutT tab[TAB ? RMAX * RMAX + 2 : 1];
    if constexpr (TAB)
        for (int i = threadIdx.x; i < R * R + 2; i += 256) tab[i] = (OutT)__ldg(table + i);
    const int x0 = blockIdx.x * DP_TW, y0 = blockIdx.y * DP_TH;
    const long long fo = (long long)blockIdx.z * hgt * wid;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool none = !anyfg[blockIdx.z];
    // stage: one warp per window row, four columns per lane (unrolled: several rows' loads in flight)
#pragma unroll 4
    for (int r = warp; r < ROWS; r += 8) {
        const int yy = y0 - RMAX + r, xx = x0 + 4 * lane;
        unsigned v[4] = {DP_INF, DP_INF, DP_INF, DP_INF};
        if (yy >= 0 && yy < hgt) {
            const unsigned char *row = g8 + fo + (long long)yy * wid;
            if (xx + 3 < wid && (wid & 3) == 0) {
                const unsigned u = __ldg(reinterpret_cast<const unsigned *>(row + xx));
#pragma unroll
                for (int k = 0; k < 4; ++k) { const unsigned g = (u >> (8 * k)) & 0xffu; if (g != INF8) v[k] = g * g; }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (xx + k < wid) { const unsigned g = row[xx + k]; if (g != INF8) v[k] = g * g; }
            }
        }
        *reinterpret_cast<uint2 *>(s2 + r * DP_TW + 4 * lane) = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
    }
    __syncthreads();
    const int pair = (warp & 1) * 32 + lane, grp = warp >> 1;          // columns x0 + 2 pair, +1; rows y0 + 8 grp ..
    const unsigned *w32 = reinterpret_cast<const unsigned *>(s2) + (grp * DP_JR) * (DP_TW / 2) + pair;
    unsigned best[DP_JR];
#pragma unroll
    for (int j = 0; j < DP_JR; ++j) best[j] = DP_INF * 0x10001u;
#pragma unroll
    for (int r = 0; r < DP_JR + 2 * RMAX; ++r) {                     // window row y0 + 8 grp - RMAX + r
        const unsigned g = w32[r * (DP_TW / 2)];
#pragma unroll
        for (int j = 0; j < DP_JR; ++j) {
            const int dy = r - RMAX - j;                                 // a constant after unrolling
            if (dy >= -RMAX && dy <= RMAX) best[j] = __viaddmin_u16x2(g, (unsigned)(dy * dy) * 0x10001u, best[j]);
        }
    }
    const unsigned R2 = (unsigned)R * R;
    const int x = x0 + 2 * pair;
    if (x >= wid) return;
    OutT *o = out + fo + (long long)(y0 + grp * DP_JR) * wid + x;
    const bool both = x + 1 < wid && (wid & 1) == 0;
    if (none) {                                                          // cold: frame without any seed
        for (int j = 0; j < DP_JR && y0 + grp * DP_JR + j < hgt; ++j, o += wid)
            for (int e = 0; e < 2 && x + e < wid; ++e) {
                const unsigned b = (best[j] >> (16 * e)) & 0xffffu;
                o[e] = b == 0 ? (OutT)2.0 : (OutT)seedless_weight(table, y0 + grp * DP_JR + j, x + e, R2, w0e, denom);
            }
        return;
    }
#pragma unroll
    for (int j = 0; j < DP_JR; ++j, o += wid) {
        if (y0 + grp * DP_JR + j >= hgt) break;
        OutT w[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            // table[d2] for d2 <= R^2, table[R^2 + 1] = 1 (beyond the cut-off); foreground (distance 0) = 2
            const unsigned b = (best[j] >> (16 * e)) & 0xffffu;
            OutT v;
            if constexpr (TAB) v = tab[min(b, R2 + 1u)];
            else v = (OutT)__ldg(table + min(b, R2 + 1u));
            w[e] = b == 0 ? (OutT)2.0 : v;
        }
        if (both) {
            if constexpr (sizeof(OutT) == 4) *reinterpret_cast<float2 *>(o) = make_float2((float)w[0], (float)w[1]);
            else *reinterpret_cast<double2 *>(o) = make_double2((double)w[0], (double)w[1]);
        } else {
            o[0] = w[0];
            if (x + 1 < wid) o[1] = w[1];
        }
    }
}

// W3 column pass on the same DPX add-min, 32-bit: a candidate is the key (squared distance << 18 | label), so
// the minimum carries the label of the nearest instance along.  Sweep 1: k1 = min over rows of keyA + dy^2
// (keyA = the row's nearest label) -> d1^2 and its label L1.  Sweep 2: the nearest pixel of a row whose label
// differs from L1 is the row's nearest label if that is not L1, else its second nearest (the row pass keeps the
// two nearest DISTINCT labels), so k2 = min over rows of (labelA != L1 ? keyA : keyB) + dy^2 -> d2^2: a compare,
// a select and one VIADDMNMX.U32 per (pixel, dy), no branches, no divergence; ~4 instructions per (pixel, dy)
// in total against ~16 for the block-pruned two-minimum scan (inst_cols_tile, kept for labels >= 2^18).
// A thread owns one column and sweeps IP_JR = 4 rows at a time (longer batches double the unrolled code, which
// already misses the instruction cache -- ncu: 70 % hits -- and made ptxas hoist the window loads until registers
// spilled; the loads are volatile for the same reason); grid (ceil(W/64), ceil(H/64), n), block 256 = 64 columns
// x 4 row groups of 16 rows; dyn smem 2 x (64 + 2 RMAX) x 64 keys.  Rows / labels beyond R hold IP_INF (cannot wrap:
// IP_INF + RMAX^2 << 18 < 2^32); exact for the same reason as the W1 pass.
constexpr int IP_TW = 64, IP_TH = 64, IP_JR = 4;
constexpr unsigned IP_LBITS = 18, IP_LMASK = (1u << IP_LBITS) - 1u, IP_INFC = 12000u, IP_INF = IP_INFC << IP_LBITS;

template <typename OutT, int RMAX>
__global__ void __launch_bounds__(256, 3)      // three blocks per SM: one stages its window while the others sweep
inst_cols_dpx(const int *__restrict__ la, const int *__restrict__ lb, const unsigned char *__restrict__ da,
              const unsigned char *__restrict__ db, const int *__restrict__ big, OutT *__restrict__ out, int hgt,
              int wid, int R, double w0, double denom, double wc0, double wc1)
{
    if (big[blockIdx.z]) return;                       // labels >= 2^18 in this frame: inst_cols_tile does it
    constexpr int ROWS = IP_TH + 2 * RMAX, GROUPS = 256 / IP_TW, PER_GROUP = IP_TH / GROUPS;
    extern __shared__ __align__(16) unsigned keys_[];
    unsigned *kA = keys_, *kB = keys_ + ROWS * IP_TW;
    const int x0 = blockIdx.x * IP_TW, y0 = blockIdx.y * IP_TH;
    const long long fo = (long long)blockIdx.z * hgt * wid;
    const int col = threadIdx.x & (IP_TW - 1), grp = threadIdx.x / IP_TW, x = x0 + col;
    // ROWS / GROUPS rows per thread, four rows (16 loads) in flight at a time: issued one row per
    // iteration the staging was a chain of L2 round trips
    for (int rb = grp; rb < ROWS; rb += 4 * GROUPS) {
        unsigned d1[4], d2[4], l1v[4], l2v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = rb + k * GROUPS, yy = y0 - RMAX + r;
            const bool ok = r < ROWS && yy >= 0 && yy < hgt && x < wid;
            const long long j = ok ? fo + (long long)yy * wid + x : 0;
            d1[k] = ok ? (unsigned)__ldg(da + j) : (unsigned)INF8;
            d2[k] = ok ? (unsigned)__ldg(db + j) : (unsigned)INF8;
            l1v[k] = ok ? (unsigned)__ldg(la + j) : 0u;
            l2v[k] = ok ? (unsigned)__ldg(lb + j) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = rb + k * GROUPS;
            if (r < ROWS) {
                kA[r * IP_TW + col] = d1[k] != INF8 ? ((d1[k] * d1[k]) << IP_LBITS) | l1v[k] : IP_INF;
                kB[r * IP_TW + col] = d2[k] != INF8 ? ((d2[k] * d2[k]) << IP_LBITS) | l2v[k] : IP_INF;
            }
        }
    }
    __syncthreads();
    const double reach = (double)(R + 1);
    const unsigned reach2 = (unsigned)((R + 1) * (R + 1));
#pragma unroll 1
    for (int sub = 0; sub < PER_GROUP / IP_JR; ++sub) {                   // IP_JR rows at a time
        const int ly = grp * PER_GROUP + sub * IP_JR;                    // first output row of this batch (in the tile)
        // volatile: keeps ptxas from hoisting dozens of window loads ahead of the sweep (which spilled registers)
        const volatile unsigned *pa = kA + ly * IP_TW + col, *pb = kB + ly * IP_TW + col;
        unsigned k1[IP_JR], k2[IP_JR], l1[IP_JR];
#pragma unroll
        for (int j = 0; j < IP_JR; ++j) k1[j] = k2[j] = 0xffffffffu;
        // Sweep 1 only has to be exact up to reach/2: d2 >= d1, so a pixel whose nearest label is further than
        // that keeps the class weight whatever d1 and d2 are, and a nearest label within reach/2 lies within
        // reach/2 ROWS.  Rows |dy| <= RH = RMAX/2 + 1 therefore give the exact d1 (and its label) for every pixel
        // that can matter and a value > reach^2/4 for all others: 39 add-mins per pixel instead of 73.
        constexpr int RH = RMAX / 2 + 1;
#pragma unroll
        for (int r = RMAX - RH; r < IP_JR + RMAX + RH; ++r) {            // window row y0 + ly - RMAX + r
            const unsigned g = pa[r * IP_TW];
#pragma unroll
            for (int j = 0; j < IP_JR; ++j) {
                const int dy = r - RMAX - j;                                 // a constant after unrolling
                if (dy >= -RH && dy <= RH) k1[j] = __viaddmin_u32(g, (unsigned)(dy * dy) << IP_LBITS, k1[j]);
            }
            if ((r & 7) == 7) asm volatile("" ::: "memory");     // keep the loads of later rows from piling up in registers
        }
        // d2 >= d1: a pixel further than reach/2 from its nearest label keeps the class weight whatever d2
        // is, and so does the foreground; a warp with no other pixel skips the second sweep (75 % of the work)
        bool need = false, uni = true;
        unsigned lsel = 0;                                               // nearest label of this thread's needed rows
#pragma unroll
        for (int j = 0; j < IP_JR; ++j) {
            l1[j] = k1[j] & IP_LMASK;
            const unsigned c1 = k1[j] >> IP_LBITS;
            const bool nj = c1 != 0 && 4u * c1 <= reach2;
            if (nj) {
                if (!need) lsel = l1[j];
                uni &= l1[j] == lsel;
                need = true;
            }
        }
        if (!__any_sync(0xffffffffu, need)) {
            // no pixel of this warp can be affected by a second label
        } else if (__all_sync(0xffffffffu, uni)) {
            // Fast path (most warps: a thread's four rows share their nearest label): the row's candidate is chosen once
            // and feeds four add-mins -- 1.5 instructions per (pixel, dy) instead of 3.  Rows that are not `need`ed may
            // see the wrong exclusion label: harmless, every labelled pixel is further than reach/2 from them, so
            // d1 + d2 > reach whatever candidate wins.
#pragma unroll
            for (int r = 0; r < IP_JR + 2 * RMAX; ++r) {
                const unsigned ga = pa[r * IP_TW], gb = pb[r * IP_TW];
                const unsigned g = (ga & IP_LMASK) != lsel ? ga : gb;
#pragma unroll
                for (int j = 0; j < IP_JR; ++j) {
                    const int dy = r - RMAX - j;
                    if (dy >= -RMAX && dy <= RMAX) k2[j] = __viaddmin_u32(g, (unsigned)(dy * dy) << IP_LBITS, k2[j]);
                }
                if ((r & 7) == 7) asm volatile("" ::: "memory");
            }
        } else {
#pragma unroll
            for (int r = 0; r < IP_JR + 2 * RMAX; ++r) {
                const unsigned ga = pa[r * IP_TW], gb = pb[r * IP_TW], lr = ga & IP_LMASK;
#pragma unroll
                for (int j = 0; j < IP_JR; ++j) {
                    const int dy = r - RMAX - j;
                    if (dy >= -RMAX && dy <= RMAX)
                        k2[j] = __viaddmin_u32(lr != l1[j] ? ga : gb, (unsigned)(dy * dy) << IP_LBITS, k2[j]);
                }
                if ((r & 3) == 3) asm volatile("" ::: "memory");
            }
        }
        if (x >= wid) continue;
        OutT *o = out + fo + (long long)(y0 + ly) * wid + x;
#pragma unroll
        for (int j = 0; j < IP_JR; ++j, o += wid) {
            if (y0 + ly + j >= hgt) break;
            const unsigned c1 = k1[j] >> IP_LBITS, c2 = k2[j] >> IP_LBITS;
            double w = wc0;
            if (c1 == 0) w = wc1;                                            // foreground: row distance 0
            else if (c2 < IP_INFC && c1 + c2 <= reach2) {                    // (d1 + d2)^2 >= d1^2 + d2^2: cheap reject
                const double sum = sqrt((double)c1) + sqrt((double)c2);
                if (sum <= reach) w = wc0 + w0 * exp(-(sum * sum) / denom);   // beyond: cannot change the result
            }
            *o = (OutT)w;
        }
    }
}

// W3 row pass on the row's RUN LIST (default when the row fits: wid <= 3072).  Same results as inst_rows_bits below,
// which answers every pixel with its own bit-scan walk (~100 instructions and up to 8 dependent label loads per
// pixel); here the row is turned into runs once (start, end, label, nearest run with ANOTHER label on either side),
// and a pixel only ranks itself among the run starts and looks at <= 4 candidates: own run, or the two runs framing
// its gap, plus their different-label neighbours -- exactly the candidates the walk would have found, in the same
// order (own, left near / far, right near / far).
// grid (hgt, n), block 256, dyn smem: 4 * (nw + 1) words + wid * 12 bytes.
__global__ void __launch_bounds__(256)
inst_rows_runs(const int *__restrict__ labels, int *__restrict__ la, int *__restrict__ lb,
               unsigned char *__restrict__ da, unsigned char *__restrict__ db, int *__restrict__ big, int hgt, int wid,
               int R)
{
    extern __shared__ unsigned bits[];
    const long long ro = ((long long)blockIdx.y * hgt + blockIdx.x) * wid;
    const int *lr = labels + ro;
    const int nw = (wid + 31) >> 5;
    unsigned *sbits = bits, *ebits = bits + (nw + 1);
    int *spre = reinterpret_cast<int *>(bits + 2 * (nw + 1)), *epre = spre + (nw + 1);   // exclusive bit counts per word
    int *rl = epre + (nw + 1);                                                     // run label
    unsigned short *rs = reinterpret_cast<unsigned short *>(rl + wid), *re = rs + wid, *pd = re + wid, *nd = pd + wid;
    constexpr unsigned short NONE = 0xffffu;
    const int lane = threadIdx.x & 31;
    for (int x = threadIdx.x; x < nw * 32; x += blockDim.x) {
        int l = 0, lp = 0, ln = 0;
        if (x < wid) {
            l = lr[x];
            lp = x > 0 ? lr[x - 1] : 0;
            ln = x + 1 < wid ? lr[x + 1] : 0;
        }
        const bool fg = l > 0;
        if (l > (int)IP_LMASK) big[blockIdx.y] = 1;          // label too wide for the packed keys (benign race)
        const unsigned sb = __ballot_sync(0xffffffffu, fg && lp != l);
        const unsigned eb = __ballot_sync(0xffffffffu, fg && ln != l);
        if (lane == 0) { sbits[x >> 5] = sb; ebits[x >> 5] = eb; }
    }
    __syncthreads();
    if (threadIdx.x < 32) {                                   // exclusive prefix of the per-word bit counts
        int cs = 0, ce = 0;
        for (int w0 = 0; w0 < nw; w0 += 32) {
            const int w = w0 + lane;
            const int ps = w < nw ? __popc(sbits[w]) : 0, pe = w < nw ? __popc(ebits[w]) : 0;
            int is = ps, ie = pe;
            for (int o = 1; o < 32; o <<= 1) {
                const int ts = __shfl_up_sync(0xffffffffu, is, o), te = __shfl_up_sync(0xffffffffu, ie, o);
                if (lane >= o) { is += ts; ie += te; }
            }
            if (w < nw) { spre[w] = cs + is - ps; epre[w] = ce + ie - pe; }
            cs += __shfl_sync(0xffffffffu, is, 31);
            ce += __shfl_sync(0xffffffffu, ie, 31);
        }
        if (lane == 0) { spre[nw] = cs; epre[nw] = ce; }
    }
    __syncthreads();
    const int nruns = spre[nw];
    for (int w = threadIdx.x; w < nw; w += blockDim.x) {      // the k-th start bit and the k-th end bit frame run k
        unsigned a = sbits[w], e = ebits[w];
        int ks = spre[w], ke = epre[w];
        while (a) { const int b = __ffs(a) - 1; a &= a - 1; const int x = (w << 5) + b; rs[ks] = (unsigned short)x; rl[ks] = lr[x]; ++ks; }
        while (e) { const int b = __ffs(e) - 1; e &= e - 1; re[ke++] = (unsigned short)((w << 5) + b); }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nruns; k += blockDim.x) {   // nearest run with another label, within reach
        const int l = rl[k];
        int j = k - 1;
        while (j >= 0 && rl[j] == l && (int)rs[k] - (int)re[j] <= R) --j;
        pd[k] = (j >= 0 && rl[j] != l) ? (unsigned short)j : NONE;
        j = k + 1;
        while (j < nruns && rl[j] == l && (int)rs[j] - (int)re[k] <= R) ++j;
        nd[k] = (j < nruns && rl[j] != l) ? (unsigned short)j : NONE;
    }
    __syncthreads();
    for (int x = threadIdx.x; x < wid; x += blockDim.x) {
        const int w = x >> 5;
        const int k = spre[w] + __popc(sbits[w] & (0xffffffffu >> (31 - (x & 31)))) - 1;   // last run starting at or before x
        Best2 b = {0, 0, INF32, INF32};
        const bool inside = k >= 0 && x <= (int)re[k];
        int lft, rgt;                                         // nearest candidate runs on either side
        if (inside) {
            best2_insert(b, rl[k], 0u);
            lft = pd[k] == NONE ? -1 : (int)pd[k];
            rgt = nd[k] == NONE ? -1 : (int)nd[k];
        } else {
            lft = k;
            rgt = k + 1 < nruns ? k + 1 : -1;
        }
        if (lft >= 0 && x - (int)re[lft] <= R) {
            best2_insert(b, rl[lft], (unsigned)(x - (int)re[lft]));
            if (!inside && pd[lft] != NONE && x - (int)re[pd[lft]] <= R)
                best2_insert(b, rl[pd[lft]], (unsigned)(x - (int)re[pd[lft]]));
        }
        if (rgt >= 0 && (int)rs[rgt] - x <= R) {
            best2_insert(b, rl[rgt], (unsigned)((int)rs[rgt] - x));
            if (!inside && nd[rgt] != NONE && (int)rs[nd[rgt]] - x <= R)
                best2_insert(b, rl[nd[rgt]], (unsigned)((int)rs[nd[rgt]] - x));
        }
        la[ro + x] = b.la;
        lb[ro + x] = b.lb;
        da[ro + x] = b.la ? (unsigned char)min(b.da, 254u) : INF8;
        db[ro + x] = b.lb ? (unsigned char)min(b.db, 254u) : INF8;
    }
}

// W3 row pass.  grid (hgt, n), block 256, dyn smem: 2 * ceil(wid/32) words (run starts, run ends)
__global__ void inst_rows_bits(const int *__restrict__ labels, int *__restrict__ la,
                               int *__restrict__ lb, unsigned char *__restrict__ da,
                               unsigned char *__restrict__ db, int *__restrict__ big, int hgt, int wid, int R)
{
    extern __shared__ unsigned bits[];
    const long long ro = ((long long)blockIdx.y * hgt + blockIdx.x) * wid;
    const int *lr = labels + ro;
    const int nw = (wid + 31) >> 5;
    unsigned *sbits = bits, *ebits = bits + nw;
    for (int x = threadIdx.x; x < nw * 32; x += blockDim.x) {
        int l = 0, lp = 0, ln = 0;
        if (x < wid) {
            l = lr[x];
            lp = x > 0 ? lr[x - 1] : 0;
            ln = x + 1 < wid ? lr[x + 1] : 0;
        }
        const bool fg = l > 0;
        if (l > (int)IP_LMASK) big[blockIdx.y] = 1;          // label too wide for the packed keys (benign race)
        const unsigned s = __ballot_sync(0xffffffffu, fg && lp != l);
        const unsigned e = __ballot_sync(0xffffffffu, fg && ln != l);
        if ((threadIdx.x & 31) == 0) { sbits[x >> 5] = s; ebits[x >> 5] = e; }
    }
    __syncthreads();
    for (int x = threadIdx.x; x < wid; x += blockDim.x) {
        const int own = lr[x] > 0 ? lr[x] : 0;
        Best2 b = {0, 0, INF32, INF32};
        if (own) best2_insert(b, own, 0u);
        const int lo = max(x - R, 0), hi = min(x + R, wid - 1);
        // ---- to the left: walk run ends, skipping runs of a label already taken on this side
        {
            int pos = own ? prev_set(sbits, x, 0) - 1 : x;       // left of the own run
            int first = own;
            for (int it = 0; it < 48 && pos >= lo; ++it) {
                const int e = prev_set(ebits, pos, lo);
                if (e < 0) break;
                const int l = lr[e];
                if (l != first) {
                    best2_insert(b, l, (unsigned)(x - e));
                    if (first) break;                             // two distinct labels on this side
                    first = l;
                }
                const int st = prev_set(sbits, e, 0);
                if (st < 0) break;
                pos = st - 1;                                     // jump over that run
            }
        }
        // ---- to the right
        {
            int pos = x;
            if (own) { const int e = next_set(ebits, x, wid - 1); pos = (e < 0 ? wid : e + 1); }
            int first = own;
            for (int it = 0; it < 48 && pos <= hi; ++it) {
                const int s = next_set(sbits, pos, hi);
                if (s < 0) break;
                const int l = lr[s];
                if (l != first) {
                    best2_insert(b, l, (unsigned)(s - x));
                    if (first) break;
                    first = l;
                }
                const int e = next_set(ebits, s, wid - 1);
                if (e < 0) break;
                pos = e + 1;
            }
        }
        la[ro + x] = b.la;
        lb[ro + x] = b.lb;
        da[ro + x] = b.la ? (unsigned char)min(b.da, 254u) : INF8;
        db[ro + x] = b.lb ? (unsigned char)min(b.db, 254u) : INF8;
    }
}

// grid (ceil(W/32), ceil(H/64), n), block 256, dyn smem rows8 * 32 * 10 + (rows8/8) * 32 bytes.
// Same block-by-block scan as edt_cols_tile: the minimum nearest-label distance of every 8-row block
// bounds BOTH candidates of its rows from below, so a block is opened only if
// dy_min^2 + block_min^2 < the current second-best squared distance.
template <typename OutT>
__global__ void __launch_bounds__(256)
inst_cols_tile(const int *__restrict__ la, const int *__restrict__ lb,
               const unsigned char *__restrict__ da, const unsigned char *__restrict__ db,
               OutT *__restrict__ out, int hgt, int wid, int R, double w0, double denom, double wc0,
               double wc1, const int *__restrict__ only_if)
{
    if (only_if && !only_if[blockIdx.z]) return;       // this frame went through inst_cols_dpx
    extern __shared__ __align__(16) unsigned char ws_[];
    const int rows = CT_H + 2 * R, nblk = (rows + 7) >> 3, rows8 = nblk * 8;
    int *sla = reinterpret_cast<int *>(ws_);
    int *slb = sla + rows8 * CT_W;
    unsigned char *sda = reinterpret_cast<unsigned char *>(slb + rows8 * CT_W);
    unsigned char *sdb = sda + rows8 * CT_W;
    unsigned char *bm = sdb + rows8 * CT_W;              // [nblk][32] block minima of sda
    const int x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
    const long long fo = (long long)blockIdx.z * hgt * wid;
    const int col = threadIdx.x & 31, x = x0 + col, warp = threadIdx.x >> 5;
    for (int r = warp; r < rows8; r += 8) {
        const int yy = y0 - R + r;
        int l1 = 0, l2 = 0;
        unsigned char d1 = INF8, d2 = INF8;
        if (r < rows && yy >= 0 && yy < hgt && x < wid) {
            const long long j = fo + (long long)yy * wid + x;
            l1 = la[j]; l2 = lb[j]; d1 = da[j]; d2 = db[j];
        }
        sla[r * CT_W + col] = l1; slb[r * CT_W + col] = l2;
        sda[r * CT_W + col] = d1; sdb[r * CT_W + col] = d2;
    }
    __syncthreads();
    for (int b = warp; b < nblk; b += 8) {
        unsigned m = INF8;
#pragma unroll
        for (int k = 0; k < 8; ++k) m = min(m, (unsigned)sda[(b * 8 + k) * CT_W + col]);
        bm[b * CT_W + col] = (unsigned char)m;
    }
    __syncthreads();
    const double reach = (double)(R + 1);
    for (int ly = warp; ly < CT_H; ly += 8) {
        const int y = y0 + ly;
        if (x >= wid || y >= hgt) continue;
        const long long idx = fo + (long long)y * wid + x;
        const int p = ly + R;
        if (sda[p * CT_W + col] == 0) { out[idx] = (OutT)wc1; continue; }        // foreground: row distance 0
        Best2 b = {0, 0, INF32, INF32};
        const int bo = p >> 3, off = p & 7;
        auto scan_block = [&](int blk, int dy0, int step) {      // rows j = 0..7 at vertical offset dy0 + step*j
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int q = (blk * 8 + j) * CT_W + col;
                const unsigned d = sda[q];
                if (d == INF8) continue;                       // no label within R in this row
                const int dy = dy0 + step * j;
                const unsigned dy2 = (unsigned)(dy * dy), c1 = dy2 + d * d;
                if (c1 >= b.db) continue;                      // cannot change either of the two best
                best2_insert(b, sla[q], c1);
                const unsigned e = sdb[q];
                if (e == INF8) continue;
                const unsigned c2 = dy2 + e * e;
                if (c2 < b.db) best2_insert(b, slb[q], c2);
            }
        };
        if (bm[bo * CT_W + col] != INF8) scan_block(bo, -off, 1);
        for (int k = 1; k < nblk; ++k) {
            const int du = off + 8 * k - 7, dd = 8 * k - off;    // nearest row of block bo-k / bo+k
            const unsigned du2 = (unsigned)(du * du), dd2 = (unsigned)(dd * dd);
            if ((du2 >= b.db && dd2 >= b.db) || (du > R && dd > R)) break;       // rows beyond R cannot matter
            if (bo - k >= 0) {
                const unsigned m = bm[(bo - k) * CT_W + col];
                if (m != INF8 && du2 + m * m < b.db) scan_block(bo - k, du + 7, -1);
            }
            if (bo + k < nblk) {
                const unsigned m = bm[(bo + k) * CT_W + col];
                if (m != INF8 && dd2 + m * m < b.db) scan_block(bo + k, dd, 1);
            }
        }
        double w = wc0;
        if (b.lb != 0) {
            const double sum = sqrt((double)b.da) + sqrt((double)b.db);
            if (sum <= reach) w = wc0 + w0 * exp(-(sum * sum) / denom);   // beyond: cannot change the result
        }
        out[idx] = (OutT)w;
    }
}

// Radius beyond which |w0|*exp(-r^2/denom) is < 2^-bits * scale, i.e. cannot change
// the rounded result (scale = magnitude of the class term it is added to).
int cutoff_radius(double w0, double denom, double scale, int out_dtype, int maxdim)
{
    const double aw = std::fabs(w0);
    if (aw == 0.0) return 1;
    if (!(scale > 0.0) || !std::isfinite(aw) || !std::isfinite(denom)) return maxdim;
    const double bits = (out_dtype == SQ_F32) ? 27.0 : 56.0;
    const double t = std::log(aw / scale) + bits * 0.6931471805599453;
    if (t <= 0.0) return 1;
    const double r = std::ceil(std::sqrt(denom * t)) + 1.0;
    if (!(r < (double)maxdim)) return maxdim;
    return (int)(r < 1.0 ? 1.0 : r);
}

int check_common(sq_handle_t h, int n, int hgt, int wid, int out_dtype)
{
    SQ_REQUIRE(h, SQ_EINVAL, "weightmap: null handle");
    SQ_REQUIRE(n >= 1 && hgt >= 1 && wid >= 1, SQ_EINVAL, "weightmap: bad shape (%d,%d,%d)", n, hgt, wid);
    SQ_REQUIRE(wid <= 32768 && hgt <= 32768, SQ_EINVAL, "weightmap: image side > 32768");
    SQ_REQUIRE(n <= 65535, SQ_EINVAL, "weightmap: more than 65535 frames per call");
    SQ_REQUIRE(out_dtype == SQ_F32 || out_dtype == SQ_F64, SQ_EINVAL, "weightmap: bad out_dtype");
    return SQ_OK;
}

}  // namespace

extern "C" int sq_weightmap_workspace_bytes(sq_handle_t h, int n, int hgt, int wid,
                                            int instance_mode, size_t *bytes)
{
    SQ_REQUIRE(bytes, SQ_EINVAL, "weightmap: null pointer");
    SQ_TRY(check_common(h, n, hgt, wid, SQ_F32));
    const size_t px = (size_t)n * hgt * wid;
    SqArena a(nullptr, 0);
    if (instance_mode) {
        a.take<int>(px); a.take<int>(px);
        a.take<unsigned short>(px); a.take<unsigned short>(px);
        a.take<int>(n);
    } else {
        a.take<unsigned short>(px);
        a.take<int>(n);
        a.take<double>(WR_MAX * WR_MAX + 2);
    }
    *bytes = a.off;
    return SQ_OK;
}

extern "C" int sq_weightmap_edt(sq_handle_t h, const uint8_t *mask, int n, int hgt, int wid,
                                double w0, double sigma, int out_dtype, void *out, int32_t *d2,
                                void *ws, size_t ws_bytes, void *stream_)
{
    SQ_TRY(check_common(h, n, hgt, wid, out_dtype));
    SQ_REQUIRE(mask && out && ws, SQ_EINVAL, "weightmap_edt: null pointer");
    const size_t px = (size_t)n * hgt * wid;
    SqArena a(ws, ws_bytes);
    unsigned short *g = a.take<unsigned short>(px);
    int *anyfg = a.take<int>(n);
    double *table = a.take<double>(WR_MAX * WR_MAX + 2);
    SQ_REQUIRE(a.ok(), SQ_ENOMEM, "weightmap_edt: workspace %zu < %zu bytes", ws_bytes, a.off);
    cudaStream_t st = (cudaStream_t)stream_;
    const double denom = 2.0 * sigma * sigma + 1e-99;          // pipeline.py:478
    const double w0e = (double)(float)w0;   // w0*(1.-image) is float32 arithmetic (image is float32)
    const int maxdim = hgt > wid ? hgt : wid;
    const int rmax = d2 ? maxdim : cutoff_radius(w0e, denom, 1.0, out_dtype, maxdim);

    SQ_CUDA(cudaMemsetAsync(anyfg, 0, n * sizeof(int), st));
    if (!d2 && rmax <= WR_MAX) {
        // bounded path (the common case: R = 32 for w0 = 10, sigma = 5)
        unsigned char *g8 = reinterpret_cast<unsigned char *>(g);
        if ((wid & 15) == 0 && (((uintptr_t)mask | (uintptr_t)g8) & 15) == 0) {
            const int threads = std::min(256, 32 * ((wid / 16 + 31) / 32));
            edt_rows_bits16<<<dim3(hgt, n), threads, (size_t)((wid + 31) / 32 + 4) * 4, st>>>(mask, g8, anyfg, hgt, wid, rmax);
        } else
            edt_rows_bits<<<dim3(hgt, n), 256, (size_t)((wid + 31) / 32 + 4) * 4, st>>>(mask, g8, anyfg, hgt, wid, rmax);
        w1_table_kernel<<<sq_div_up(rmax * rmax + 1, 256), 256, 0, st>>>(table, rmax * rmax + 1, w0e, denom);
        const dim3 tgrid(sq_div_up(wid, DP_TW), sq_div_up(hgt, DP_TH), n);
#define SQ_W1_COLS(T, RM) edt_cols_dpx<T, RM><<<tgrid, 256, 0, st>>>(g8, anyfg, table, (T *)out, hgt, wid, rmax, w0e, denom)
        if (out_dtype == SQ_F32) {
            if (rmax <= 36) SQ_W1_COLS(float, 36); else if (rmax <= 48) SQ_W1_COLS(float, 48); else SQ_W1_COLS(float, 64);
        } else {
            if (rmax <= 36) SQ_W1_COLS(double, 36); else if (rmax <= 48) SQ_W1_COLS(double, 48); else SQ_W1_COLS(double, 64);
        }
#undef SQ_W1_COLS
        SQ_CHECK_LAUNCH();
        return SQ_OK;
    }
    if ((size_t)wid * sizeof(unsigned short) > 48 * 1024)
        SQ_CUDA(cudaFuncSetAttribute(edt_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    edt_rows<<<dim3(hgt, n), ROW_THREADS, (size_t)wid * sizeof(unsigned short), st>>>(
        mask, g, anyfg, hgt, wid);
    const dim3 grid(sq_div_up(wid, 32), sq_div_up(hgt, 8), n), blk(32, 8);
    if (out_dtype == SQ_F32)
        edt_cols_weight<float><<<grid, blk, 0, st>>>(mask, g, anyfg, (float *)out, d2, hgt, wid,
                                                    rmax, w0e, denom);
    else
        edt_cols_weight<double><<<grid, blk, 0, st>>>(mask, g, anyfg, (double *)out, d2, hgt,
                                                     wid, rmax, w0e, denom);
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

extern "C" int sq_weightmap_unet(sq_handle_t h, const int32_t *labels, int n, int hgt, int wid,
                                 double w0, double sigma, const double *wc, int out_dtype,
                                 void *out, void *ws, size_t ws_bytes, void *stream_)
{
    SQ_TRY(check_common(h, n, hgt, wid, out_dtype));
    SQ_REQUIRE(labels && out && ws, SQ_EINVAL, "weightmap_unet: null pointer");
    const size_t px = (size_t)n * hgt * wid;
    SqArena a(ws, ws_bytes);
    int *la = a.take<int>(px), *lb = a.take<int>(px);
    unsigned short *da = a.take<unsigned short>(px), *db = a.take<unsigned short>(px);
    int *big = a.take<int>(n);
    SQ_REQUIRE(a.ok(), SQ_ENOMEM, "weightmap_unet: workspace %zu < %zu bytes", ws_bytes, a.off);
    cudaStream_t st = (cudaStream_t)stream_;
    const double denom = 2.0 * sigma * sigma + 1e-99;
    const double wc0 = wc ? wc[0] : 1.0, wc1 = wc ? wc[1] : 2.0;
    const int maxdim = hgt > wid ? hgt : wid;
    // d1 + d2 >= d2 >= dy: once dy passes the cut-off the gap term cannot matter
    const int rmax = cutoff_radius(w0, denom, std::fabs(wc0), out_dtype, maxdim);

    if (rmax <= WR_MAX) {
        unsigned char *da8 = reinterpret_cast<unsigned char *>(da), *db8 = reinterpret_cast<unsigned char *>(db);
        SQ_CUDA(cudaMemsetAsync(big, 0, n * sizeof(int), st));
        // row pass on the run list when its shared memory fits the default 48 KB (wid <= 3072), else the bit-scan walk
        const size_t runs_smem = (size_t)((wid + 31) / 32 + 1) * 16 + (size_t)wid * 12;
        static const bool walk_only = getenv("SQ_W3_ROWWALK") != nullptr;     // A/B switch
        if (runs_smem <= 48 * 1024 && wid < 65535 && !walk_only)
            inst_rows_runs<<<dim3(hgt, n), 256, runs_smem, st>>>(labels, la, lb, da8, db8, big, hgt, wid, rmax);
        else
            inst_rows_bits<<<dim3(hgt, n), 256, (size_t)((wid + 31) / 32) * 8, st>>>(labels, la, lb, da8, db8, big, hgt, wid, rmax);
        {
            // DPX column pass (labels < 2^18); frames flagged by the row pass fall through to the scan kernel
            const dim3 dgrid(sq_div_up(wid, IP_TW), sq_div_up(hgt, IP_TH), n);
#define SQ_W3_COLS(T, RM)                                                                                          \
    do {                                                                                                           \
        auto k = inst_cols_dpx<T, RM>;                                                                             \
        const size_t dsm = (size_t)2 * (IP_TH + 2 * RM) * IP_TW * sizeof(unsigned);                                \
        SQ_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));                  \
        k<<<dgrid, 256, dsm, st>>>(la, lb, da8, db8, big, (T *)out, hgt, wid, rmax, w0, denom, wc0, wc1);           \
    } while (0)
            if (out_dtype == SQ_F32) {
                if (rmax <= 36) SQ_W3_COLS(float, 36); else if (rmax <= 48) SQ_W3_COLS(float, 48); else SQ_W3_COLS(float, 64);
            } else {
                if (rmax <= 36) SQ_W3_COLS(double, 36); else if (rmax <= 48) SQ_W3_COLS(double, 48); else SQ_W3_COLS(double, 64);
            }
#undef SQ_W3_COLS
            SQ_CHECK_LAUNCH();
        }
        const dim3 tgrid(sq_div_up(wid, CT_W), sq_div_up(hgt, CT_H), n);
        const int rows8 = (CT_H + 2 * rmax + 7) / 8 * 8;
        const size_t sm = (size_t)rows8 * CT_W * 10 + (size_t)(rows8 / 8) * CT_W;
        if (out_dtype == SQ_F32) {
            auto k = inst_cols_tile<float>;
            if (sm > 48 * 1024) SQ_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            k<<<tgrid, 256, sm, st>>>(la, lb, da8, db8, (float *)out, hgt, wid, rmax, w0, denom, wc0, wc1, big);
        } else {
            auto k = inst_cols_tile<double>;
            if (sm > 48 * 1024) SQ_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            k<<<tgrid, 256, sm, st>>>(la, lb, da8, db8, (double *)out, hgt, wid, rmax, w0, denom, wc0, wc1, big);
        }
        SQ_CHECK_LAUNCH();
        return SQ_OK;
    }
    inst_rows<<<dim3(hgt, n), ROW_THREADS, 0, st>>>(labels, la, lb, da, db, hgt, wid);
    const dim3 grid(sq_div_up(wid, 32), sq_div_up(hgt, 8), n), blk(32, 8);
    if (out_dtype == SQ_F32)
        inst_cols_weight<float><<<grid, blk, 0, st>>>(labels, la, lb, da, db, (float *)out, hgt,
                                                     wid, rmax, w0, denom, wc0, wc1);
    else
        inst_cols_weight<double><<<grid, blk, 0, st>>>(labels, la, lb, da, db, (double *)out, hgt,
                                                      wid, rmax, w0, denom, wc0, wc1);
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

extern "C" int sq_weightmap_edt_host(sq_handle_t h, const uint8_t *mask_host, int n, int hgt,
                                     int wid, double w0, double sigma, int out_dtype,
                                     void *out_host, int32_t *d2_host)
{
    SQ_TRY(check_common(h, n, hgt, wid, out_dtype));
    SQ_REQUIRE(mask_host && out_host, SQ_EINVAL, "weightmap_edt_host: null pointer");
    SqHostCall call(h);
    SQ_CUDA(cudaSetDevice(h->device));
    const size_t px = (size_t)n * hgt * wid;
    const size_t esz = out_dtype == SQ_F32 ? 4 : 8;
    size_t ws_bytes = 0;
    SQ_TRY(sq_weightmap_workspace_bytes(h, n, hgt, wid, 0, &ws_bytes));
    SQ_TRY(sq_reserve_device(h, sq_align_up(px) + sq_align_up(px * esz) + sq_align_up(px * 4) +
                                    ws_bytes + 1024));
    SqArena a(h->dev_arena, h->dev_arena_bytes);
    uint8_t *mask = a.take<uint8_t>(px);
    char *out = a.take<char>(px * esz);
    int32_t *d2 = d2_host ? a.take<int32_t>(px) : nullptr;
    void *ws = a.take<char>(ws_bytes);
    cudaStream_t st = h->stream;
    SQ_CUDA(cudaMemcpyAsync(mask, mask_host, px, cudaMemcpyHostToDevice, st));
    SQ_TRY(sq_weightmap_edt(h, mask, n, hgt, wid, w0, sigma, out_dtype, out, d2, ws, ws_bytes, st));
    SQ_CUDA(cudaMemcpyAsync(out_host, out, px * esz, cudaMemcpyDeviceToHost, st));
    if (d2_host) SQ_CUDA(cudaMemcpyAsync(d2_host, d2, px * 4, cudaMemcpyDeviceToHost, st));
    SQ_CUDA(cudaStreamSynchronize(st));
    return SQ_OK;
}

extern "C" int sq_weightmap_unet_host(sq_handle_t h, const int32_t *labels_host, int n, int hgt,
                                      int wid, double w0, double sigma, const double *wc,
                                      int out_dtype, void *out_host)
{
    SQ_TRY(check_common(h, n, hgt, wid, out_dtype));
    SQ_REQUIRE(labels_host && out_host, SQ_EINVAL, "weightmap_unet_host: null pointer");
    SqHostCall call(h);
    SQ_CUDA(cudaSetDevice(h->device));
    const size_t px = (size_t)n * hgt * wid;
    const size_t esz = out_dtype == SQ_F32 ? 4 : 8;
    size_t ws_bytes = 0;
    SQ_TRY(sq_weightmap_workspace_bytes(h, n, hgt, wid, 1, &ws_bytes));
    SQ_TRY(sq_reserve_device(h, sq_align_up(px * 4) + sq_align_up(px * esz) + ws_bytes + 1024));
    SqArena a(h->dev_arena, h->dev_arena_bytes);
    int32_t *labels = a.take<int32_t>(px);
    char *out = a.take<char>(px * esz);
    void *ws = a.take<char>(ws_bytes);
    cudaStream_t st = h->stream;
    SQ_CUDA(cudaMemcpyAsync(labels, labels_host, px * 4, cudaMemcpyHostToDevice, st));
    SQ_TRY(sq_weightmap_unet(h, labels, n, hgt, wid, w0, sigma, wc, out_dtype, out, ws, ws_bytes, st));
    SQ_CUDA(cudaMemcpyAsync(out_host, out, px * esz, cudaMemcpyDeviceToHost, st));
    SQ_CUDA(cudaStreamSynchronize(st));
    return SQ_OK;
}
