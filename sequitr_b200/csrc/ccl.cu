// Label-and-localise on the GPU: run-based union-find connected-component labelling with
// centroid reduction.  Replaces the per-frame / per-class SciPy loop of
// utils.CentroidWriter.write (reference utils.py:531-566: ndimage.label :547,
// center_of_mass :550).  Integer work, bit-exact against the oracle.
//
// One pass handles ALL classes: two voxels are connected when they share a face and carry
// the same non-zero value, so label(out == c) for every c falls out of a single union-find.
// The mask is read as horizontal RUNS of equal non-zero value (a segmentation mask of a few
// hundred cells has ~10^4 runs against 4*10^6 pixels): only the run-extraction kernel
// touches every voxel (1 B/voxel, once); everything else -- unions
// between vertically (and, in 3-D, depth-) adjacent runs, path compression, ordering,
// 64-bit centroid sums -- is O(runs).  Run slots are allocated in raster order by a prefix
// sum, so the root of a component (minimum slot, atomicMin union) is the run holding the
// component's first voxel in raster order = SciPy's numbering order.
//
// Kernels (HBM-bound; algorithmic traffic 1 B/voxel in, rows out):
//   run_scan_emit runs per 512-voxel row segment (16 voxels per lane, byte-wise SIMD compares + shuffles), their
//                 raster-order slots by a single-pass decoupled look-back scan, run start / end lists, parent[i] = i
//   run_merge     unions with overlapping runs of the previous row / previous plane
//   run_compress  full path compression + root count per 2048-run chunk (only chunks that hold runs)
//   root_emit     raster-ordered root list (chunk offsets summed in the block: no scan launch)
//   ccl_order     stable counting sort of the roots by class -> table row index
//   run_accum     64-bit sums (count, sum z, sum y, sum x) per row from run endpoints
//   ccl_finalize  fp64 divide exactly as center_of_mass does, float32 rows
#include "sq_common.cuh"
#include <algorithm>

namespace {

constexpr int PPL = 16;              // voxels per lane (one 16-byte load)
constexpr int SEG = 32 * PPL;        // voxels per warp segment
constexpr int CHUNK = 2048;          // runs per compress/emit block
constexpr int CHUNK_THREADS = 256;

struct Dims {
    int n, D, H, W;
    int rows;                         // D*H lines per frame
    int vox;                          // rows*W  (< 2^31)
    int nseg;                         // segments per line
    int nsegs;                        // rows*nseg segments per frame
    int maxruns;                      // run capacity per frame
    int nchunks;                      // ceil(maxruns / CHUNK)
    int vec;                          // 1: 16-byte mask loads are aligned
};

__device__ __forceinline__ int find_root(const int *P, int i)
{
    int p = __ldcg(P + i);
    while (p != i) {
        i = p;
        p = __ldcg(P + i);
    }
    return i;
}

// lock-free union by minimum slot (Playne/Komura style)
__device__ __forceinline__ void unite(int *P, int a, int b)
{
    bool done = false;
    while (!done) {
        a = find_root(P, a);
        b = find_root(P, b);
        if (a < b) {
            int old = atomicMin(P + b, a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            int old = atomicMin(P + a, b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    }
}

// start / end bits of the PPL = 16 voxels this lane owns in its segment; returns popc(start) | popc(end)<<16.
// Byte-wise SIMD: a voxel starts a run when it is non-zero and differs from its left neighbour, ends one
// when it differs from its right neighbour (four voxels per compare).
__device__ __forceinline__ unsigned segment_bits(const uint8_t *__restrict__ line, int x, int W, int vec,
                                                 int lane, unsigned &sb, unsigned &eb)
{
    unsigned w[4] = {0u, 0u, 0u, 0u};
    if (vec && x + PPL - 1 < W) {
        const uint4 v = *reinterpret_cast<const uint4 *>(line + x);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < PPL; ++k)
            if (x + k < W) w[k >> 2] |= (unsigned)line[x + k] << (8 * (k & 3));
    }
    unsigned prev = __shfl_up_sync(0xffffffffu, w[3] >> 24, 1);
    unsigned next = __shfl_down_sync(0xffffffffu, w[0] & 0xffu, 1);
    if (lane == 0) prev = 0;                       // segment borders always cut a run
    if (lane == 31) next = 0;
    sb = 0;
    eb = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const unsigned lw = (w[i] << 8) | (i ? w[i - 1] >> 24 : prev);
        const unsigned rw = (w[i] >> 8) | ((i < 3 ? w[i + 1] & 0xffu : next) << 24);
        const unsigned nz = __vcmpne4(w[i], 0u);
        const unsigned s4 = nz & __vcmpne4(w[i], lw) & 0x01010101u, e4 = nz & __vcmpne4(w[i], rw) & 0x01010101u;
        sb |= ((s4 * 0x10204080u) >> 28) << (4 * i);           // bytes 0..3 -> bits 0..3
        eb |= ((e4 * 0x10204080u) >> 28) << (4 * i);
    }
    return (unsigned)__popc(sb) | ((unsigned)__popc(eb) << 16);
}

__device__ __forceinline__ unsigned warp_inclusive_scan(unsigned v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Single-pass front end: runs per segment, their raster-order slots (an exclusive prefix sum over the
// frame's segments) and the run lists, in ONE kernel that reads the mask once.  A block owns SPB = 128
// consecutive segments (warp w: segments 16w .. 16w+15; 64 blocks per 2048^2 frame, so a look-back is two rounds); the prefix across blocks is a decoupled look-back
// (Merrill & Garland): every block publishes its aggregate, then its inclusive prefix, in one 64-bit status
// word (flag << 32 | value) and warp 0 walks back over its predecessors' words, 32 at a time.  Blocks take
// their index from a per-frame ticket counter, so a block only ever waits for blocks that are already running.
// grid (ceil(nsegs / SPB), n), block 256.  status / ticket are zeroed by the caller.
constexpr int SPB = 128, SPW = SPB / 8, SPL = SPB / 32;     // segments per block / per warp / per lane of the block scan
constexpr unsigned long long ST_AGG = 1ull << 32, ST_PREFIX = 2ull << 32;

__global__ void __launch_bounds__(256)
run_scan_emit(const uint8_t *__restrict__ mask, unsigned long long *__restrict__ status, int *__restrict__ ticket,
              int *__restrict__ seg_off, int *__restrict__ totals, int *__restrict__ run_start,
              int *__restrict__ run_end, int *__restrict__ parent, Dims dm, int nblk)
{
    __shared__ int s_bid, s_excl, s_tot[SPB];
    const int f = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_bid = atomicAdd(ticket + f, 1);
    __syncthreads();
    const int bid = s_bid;
    const uint8_t *fm = mask + (long long)f * dm.vox;
    unsigned sb[SPW], eb[SPW], pre[SPW];
#pragma unroll
    for (int k = 0; k < SPW; ++k) {
        const int s = bid * SPB + wid * SPW + k;
        sb[k] = eb[k] = pre[k] = 0;
        unsigned tot = 0;
        if (s < dm.nsegs) {
            const int line = s / dm.nseg, seg = s - line * dm.nseg;
            const unsigned c = segment_bits(fm + (long long)line * dm.W, seg * SEG + lane * PPL, dm.W, dm.vec, lane, sb[k], eb[k]);
            const unsigned incl = warp_inclusive_scan(c, lane);
            pre[k] = incl - c;                                  // exclusive (starts | ends << 16) within the segment
            tot = __shfl_sync(0xffffffffu, incl, 31) & 0xffffu;
        }
        if (lane == 0) s_tot[wid * SPW + k] = (int)tot;
    }
    __syncthreads();
    if (wid == 0) {
        int v[SPL], lsum = 0;
#pragma unroll
        for (int q = 0; q < SPL; ++q) { v[q] = s_tot[lane * SPL + q]; lsum += v[q]; }
        const int incl = (int)warp_inclusive_scan((unsigned)lsum, lane);
        const int agg = __shfl_sync(0xffffffffu, incl, 31);
        int run = incl - lsum;
#pragma unroll
        for (int q = 0; q < SPL; ++q) { s_tot[lane * SPL + q] = run; run += v[q]; }     // exclusive prefix within the block
        volatile unsigned long long *st = status + (long long)f * nblk;
        int excl = 0;
        if (bid > 0) {
            if (lane == 0) st[bid] = ST_AGG | (unsigned)agg;
            int j = bid - 1;
            while (true) {
                const int idx = j - lane;
                unsigned long long w = ST_PREFIX;               // before the first block: prefix 0
                if (idx >= 0) {
                    do { w = st[idx]; } while ((w >> 32) == 0);
                }
                const unsigned pm = __ballot_sync(0xffffffffu, (w >> 32) == 2);
                // sum the aggregates up to (and including) the nearest predecessor that already has its prefix
                const int first = pm ? __ffs(pm) - 1 : 31;
                int val = (lane <= first) ? (int)(unsigned)w : 0;
                for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
                excl += __shfl_sync(0xffffffffu, val, 0);
                if (pm) break;
                j -= 32;
            }
        }
        if (lane == 0) {
            __threadfence();
            st[bid] = ST_PREFIX | (unsigned)(excl + agg);
            s_excl = excl;
            if (bid == nblk - 1) {
                totals[f] = excl + agg;
                seg_off[(long long)f * (dm.nsegs + 1) + dm.nsegs] = excl + agg;
            }
        }
    }
    __syncthreads();
    const long long rb = (long long)f * dm.maxruns;
#pragma unroll
    for (int k = 0; k < SPW; ++k) {
        const int s = bid * SPB + wid * SPW + k;
        if (s >= dm.nsegs) break;
        const int base = s_excl + s_tot[wid * SPW + k];
        if (lane == 0) seg_off[(long long)f * (dm.nsegs + 1) + s] = base;
        unsigned a = sb[k], e = eb[k];
        if (!(a | e)) continue;
        const int line = s / dm.nseg, x = (s - line * dm.nseg) * SEG + lane * PPL;
        int is = base + (int)(pre[k] & 0xffffu), ie = base + (int)(pre[k] >> 16);
        while (a) {                                             // set bits in ascending order = raster order
            const int b = __ffs(a) - 1;
            a &= a - 1;
            run_start[rb + is] = line * dm.W + x + b;
            parent[rb + is] = is;
            ++is;
        }
        while (e) {
            const int b = __ffs(e) - 1;
            e &= e - 1;
            run_end[rb + ie++] = x + b;
        }
    }
}

// unions of run i with the overlapping same-class runs of line `pl` (frame-local pointers)
__device__ __forceinline__ void merge_with_line(const uint8_t *mk, const int *seg_off,
                                                const int *run_start, const int *run_end, int *parent,
                                                const Dims &dm, int i, int pl, int x0, int x1, uint8_t c)
{
    // runs never cross a segment border, so [x0, x1] lies in ONE segment and only the runs of that segment
    // of line `pl` can overlap it
    const int sg = pl * dm.nseg + x0 / SEG;
    const int jb = seg_off[sg], je = seg_off[sg + 1];
    for (int j = jb; j < je; ++j) {
        const int s = run_start[j];
        const int xs = s - pl * dm.W;
        if (xs > x1) break;                                    // runs of a line are sorted by x
        if (run_end[j] >= x0 && mk[s] == c) unite(parent, i, j);
    }
}

// grid (blocks, n): grid-stride over the runs of a frame
__global__ void run_merge(const uint8_t *__restrict__ mask, const int *__restrict__ seg_off,
                          const int *__restrict__ totals, const int *__restrict__ run_start,
                          const int *__restrict__ run_end, int *__restrict__ parent, Dims dm)
{
    const int f = blockIdx.y;
    const int R = totals[f];
    const uint8_t *mk = mask + (long long)f * dm.vox;
    const int *so = seg_off + (long long)f * (dm.nsegs + 1);
    const int *rs = run_start + (long long)f * dm.maxruns;
    const int *re = run_end + (long long)f * dm.maxruns;
    int *pa = parent + (long long)f * dm.maxruns;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R; i += gridDim.x * blockDim.x) {
        const int s = rs[i];
        const int line = s / dm.W, x0 = s - line * dm.W, x1 = re[i];
        const uint8_t c = mk[s];
        // a run cut by a segment border continues in the previous slot
        if (x0 > 0 && (x0 % SEG) == 0 && i > 0) {
            const int ps = rs[i - 1];
            if (ps / dm.W == line && re[i - 1] == x0 - 1 && mk[ps] == c) unite(pa, i, i - 1);
        }
        const int y = line % dm.H;
        if (y > 0) merge_with_line(mk, so, rs, re, pa, dm, i, line - 1, x0, x1, c);
        if (line >= dm.H) merge_with_line(mk, so, rs, re, pa, dm, i, line - dm.H, x0, x1, c);
    }
}

__device__ __forceinline__ int block_sum_256(int v, int *sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = 0;
    if (threadIdx.x < 32) {
        r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    return r;                                                  // valid in thread 0
}

// grid (blocks, n), block 256: full path compression and the root count of every 2048-run chunk that holds
// runs (chunks past the frame's last run are never touched: root_emit only reads the first ceil(R / CHUNK))
__global__ void run_compress(int *__restrict__ parent, const int *__restrict__ totals,
                             int *__restrict__ chunk_count, Dims dm)
{
    __shared__ int sh[8];
    const int f = blockIdx.y;
    const int R = totals[f];
    int *pa = parent + (long long)f * dm.maxruns;
    for (int chunk = blockIdx.x; chunk * CHUNK < R; chunk += gridDim.x) {
        const int base = chunk * CHUNK;
        int cnt = 0;
        // the thread's 8 runs walk to their roots in LOCK-STEP: 8 independent pointer chases in flight per
        // thread instead of one after the other (the walk is L2-latency bound: tall components leave chains
        // as deep as their height)
        constexpr int NW = CHUNK / CHUNK_THREADS;
        int x[NW], first[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const int i = base + k * CHUNK_THREADS + threadIdx.x;
            first[k] = (i < R) ? __ldcg(pa + i) : -1;
            x[k] = first[k];
        }
        bool moving = true;
        while (moving) {
            moving = false;
            int p[NW];
#pragma unroll
            for (int k = 0; k < NW; ++k) p[k] = (x[k] >= 0) ? __ldcg(pa + x[k]) : -1;
#pragma unroll
            for (int k = 0; k < NW; ++k)
                if (p[k] != x[k]) { x[k] = p[k]; moving = true; }
        }
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const int i = base + k * CHUNK_THREADS + threadIdx.x;
            if (i < R) {
                if (x[k] != first[k]) pa[i] = x[k];
                cnt += (x[k] == i);
            }
        }
        __syncthreads();                                       // sh is reused across chunks
        const int tot = block_sum_256(cnt, sh);
        if (threadIdx.x == 0) chunk_count[(long long)f * (dm.nchunks + 1) + chunk] = tot;
    }
}

// grid (blocks, n), block 256: raster-ordered root list; thread t owns 8 consecutive runs of a chunk.  The
// chunk's offset is the sum of the counts of the chunks before it (a handful for a real frame), added up by
// the block itself; the block of chunk 0 also publishes the frame's root count.
__global__ void root_emit(const int *__restrict__ parent, const int *__restrict__ totals,
                          const int *__restrict__ chunk_count, int *__restrict__ nroots,
                          int *__restrict__ rootlist, Dims dm, int max_rows)
{
    __shared__ int warp_sums[8], sh[8], s_off;
    const int f = blockIdx.y;
    const int R = totals[f];
    const int nact = (R + CHUNK - 1) / CHUNK;
    const int *cc = chunk_count + (long long)f * (dm.nchunks + 1);
    const int *pa = parent + (long long)f * dm.maxruns;
    if (blockIdx.x == 0 && nact == 0 && threadIdx.x == 0) nroots[f] = 0;
    for (int chunk = blockIdx.x; chunk < nact; chunk += gridDim.x) {
        int before = 0, all = 0;
        for (int i = threadIdx.x; i < nact; i += blockDim.x) {
            const int c = cc[i];
            all += c;
            if (i < chunk) before += c;
        }
        __syncthreads();
        const int off = block_sum_256(before, sh);
        if (threadIdx.x == 0) s_off = off;
        __syncthreads();
        if (chunk == 0) {
            const int tot = block_sum_256(all, sh);
            if (threadIdx.x == 0) nroots[f] = tot;
            __syncthreads();
        }
        if (cc[chunk] != 0) {
            const int base = chunk * CHUNK + threadIdx.x * 8;
            int flags = 0, cnt = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = base + k;
                if (i < R && pa[i] == i) { flags |= 1 << k; ++cnt; }
            }
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            int s2 = cnt;
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, s2, o);
                if (lane >= o) s2 += t;
            }
            if (lane == 31) warp_sums[wid] = s2;
            __syncthreads();
            int pre = 0;
            for (int w = 0; w < wid; ++w) pre += warp_sums[w];
            int pos = s_off + pre + s2 - cnt;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (flags & (1 << k)) {
                    if (pos < max_rows) rootlist[(long long)f * max_rows + pos] = base + k;
                    ++pos;
                }
        }
        __syncthreads();
    }
}

// grid (n), block 256: stable counting sort of the frame's root runs by class value.
// Writes parent[root] = -2 - row, sorted_root (root run slot of each row), class_base.
// The class histogram uses shared-memory atomics; the stable rank of root j is base[class] +
// #(i < j with the same class).  For up to ORDER_PAR roots every thread counts its predecessors
// in parallel (O(n^2/256), a few microseconds for the hundreds of cells of a real frame); beyond
// that thread c walks the raster-ordered list for class c.
constexpr int ORDER_PAR = 4096;

__global__ void ccl_order(const uint8_t *__restrict__ mask, const int *__restrict__ run_start,
                          int *__restrict__ parent, const int *__restrict__ rootlist,
                          const int *__restrict__ nroots, int *__restrict__ sorted_root,
                          int *__restrict__ class_base, Dims dm, int max_rows)
{
    __shared__ __align__(4) uint8_t cls[ORDER_PAR];
    __shared__ int cnt_sh[256];
    const int f = blockIdx.x;
    const int c = threadIdx.x;
    const int n = min(nroots[f], max_rows);
    const uint8_t *mk = mask + (long long)f * dm.vox;
    const int *rs = run_start + (long long)f * dm.maxruns;
    int *pa = parent + (long long)f * dm.maxruns;
    const int *rl = rootlist + (long long)f * max_rows;
    int *sr = sorted_root + (long long)f * max_rows;

    cnt_sh[c] = 0;
    __syncthreads();
    // class of every root: three dependent loads (root slot -> first voxel -> class), four roots per thread in flight
    for (int j0 = 0; j0 < n; j0 += 1024) {
        int r[4], sv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int j = j0 + q * 256 + c; r[q] = j < n ? rl[j] : -1; }
#pragma unroll
        for (int q = 0; q < 4; ++q) sv[q] = r[q] >= 0 ? rs[r[q]] : 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (r[q] < 0) continue;
            const int j = j0 + q * 256 + c;
            const uint8_t v = mk[sv[q]];
            if (j < ORDER_PAR) cls[j] = v;
            atomicAdd(&cnt_sh[v], 1);
        }
    }
    __syncthreads();
    const int mine = cnt_sh[c];
    __syncthreads();
    if (c == 0) {
        int run = 0;
        for (int k = 0; k < 256; ++k) { int t = cnt_sh[k]; cnt_sh[k] = run; run += t; }
    }
    __syncthreads();
    class_base[f * 256 + c] = cnt_sh[c];
    if (n <= ORDER_PAR) {
        for (int j = c; j < n; j += 256) {
            // stable rank = earlier roots of the same class, four class bytes per compare
            const uint8_t v = cls[j];
            const unsigned vv = (unsigned)v * 0x01010101u;
            const unsigned *cw = reinterpret_cast<const unsigned *>(cls);
            const int full = j >> 2, rem = j & 3;
            int rank = 0;
            for (int i = 0; i < full; ++i) rank += __popc(__vcmpeq4(cw[i], vv) & 0x01010101u);
            if (rem) rank += __popc(__vcmpeq4(cw[full], vv) & (0x01010101u & ((1u << (8 * rem)) - 1u)));
            const int row = cnt_sh[v] + rank;
            const int r = rl[j];
            pa[r] = -2 - row;
            sr[row] = r;
        }
    } else if (mine > 0) {
        int k = cnt_sh[c];
        for (int j = 0; j < n; ++j) {
            const int r = rl[j];
            if (mk[rs[r]] == c) { pa[r] = -2 - k; sr[k] = r; ++k; }
        }
    }
}

// grid (blocks, n): grid-stride over runs; closed-form sums from the run endpoints
__global__ void run_accum(const uint8_t *__restrict__ mask, const int *__restrict__ totals,
                          const int *__restrict__ run_start, const int *__restrict__ run_end,
                          const int *__restrict__ parent, const int *__restrict__ class_base,
                          unsigned long long *__restrict__ acc, int *__restrict__ labels_out, Dims dm,
                          int max_rows)
{
    const int f = blockIdx.y;
    const int R = totals[f];
    const int *rs = run_start + (long long)f * dm.maxruns;
    const int *re = run_end + (long long)f * dm.maxruns;
    const int *pa = parent + (long long)f * dm.maxruns;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R; i += gridDim.x * blockDim.x) {
        const int l = pa[i];
        int row = -1;
        if (l <= -2) row = -2 - l;
        else {
            const int rr = pa[l];
            if (rr <= -2) row = -2 - rr;
        }
        if (row < 0) continue;                                 // component beyond max_rows
        const int s = rs[i];
        const int line = s / dm.W, x0 = s - line * dm.W, x1 = re[i];
        const unsigned long long len = (unsigned long long)(x1 - x0 + 1);
        unsigned long long *a = acc + ((long long)f * max_rows + row) * 4;
        atomicAdd(a + 0, len);
        if (dm.D > 1) atomicAdd(a + 1, len * (unsigned long long)(line / dm.H));
        atomicAdd(a + 2, len * (unsigned long long)(line % dm.H));
        atomicAdd(a + 3, (unsigned long long)(x0 + x1) * len / 2);
        if (labels_out) {
            const int lab = row - class_base[f * 256 + mask[(long long)f * dm.vox + s]] + 1;
            int *lo = labels_out + (long long)f * dm.vox + s;
            for (int k = 0; k <= x1 - x0; ++k) lo[k] = lab;
        }
    }
}

// grid (ceil(max_rows/256), n)
__global__ void ccl_finalize(const uint8_t *__restrict__ mask, const int *__restrict__ run_start,
                             const int *__restrict__ sorted_root,
                             const unsigned long long *__restrict__ acc,
                             const int *__restrict__ nroots, float *__restrict__ table,
                             int *__restrict__ counts, Dims dm, int max_rows, int frame0)
{
    const int f = blockIdx.y;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int tot = nroots[f];
    if (row == 0) counts[f] = tot;
    if (row >= min(tot, max_rows)) return;
    const int r = sorted_root[(long long)f * max_rows + row];
    const unsigned long long c = mask[(long long)f * dm.vox + run_start[(long long)f * dm.maxruns + r]];
    const unsigned long long *a = acc + ((long long)f * max_rows + row) * 4;
    // center_of_mass: sum(c * index) / sum(c), both exact integers in fp64
    const double den = (double)(c * a[0]);
    const double cz = (double)(c * a[1]) / den;
    const double cy = (double)(c * a[2]) / den;
    const double cx = (double)(c * a[3]) / den;
    float *t = table + ((long long)f * max_rows + row) * 5;
    t[0] = (float)(frame0 + f);
    if (dm.D > 1) { t[1] = (float)cz; t[2] = (float)cy; t[3] = (float)cx; }
    else          { t[1] = (float)cy; t[2] = (float)cx; t[3] = 0.0f; }
    t[4] = (float)c;
}

struct Workspace {
    int *seg_off, *totals, *run_start, *run_end, *parent;
    int *chunk_count, *nroots, *rootlist, *sorted_root, *class_base;
    // zeroed before every call, contiguous: look-back status words, centroid accumulators, block tickets
    unsigned long long *status, *acc;
    int *ticket;
    size_t zero_bytes;
    int nblk;
    size_t bytes;
};

Workspace carve(void *p, size_t avail, const Dims &dm, int max_rows)
{
    SqArena a(p, avail);
    Workspace w;
    w.nblk = sq_div_up(dm.nsegs, SPB);
    w.seg_off = a.take<int>((size_t)dm.n * (dm.nsegs + 1));
    w.totals = a.take<int>(dm.n);
    w.run_start = a.take<int>((size_t)dm.n * dm.maxruns);
    w.run_end = a.take<int>((size_t)dm.n * dm.maxruns);
    w.parent = a.take<int>((size_t)dm.n * dm.maxruns);
    w.chunk_count = a.take<int>((size_t)dm.n * (dm.nchunks + 1));
    w.nroots = a.take<int>(dm.n);
    w.rootlist = a.take<int>((size_t)dm.n * max_rows);
    w.sorted_root = a.take<int>((size_t)dm.n * max_rows);
    w.class_base = a.take<int>((size_t)dm.n * 256);
    const size_t z0 = a.off;
    w.status = a.take<unsigned long long>((size_t)dm.n * w.nblk);
    w.acc = a.take<unsigned long long>((size_t)dm.n * max_rows * 4);
    w.ticket = a.take<int>(dm.n);
    w.zero_bytes = a.off - z0;
    w.bytes = a.off;
    return w;
}

int make_dims(int n, int d, int hgt, int wid, int max_rows, const void *mask, Dims *dm)
{
    SQ_REQUIRE(n >= 1 && d >= 1 && hgt >= 1 && wid >= 1, SQ_EINVAL,
               "label: bad shape (%d,%d,%d,%d)", n, d, hgt, wid);
    SQ_REQUIRE(max_rows >= 1, SQ_EINVAL, "label: max_rows must be >= 1");
    SQ_REQUIRE(n <= 65535, SQ_EINVAL, "label: more than 65535 frames per call");
    const long long vox = (long long)d * hgt * wid;
    SQ_REQUIRE(vox < (1ll << 31) - CHUNK, SQ_EINVAL, "label: frame too large (%lld voxels)", vox);
    dm->n = n; dm->D = d; dm->H = hgt; dm->W = wid;
    dm->rows = d * hgt;
    dm->vox = (int)vox;
    dm->nseg = sq_div_up(wid, SEG);
    dm->nsegs = dm->rows * dm->nseg;
    // worst case: alternating values, plus one cut per segment border
    const long long mr = (long long)dm->rows * ((wid + 1) / 2 + dm->nseg);
    SQ_REQUIRE(mr < (1ll << 31) - CHUNK, SQ_EINVAL, "label: frame too large");
    dm->maxruns = (int)mr;
    dm->nchunks = sq_div_up(mr, CHUNK);
    dm->vec = (wid % PPL == 0) && (((uintptr_t)mask & (PPL - 1)) == 0);     // 16-byte mask loads are aligned
    return SQ_OK;
}

}  // namespace

extern "C" int sq_label_workspace_bytes(sq_handle_t h, int n, int d, int hgt, int wid,
                                        int max_rows, size_t *bytes)
{
    SQ_REQUIRE(h && bytes, SQ_EINVAL, "label: null handle/pointer");
    Dims dm;
    SQ_TRY(make_dims(n, d, hgt, wid, max_rows, nullptr, &dm));
    *bytes = carve(nullptr, 0, dm, max_rows).bytes;
    return SQ_OK;
}

extern "C" int sq_label_centroids(sq_handle_t h, const uint8_t *mask, int n, int d, int hgt,
                                  int wid, int frame0, int32_t *labels, float *table,
                                  int32_t *counts, int max_rows, void *ws, size_t ws_bytes,
                                  void *stream_)
{
    SQ_REQUIRE(h && mask && table && counts && ws, SQ_EINVAL, "label: null pointer");
    Dims dm;
    SQ_TRY(make_dims(n, d, hgt, wid, max_rows, mask, &dm));
    Workspace w = carve(ws, ws_bytes, dm, max_rows);
    SQ_REQUIRE(w.bytes <= ws_bytes, SQ_ENOMEM, "label: workspace %zu < %zu bytes", ws_bytes, w.bytes);
    cudaStream_t st = (cudaStream_t)stream_;

    SQ_CUDA(cudaMemsetAsync(w.status, 0, w.zero_bytes, st));
    if (labels) SQ_CUDA(cudaMemsetAsync(labels, 0, (size_t)n * dm.vox * sizeof(int32_t), st));
    run_scan_emit<<<dim3(w.nblk, n), 256, 0, st>>>(mask, w.status, w.ticket, w.seg_off, w.totals, w.run_start, w.run_end,
                                                   w.parent, dm, w.nblk);
    SQ_CHECK_LAUNCH();
    // O(runs) kernels: the run count lives on the device, so launch a fixed grid and stride
    const int rblocks = std::min(dm.nchunks * (CHUNK / 256), 4 * h->sm_count);
    const int cblocks = std::min(dm.nchunks, std::max(1, 2 * h->sm_count / n));
    run_merge<<<dim3(rblocks, n), 256, 0, st>>>(mask, w.seg_off, w.totals, w.run_start, w.run_end,
                                                w.parent, dm);
    run_compress<<<dim3(cblocks, n), CHUNK_THREADS, 0, st>>>(w.parent, w.totals, w.chunk_count, dm);
    root_emit<<<dim3(cblocks, n), CHUNK_THREADS, 0, st>>>(w.parent, w.totals, w.chunk_count, w.nroots, w.rootlist, dm,
                                                         max_rows);
    ccl_order<<<n, 256, 0, st>>>(mask, w.run_start, w.parent, w.rootlist, w.nroots, w.sorted_root,
                                 w.class_base, dm, max_rows);
    SQ_CHECK_LAUNCH();
    run_accum<<<dim3(rblocks, n), 256, 0, st>>>(mask, w.totals, w.run_start, w.run_end, w.parent,
                                                w.class_base, w.acc, labels, dm, max_rows);
    ccl_finalize<<<dim3(sq_div_up(max_rows, 256), n), 256, 0, st>>>(
        mask, w.run_start, w.sorted_root, w.acc, w.nroots, table, counts, dm, max_rows, frame0);
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

extern "C" int sq_label_centroids_host(sq_handle_t h, const uint8_t *mask_host, int n, int d,
                                       int hgt, int wid, int frame0, int32_t *labels_host,
                                       float *table_host, int32_t *counts_host, int max_rows)
{
    SQ_REQUIRE(h && mask_host && table_host && counts_host, SQ_EINVAL, "label_host: null pointer");
    SqHostCall call(h);
    Dims dm;
    SQ_TRY(make_dims(n, d, hgt, wid, max_rows, nullptr, &dm));
    SQ_CUDA(cudaSetDevice(h->device));
    const size_t nvox = (size_t)n * dm.vox;
    SqArena probe(nullptr, 0);
    probe.take<uint8_t>(nvox);
    if (labels_host) probe.take<int32_t>(nvox);
    probe.take<float>((size_t)n * max_rows * 5);
    probe.take<int32_t>(n);
    const size_t ws_bytes = carve(nullptr, 0, dm, max_rows).bytes;
    SQ_TRY(sq_reserve_device(h, probe.off + ws_bytes + 256));
    SqArena a(h->dev_arena, h->dev_arena_bytes);
    uint8_t *mask = a.take<uint8_t>(nvox);
    int32_t *labels = labels_host ? a.take<int32_t>(nvox) : nullptr;
    float *table = a.take<float>((size_t)n * max_rows * 5);
    int32_t *counts = a.take<int32_t>(n);
    void *ws = a.take<char>(ws_bytes);
    cudaStream_t st = h->stream;
    SQ_CUDA(cudaMemcpyAsync(mask, mask_host, nvox, cudaMemcpyHostToDevice, st));
    SQ_TRY(sq_label_centroids(h, mask, n, d, hgt, wid, frame0, labels, table, counts, max_rows,
                              ws, ws_bytes, st));
    SQ_CUDA(cudaMemcpyAsync(counts_host, counts, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SQ_CUDA(cudaMemcpyAsync(table_host, table, (size_t)n * max_rows * 5 * sizeof(float),
                            cudaMemcpyDeviceToHost, st));
    if (labels_host)
        SQ_CUDA(cudaMemcpyAsync(labels_host, labels, nvox * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, st));
    SQ_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n; ++i)
        SQ_REQUIRE(counts_host[i] <= max_rows, SQ_EOVERFLOW,
                   "label: frame %d has %d components > max_rows=%d", i, counts_host[i], max_rows);
    return SQ_OK;
}
