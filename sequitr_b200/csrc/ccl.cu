// Label-and-localise on the GPU: union-find connected-component labelling with
// centroid reduction.  Replaces the per-frame / per-class SciPy loop of
// utils.CentroidWriter.write (reference utils.py:531-566: ndimage.label :547,
// center_of_mass :550).  Integer work, bit-exact against the oracle.
//
// One pass handles ALL classes: two voxels are connected when they share a face
// and carry the same non-zero value, so label(out == c) for every c falls out of
// a single union-find whose roots are the minimum linear index of each component
// (= the component's first voxel in raster order = SciPy's numbering order).
//
// Kernels (HBM-bound; algorithmic traffic 1 B/voxel in, rows out):
//   ccl_init      run-start labels inside 32-voxel row segments (warp ballot)
//   ccl_merge     unions across segment / row / plane borders (atomicMin union-find)
//   ccl_compress  full path compression + root count per 2048-voxel chunk
//   ccl_scan      per-frame exclusive scan of the chunk counts
//   ccl_emit      raster-ordered root list
//   ccl_order     stable counting sort of the roots by class -> row index
//   ccl_accum     warp-aggregated 64-bit sums (count, sum z, sum y, sum x) per row
//   ccl_finalize  fp64 divide exactly as center_of_mass does, float32 rows
#include "sq_common.cuh"

namespace {

constexpr int CHUNK = 2048;          // voxels per compress/emit block
constexpr int CHUNK_THREADS = 256;

struct Dims {
    int n, D, H, W;
    int vox;                          // D*H*W  (< 2^31)
    int nchunks;
};

__device__ __forceinline__ int find_root(const int *L, int i)
{
    int p = __ldcg(L + i);
    while (p != i) {
        i = p;
        p = __ldcg(L + i);
    }
    return i;
}

// lock-free union by minimum index (Playne/Komura style)
__device__ __forceinline__ void unite(int *L, int a, int b)
{
    bool done = false;
    while (!done) {
        a = find_root(L, a);
        b = find_root(L, b);
        if (a < b) {
            int old = atomicMin(L + b, a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            int old = atomicMin(L + a, b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    }
}

// grid (ceil(W/32), ceil(H/8), n*D), block (32, 8): one warp = 32 consecutive x
__global__ void ccl_init(const uint8_t *__restrict__ mask, int *__restrict__ L, Dims dm)
{
    const int lane = threadIdx.x;
    const int x = blockIdx.x * 32 + lane;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int zf = blockIdx.z;                    // frame * D + z
    const int z = zf % dm.D;
    const bool inside = (x < dm.W) && (y < dm.H);
    const long long g = ((long long)zf * dm.H + y) * dm.W + x;
    const uint8_t v = inside ? mask[g] : (uint8_t)0;
    const uint8_t lv = (uint8_t)__shfl_up_sync(0xffffffffu, (int)v, 1);
    const bool cont = (lane > 0) && (v != 0) && (lv == v);
    const unsigned bits = __ballot_sync(0xffffffffu, cont);
    if (!inside) return;
    const unsigned m = ~bits & ((2u << lane) - 1u);
    const int s = 31 - __clz(m);
    const int idx = (z * dm.H + y) * dm.W + x;
    if (v) L[g] = idx - (lane - s);            // background entries of L are never read
}

__global__ void ccl_merge(const uint8_t *__restrict__ mask, int *__restrict__ L, Dims dm)
{
    const int lane = threadIdx.x;
    const int x = blockIdx.x * 32 + lane;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int zf = blockIdx.z;
    const int z = zf % dm.D;
    if (x >= dm.W || y >= dm.H) return;
    const long long fo = (long long)(zf / dm.D) * dm.vox;      // frame offset
    const uint8_t *mk = mask + fo;
    int *Lf = L + fo;
    const int idx = (z * dm.H + y) * dm.W + x;
    const uint8_t v = mk[idx];
    if (!v) return;
    const bool left_same = (x > 0) && (mk[idx - 1] == v);
    if (lane == 0 && left_same) unite(Lf, idx, idx - 1);
    if (y > 0 && mk[idx - dm.W] == v) {
        const bool covered = left_same && (mk[idx - dm.W - 1] == v);
        if (!covered) unite(Lf, idx, idx - dm.W);
    }
    if (z > 0) {
        const int plane = dm.H * dm.W;
        if (mk[idx - plane] == v) {
            const bool covered = left_same && (mk[idx - plane - 1] == v);
            if (!covered) unite(Lf, idx, idx - plane);
        }
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

// grid (nchunks, n), block 256
__global__ void ccl_compress(const uint8_t *__restrict__ mask, int *__restrict__ L,
                             int *__restrict__ chunk_count, Dims dm)
{
    __shared__ int sh[8];
    int *Lf = L + (long long)blockIdx.y * dm.vox;
    const uint8_t *mk = mask + (long long)blockIdx.y * dm.vox;
    const int base = blockIdx.x * CHUNK;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < CHUNK / CHUNK_THREADS; ++k) {
        const int i = base + k * CHUNK_THREADS + threadIdx.x;
        if (i < dm.vox && mk[i]) {
            const int l = Lf[i];
            const int r = find_root(Lf, i);
            if (r != l) Lf[i] = r;
            cnt += (r == i);
        }
    }
    const int tot = block_sum_256(cnt, sh);
    if (threadIdx.x == 0) chunk_count[blockIdx.y * dm.nchunks + blockIdx.x] = tot;
}

// grid (n), block 1024: exclusive scan of chunk counts of one frame
__global__ void ccl_scan(const int *__restrict__ chunk_count, int *__restrict__ chunk_off,
                         int *__restrict__ totals, Dims dm)
{
    __shared__ int warp_sums[32];
    __shared__ int carry_sh;
    const int f = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (int start = 0; start < dm.nchunks; start += 1024) {
        const int i = start + threadIdx.x;
        const int v = (i < dm.nchunks) ? chunk_count[f * dm.nchunks + i] : 0;
        int s = v;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_sums[wid] = s;
        __syncthreads();
        if (wid == 0) {
            int w = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;                               // inclusive over warps
        }
        __syncthreads();
        const int carry = carry_sh;
        const int incl = s + (wid > 0 ? warp_sums[wid - 1] : 0);
        if (i < dm.nchunks) chunk_off[f * dm.nchunks + i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_sh = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[f] = carry_sh;
}

// grid (nchunks, n), block 256: thread t owns 8 consecutive voxels (keeps raster order)
__global__ void ccl_emit(const uint8_t *__restrict__ mask, const int *__restrict__ L,
                         const int *__restrict__ chunk_count, const int *__restrict__ chunk_off,
                         int *__restrict__ rootlist, Dims dm, int max_rows)
{
    __shared__ int warp_sums[8];
    const int f = blockIdx.y;
    if (chunk_count[f * dm.nchunks + blockIdx.x] == 0) return;
    const int *Lf = L + (long long)f * dm.vox;
    const uint8_t *mk = mask + (long long)f * dm.vox;
    const int base = blockIdx.x * CHUNK + threadIdx.x * 8;
    int flags = 0, cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = base + k;
        if (i < dm.vox && mk[i] && Lf[i] == i) { flags |= 1 << k; ++cnt; }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int s = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    if (lane == 31) warp_sums[wid] = s;
    __syncthreads();
    int pre = 0;
    for (int w = 0; w < wid; ++w) pre += warp_sums[w];
    int pos = chunk_off[f * dm.nchunks + blockIdx.x] + pre + s - cnt;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (flags & (1 << k)) {
            if (pos < max_rows) rootlist[(long long)f * max_rows + pos] = base + k;
            ++pos;
        }
}

// grid (n), block 256: stable counting sort of the frame's roots by class value.
// Writes L[root] = -2 - row, sorted_root, class_base.  The class histogram uses shared-memory
// atomics; the stable rank of root j is base[class] + #(i < j with the same class).  For up to
// ORDER_PAR roots every thread counts its predecessors in parallel (O(n^2/256), a few
// microseconds for the hundreds of cells of a real frame); beyond that thread c walks the
// raster-ordered list for class c.
constexpr int ORDER_PAR = 4096;

__global__ void ccl_order(const uint8_t *__restrict__ mask, int *__restrict__ L,
                          const int *__restrict__ rootlist, const int *__restrict__ totals,
                          int *__restrict__ sorted_root, int *__restrict__ class_base,
                          Dims dm, int max_rows)
{
    __shared__ uint8_t cls[ORDER_PAR];
    __shared__ int cnt_sh[256];
    const int f = blockIdx.x;
    const int c = threadIdx.x;
    const int n = min(totals[f], max_rows);
    const uint8_t *mk = mask + (long long)f * dm.vox;
    int *Lf = L + (long long)f * dm.vox;
    const int *rl = rootlist + (long long)f * max_rows;
    int *sr = sorted_root + (long long)f * max_rows;

    cnt_sh[c] = 0;
    __syncthreads();
    for (int j = c; j < n; j += 256) {
        const uint8_t v = mk[rl[j]];
        if (j < ORDER_PAR) cls[j] = v;
        atomicAdd(&cnt_sh[v], 1);
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
            const uint8_t v = cls[j];
            int rank = 0;
            for (int i = 0; i < j; ++i) rank += (cls[i] == v);
            const int row = cnt_sh[v] + rank;
            const int r = rl[j];
            Lf[r] = -2 - row;
            sr[row] = r;
        }
    } else if (mine > 0) {
        int k = cnt_sh[c];
        for (int j = 0; j < n; ++j) {
            const int r = rl[j];
            if (mk[r] == c) { Lf[r] = -2 - k; sr[k] = r; ++k; }
        }
    }
}

__device__ __forceinline__ int sum_of_set_bit_positions(unsigned g)
{
    return __popc(g & 0xAAAAAAAAu) + 2 * __popc(g & 0xCCCCCCCCu) + 4 * __popc(g & 0xF0F0F0F0u) +
           8 * __popc(g & 0xFF00FF00u) + 16 * __popc(g & 0xFFFF0000u);
}

// same launch shape as ccl_init
__global__ void ccl_accum(const uint8_t *__restrict__ mask, const int *__restrict__ L,
                          const int *__restrict__ class_base, unsigned long long *__restrict__ acc,
                          int *__restrict__ labels_out, Dims dm, int max_rows)
{
    const int lane = threadIdx.x;
    const int x = blockIdx.x * 32 + lane;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int zf = blockIdx.z;
    const int z = zf % dm.D, f = zf / dm.D;
    const bool inside = (x < dm.W) && (y < dm.H);
    const long long fo = (long long)f * dm.vox;
    const int idx = (z * dm.H + y) * dm.W + x;
    int row = -1;
    uint8_t v = 0;
    if (inside) {
        v = mask[fo + idx];
        if (v) {
            const int l = L[fo + idx];
            if (l <= -2) row = -2 - l;
            else if (l >= 0) {
                const int rr = L[fo + l];
                if (rr <= -2) row = -2 - rr;
            }
        }
        if (labels_out) labels_out[fo + idx] = (row >= 0) ? row - class_base[f * 256 + v] + 1 : 0;
    }
    const unsigned act = __ballot_sync(0xffffffffu, row >= 0);
    if (row >= 0) {
        const unsigned g = __match_any_sync(act, row);
        if (lane == __ffs(g) - 1) {
            const unsigned long long cnt = __popc(g);
            unsigned long long *a = acc + ((long long)f * max_rows + row) * 4;
            const int x0 = blockIdx.x * 32;
            atomicAdd(a + 0, cnt);
            if (dm.D > 1) atomicAdd(a + 1, cnt * (unsigned long long)z);
            atomicAdd(a + 2, cnt * (unsigned long long)y);
            atomicAdd(a + 3, cnt * (unsigned long long)x0 +
                                 (unsigned long long)sum_of_set_bit_positions(g));
        }
    }
}

// grid (ceil(max_rows/256), n)
__global__ void ccl_finalize(const uint8_t *__restrict__ mask, const int *__restrict__ sorted_root,
                             const unsigned long long *__restrict__ acc,
                             const int *__restrict__ totals, float *__restrict__ table,
                             int *__restrict__ counts, Dims dm, int max_rows, int frame0)
{
    const int f = blockIdx.y;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int tot = totals[f];
    if (row == 0) counts[f] = tot;
    if (row >= min(tot, max_rows)) return;
    const int r = sorted_root[(long long)f * max_rows + row];
    const unsigned long long c = mask[(long long)f * dm.vox + r];
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
    int *L, *chunk_count, *chunk_off, *totals, *rootlist, *sorted_root, *class_base;
    unsigned long long *acc;
    size_t bytes;
};

Workspace carve(void *p, size_t avail, const Dims &dm, int max_rows)
{
    SqArena a(p, avail);
    Workspace w;
    w.L = a.take<int>((size_t)dm.n * dm.vox);
    w.chunk_count = a.take<int>((size_t)dm.n * dm.nchunks);
    w.chunk_off = a.take<int>((size_t)dm.n * dm.nchunks);
    w.totals = a.take<int>(dm.n);
    w.rootlist = a.take<int>((size_t)dm.n * max_rows);
    w.sorted_root = a.take<int>((size_t)dm.n * max_rows);
    w.class_base = a.take<int>((size_t)dm.n * 256);
    w.acc = a.take<unsigned long long>((size_t)dm.n * max_rows * 4);
    w.bytes = a.off;
    return w;
}

int make_dims(int n, int d, int hgt, int wid, int max_rows, Dims *dm)
{
    SQ_REQUIRE(n >= 1 && d >= 1 && hgt >= 1 && wid >= 1, SQ_EINVAL,
               "label: bad shape (%d,%d,%d,%d)", n, d, hgt, wid);
    SQ_REQUIRE(max_rows >= 1, SQ_EINVAL, "label: max_rows must be >= 1");
    const long long vox = (long long)d * hgt * wid;
    SQ_REQUIRE(vox < (1ll << 31) - CHUNK, SQ_EINVAL, "label: frame too large (%lld voxels)", vox);
    SQ_REQUIRE(d <= 65535, SQ_EINVAL, "label: depth %d > 65535", d);
    dm->n = n; dm->D = d; dm->H = hgt; dm->W = wid;
    dm->vox = (int)vox;
    dm->nchunks = sq_div_up(vox, CHUNK);
    return SQ_OK;
}

}  // namespace

extern "C" int sq_label_workspace_bytes(sq_handle_t h, int n, int d, int hgt, int wid,
                                        int max_rows, size_t *bytes)
{
    SQ_REQUIRE(h && bytes, SQ_EINVAL, "label: null handle/pointer");
    Dims dm;
    SQ_TRY(make_dims(n, d, hgt, wid, max_rows, &dm));
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
    SQ_TRY(make_dims(n, d, hgt, wid, max_rows, &dm));
    Workspace w = carve(ws, ws_bytes, dm, max_rows);
    SQ_REQUIRE(w.bytes <= ws_bytes, SQ_ENOMEM, "label: workspace %zu < %zu bytes", ws_bytes, w.bytes);
    cudaStream_t st = (cudaStream_t)stream_;

    // planes go on grid.z (max 65535): split very deep batches
    const dim3 blk(32, 8);
    const int planes = n * d;
    SQ_CUDA(cudaMemsetAsync(w.acc, 0, (size_t)n * max_rows * 4 * sizeof(unsigned long long), st));
    for (int p0 = 0; p0 < planes; p0 += 65535 / d * d) {
        // chunks of whole frames so that frame = zf / D stays valid
        const int np = min(planes - p0, 65535 / d * d);
        Dims sub = dm;
        const long long off = (long long)(p0 / d) * dm.vox;
        sub.n = np / d;
        const dim3 grid(sq_div_up(wid, 32), sq_div_up(hgt, 8), np);
        ccl_init<<<grid, blk, 0, st>>>(mask + off, w.L + off, sub);
        ccl_merge<<<grid, blk, 0, st>>>(mask + off, w.L + off, sub);
    }
    SQ_CHECK_LAUNCH();
    ccl_compress<<<dim3(dm.nchunks, n), CHUNK_THREADS, 0, st>>>(mask, w.L, w.chunk_count, dm);
    ccl_scan<<<n, 1024, 0, st>>>(w.chunk_count, w.chunk_off, w.totals, dm);
    ccl_emit<<<dim3(dm.nchunks, n), CHUNK_THREADS, 0, st>>>(mask, w.L, w.chunk_count, w.chunk_off,
                                                           w.rootlist, dm, max_rows);
    ccl_order<<<n, 256, 0, st>>>(mask, w.L, w.rootlist, w.totals, w.sorted_root, w.class_base,
                                 dm, max_rows);
    SQ_CHECK_LAUNCH();
    for (int p0 = 0; p0 < planes; p0 += 65535 / d * d) {
        const int np = min(planes - p0, 65535 / d * d);
        Dims sub = dm;
        const int f0 = p0 / d;
        const long long off = (long long)f0 * dm.vox;
        sub.n = np / d;
        const dim3 grid(sq_div_up(wid, 32), sq_div_up(hgt, 8), np);
        ccl_accum<<<grid, blk, 0, st>>>(mask + off, w.L + off, w.class_base + f0 * 256,
                                        w.acc + (long long)f0 * max_rows * 4,
                                        labels ? labels + off : nullptr, sub, max_rows);
    }
    ccl_finalize<<<dim3(sq_div_up(max_rows, 256), n), 256, 0, st>>>(
        mask, w.sorted_root, w.acc, w.totals, table, counts, dm, max_rows, frame0);
    SQ_CHECK_LAUNCH();
    return SQ_OK;
}

extern "C" int sq_label_centroids_host(sq_handle_t h, const uint8_t *mask_host, int n, int d,
                                       int hgt, int wid, int frame0, int32_t *labels_host,
                                       float *table_host, int32_t *counts_host, int max_rows)
{
    SQ_REQUIRE(h && mask_host && table_host && counts_host, SQ_EINVAL, "label_host: null pointer");
    Dims dm;
    SQ_TRY(make_dims(n, d, hgt, wid, max_rows, &dm));
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
