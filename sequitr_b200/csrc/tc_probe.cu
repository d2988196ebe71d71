// Stand-alone hardware probe for the tcgen05 / TMA conventions the convolution
// kernels rely on (run on a B200: `tc_probe <case>`).  Not part of the library.
//
// The host builds a byte image of shared memory plus a list of (a_off, b_off)
// descriptor offsets; the kernel copies the image into shared memory, issues one
// tcgen05.mma per list entry into one TMEM accumulator and dumps D.  The host then
// evaluates several layout hypotheses and reports which one reproduces D exactly.
#include "tc_common.cuh"
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e = (x);                                                          \
        if (e != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(2);                                                                  \
        }                                                                             \
    } while (0)

struct MmaList {
    int n;
    uint32_t a_off[40], b_off[40];
};

__global__ void __launch_bounds__(128, 1)
probe_mma(const uint8_t *img, int img_bytes, MmaList L, uint32_t lbo_a, uint32_t sbo_a,
          uint32_t lbo_b, uint32_t sbo_b, int N, float *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid * 16; i < img_bytes; i += 128 * 16)
        *(uint4 *)(smem + i) = *(const uint4 *)(img + i);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (warp == 0) { tc::tmem_alloc(&tmem_base, 256); tc::tmem_relinquish(); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        const uint32_t base = tc::smem_u32(smem);
        const uint32_t idesc = tc::instr_desc_bf16(128, N);
        for (int i = 0; i < L.n; ++i)
            tc::umma_bf16(tb, tc::smem_desc(base + L.a_off[i], lbo_a, sbo_a),
                          tc::smem_desc(base + L.b_off[i], lbo_b, sbo_b), idesc, i > 0);
        tc::umma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    for (int c = 0; c < N; c += 8) {
        uint32_t v[8];
        tc::tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + c, v);
        tc::tmem_ld_wait();
        for (int j = 0; j < 8; ++j) out[(size_t)tid * N + c + j] = __uint_as_float(v[j]);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, 256);
}

__global__ void probe_tma(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int c3,
                          int bytes, uint8_t *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_barrier_init();
        tc::mbar_arrive_expect_tx(&bar, bytes);
        tc::tma_load_4d(smem, &map, &bar, c0, c1, c2, c3);
    }
    __syncthreads();
    tc::mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

// MMA issue-rate probe: one thread issues `reps` back-to-back 128 x N x 16 MMAs (operands in shared
// memory, `ntap` different A start offsets round-robin, accumulating into `nacc` accumulators
// round-robin) and times them with clock64: the cost model behind DESIGN.md section 4.
__global__ void __launch_bounds__(128, 4)
probe_rate(int N, int reps, uint32_t lbo_a, uint32_t sbo_a, int ntap, uint32_t tap_stride, int nacc,
           int a_bytes, long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid * 4; i < a_bytes + 8192; i += 128 * 4)
        *(uint32_t *)(smem + i) = 0x3c003c00u + (uint32_t)(i * 2654435761u >> 28);   // small bf16 values
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (warp == 0) { tc::tmem_alloc(&tmem_base, 128); tc::tmem_relinquish(); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        const uint32_t base = tc::smem_u32(smem);
        const uint32_t idesc = tc::instr_desc_bf16(128, N);
        const uint64_t bdesc = tc::smem_desc(base + a_bytes, N * 16, 128);
        const long long t0 = clock64();
        // descriptors precomputed; 6 MMAs per loop trip (taps round-robin, accumulators alternate)
        uint64_t ad[3];
        for (int k = 0; k < 3; ++k) ad[k] = tc::smem_desc(base + (k % ntap) * tap_stride, lbo_a, sbo_a);
        const uint32_t d1 = tb + (nacc > 1 ? N : 0);
        for (int i = 0; i < reps; i += 6) {
            tc::umma_bf16(tb, ad[0], bdesc, idesc, 1);
            tc::umma_bf16(d1, ad[1], bdesc, idesc, 1);
            tc::umma_bf16(tb, ad[2], bdesc, idesc, 1);
            tc::umma_bf16(d1, ad[0], bdesc, idesc, 1);
            tc::umma_bf16(tb, ad[1], bdesc, idesc, 1);
            tc::umma_bf16(d1, ad[2], bdesc, idesc, 1);
        }
        tc::umma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, 128);
}

// TMA load-rate probe: the conv kernels' producer loop alone (5-D box loads into an smem ring, a
// consumer thread that frees each stage as soon as it lands), to measure bytes/clk/SM per box shape.
__global__ void __launch_bounds__(160, 4)
probe_tma_rate(const __grid_constant__ CUtensorMap map, int box_bytes, int nst, int tiles_x, int tiles_y,
               int nimg, int step_x, int step_y, int xoff, int ksteps, int nprod, long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full_bar[16], empty_bar[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int ntiles = nimg * tiles_x * tiles_y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = clock64();
    if (warp < nprod && lane == 0) {
        // producer warp `warp` issues the loads of global steps i = warp, warp + nprod, ...
        long long i = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, n = t / (tiles_x * tiles_y);
            for (int ks = 0; ks < ksteps; ++ks, ++i) {
                if ((int)(i % nprod) != warp) continue;
                const int stage = (int)(i % nst);
                const uint32_t phase = (uint32_t)((i / nst) & 1);
                tc::mbar_wait(&empty_bar[stage], phase ^ 1);
                tc::mbar_arrive_expect_tx(&full_bar[stage], box_bytes);
                tc::tma_load_5d(smem + (size_t)stage * box_bytes, &map, &full_bar[stage], (tx * step_x + xoff) * 8,
                                ty * step_y - 1, ks * 2, 0, n);
            }
        }
    } else if (warp == 4 && lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
            for (int ks = 0; ks < ksteps; ++ks) {
                tc::mbar_wait(&full_bar[stage], phase);
                tc::mbar_arrive(&empty_bar[stage]);
                if (++stage == nst) { stage = 0; phase ^= 1; }
            }
        out[blockIdx.x] = clock64() - t0;
    }
}

// TMEM read-rate probe: every warp of the CTA streams its lane quarter of the 512 TMEM columns with
// tcgen05.ld 32x32b.x16 (or .x32 via two back-to-back x16 on adjacent columns), `depth` loads in flight
// before each wait.
template <int DEPTH>
__global__ void __launch_bounds__(512, 1)
probe_ldtm(int reps, long long *out, uint32_t *sink)
{
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tc::tmem_alloc(&tmem_base, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tb = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        uint32_t v[DEPTH][16];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) tc::tmem_ld16(tb + ((r * DEPTH + d) * 16) % 512, v[d]);
        tc::tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) acc ^= v[d][0] ^ v[d][15];
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

// Do tcgen05.mma and tcgen05.ld overlap?  Warp 0 lane 0 issues `reps` MMAs (N columns at TMEM col 0..),
// warps 4..4+ldw-1 meanwhile stream tcgen05.ld from columns 256.. of the same CTA's TMEM.
__global__ void __launch_bounds__(384, 1)
probe_overlap(int N, int reps, int ldw, int ldreps, long long *out, uint32_t *sink)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid * 4; i < 16384; i += 384 * 4) *(uint32_t *)(smem + i) = 0x3c003c00u;
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (warp == 0) { tc::tmem_alloc(&tmem_base, 512); tc::tmem_relinquish(); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tb = tmem_base;
    const long long t0 = clock64();
    if (tid == 0) {
        const uint32_t base = tc::smem_u32(smem);
        const uint32_t idesc = tc::instr_desc_bf16(128, N);
        const uint64_t ad = tc::smem_desc(base, 2048, 128), bd = tc::smem_desc(base + 4096, N * 16, 128);
        for (int i = 0; i < reps; i += 4) {
            tc::umma_bf16(tb, ad, bd, idesc, 1);
            tc::umma_bf16(tb + 128, ad, bd, idesc, 1);
            tc::umma_bf16(tb, ad, bd, idesc, 1);
            tc::umma_bf16(tb + 128, ad, bd, idesc, 1);
        }
        tc::umma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        out[blockIdx.x * 2] = clock64() - t0;
    } else if (warp >= 4 && warp < 4 + ldw) {
        uint32_t acc = 0;
        const uint32_t lb = tb + 256 + ((uint32_t)((warp & 3) * 32) << 16);
        for (int r = 0; r < ldreps; ++r) {
            uint32_t v0[16], v1[16];
            tc::tmem_ld16(lb + (r * 32) % 256, v0);
            tc::tmem_ld16(lb + (r * 32 + 16) % 256, v1);
            tc::tmem_ld_wait();
            acc ^= v0[3] ^ v1[7];
        }
        if (acc == 0x12345678u) sink[0] = acc;
        if (tid == 128) out[blockIdx.x * 2 + 1] = clock64() - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, 512);
}

static uint16_t f2bf(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return (uint16_t)(u >> 16);          // inputs are small integers: exact
}
static float bf2f(uint16_t h)
{
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// element (r, k) of a K-major no-swizzle operand under hypothesis `hyp`
//   hyp 0: K-chunk stride = LBO, row-group stride = SBO   (expected)
//   hyp 1: roles swapped
static float elem(const std::vector<uint8_t> &img, uint32_t off, uint32_t lbo, uint32_t sbo, int r,
                  int k, int hyp)
{
    const uint32_t kc = k / 8, rg = r / 8;
    const size_t o = off + (hyp == 0 ? kc * lbo + rg * sbo : kc * sbo + rg * lbo) + (r % 8) * 16 +
                     (k % 8) * 2;
    if (o + 2 > img.size()) return 1e30f;
    uint16_t h;
    memcpy(&h, &img[o], 2);
    return bf2f(h);
}

static int run_mma_case(const char *name, std::vector<uint8_t> img, const MmaList &L, uint32_t lbo_a,
                        uint32_t sbo_a, uint32_t lbo_b, uint32_t sbo_b, int N)
{
    while (img.size() % 16) img.push_back(0);
    uint8_t *dimg;
    float *dout;
    CK(cudaMalloc(&dimg, img.size()));
    CK(cudaMalloc(&dout, 128 * N * sizeof(float)));
    CK(cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    probe_mma<<<1, 128, img.size() + 1024>>>(dimg, (int)img.size(), L, lbo_a, sbo_a, lbo_b, sbo_b, N,
                                             dout);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * N);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    int verdict = -1;
    for (int hyp = 0; hyp < 2; ++hyp) {
        double maxerr = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < N; ++n) {
                double acc = 0;
                for (int i = 0; i < L.n; ++i)
                    for (int k = 0; k < 16; ++k)
                        acc += (double)elem(img, L.a_off[i], lbo_a, sbo_a, r, k, hyp) *
                               (double)elem(img, L.b_off[i], lbo_b, sbo_b, n, k, hyp);
                double e = fabs(acc - out[(size_t)r * N + n]);
                if (e > maxerr) maxerr = e;
            }
        printf("PROBE %s hyp%d maxerr=%g\n", name, hyp, maxerr);
        if (maxerr == 0 && verdict < 0) verdict = hyp;
    }
    printf("PROBE %s VERDICT %s\n", name, verdict == 0 ? "OK(lbo=K,sbo=MN)" : verdict == 1 ? "SWAPPED" : "MISMATCH");
    printf("PROBE %s sample D[0][0..3]= %g %g %g %g  D[9][1]=%g D[127][N-1]=%g\n", name, out[0], out[1],
           out[2], out[3], out[9 * N + 1], out[127 * N + N - 1]);
    return verdict == 0 ? 0 : 1;
}

static void fill_random(std::vector<uint8_t> &img, size_t off, size_t bytes, unsigned seed)
{
    if (img.size() < off + bytes) img.resize(off + bytes, 0);
    unsigned s = seed * 2654435761u + 12345u;
    for (size_t i = 0; i < bytes; i += 2) {
        s = s * 1664525u + 1013904223u;
        const float v = (float)((int)((s >> 24) % 7) - 3);
        const uint16_t h = f2bf(v);
        memcpy(&img[off + i], &h, 2);
    }
}

int main(int argc, char **argv)
{
    const char *which = argc > 1 ? argv[1] : "gemm";
    int dev = 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("PROBE device %s sm_%d%d smem/block optin %zu\n", prop.name, prop.major, prop.minor,
           (size_t)prop.sharedMemPerBlockOptin);

    if (!strcmp(which, "gemm") || !strcmp(which, "gemm256")) {
        // packed K-major operands: [kchunk][rowgroup][8 rows][16 B]
        const int N = !strcmp(which, "gemm256") ? 256 : 32, K = 64;
        const uint32_t lbo_a = 128 * 16, sbo_a = 128, lbo_b = N * 16, sbo_b = 128;
        std::vector<uint8_t> img;
        const uint32_t a0 = 0, b0 = (K / 8) * lbo_a;
        fill_random(img, a0, (K / 8) * lbo_a, 1);
        fill_random(img, b0, (K / 8) * lbo_b, 2);
        MmaList L;
        L.n = K / 16;
        for (int i = 0; i < L.n; ++i) { L.a_off[i] = a0 + i * 2 * lbo_a; L.b_off[i] = b0 + i * 2 * lbo_b; }
        return run_mma_case(which, img, L, lbo_a, sbo_a, lbo_b, sbo_b, N);
    }
    if (!strcmp(which, "conv16") || !strcmp(which, "conv64")) {
        // halo patch [C/8][PH][PW][8 ch]; tap (ky,kx) = start address shifted by (ky*PW+kx)*16 B
        const int C = !strcmp(which, "conv64") ? 64 : 16, N = C, PW = 10, PH = 18;
        const uint32_t lbo_a = PH * PW * 16, sbo_a = PW * 16, lbo_b = N * 16, sbo_b = 128;
        std::vector<uint8_t> img;
        const uint32_t a_bytes = (C / 8) * lbo_a;
        const uint32_t b0 = (a_bytes + 1023) / 1024 * 1024;
        fill_random(img, 0, a_bytes, 3);
        fill_random(img, b0, 9 * (C / 8) * lbo_b, 4);
        MmaList L;
        L.n = 0;
        for (int tap = 0; tap < 9; ++tap)
            for (int ks = 0; ks < C / 16; ++ks) {
                L.a_off[L.n] = ((tap / 3) * PW + (tap % 3)) * 16 + ks * 2 * lbo_a;
                L.b_off[L.n] = b0 + (tap * (C / 8) + ks * 2) * lbo_b;
                ++L.n;
            }
        return run_mma_case(which, img, L, lbo_a, sbo_a, lbo_b, sbo_b, N);
    }
    if (!strcmp(which, "tma")) {
        // global [N=2][CB=3][H=20][W=24][8] bf16 viewed as 4-D (W*8, H, CB, N)
        const int NB = 2, CB = 3, H = 20, W = 24, PW = 10, PH = 18, CBB = 2;
        std::vector<uint16_t> g((size_t)NB * CB * H * W * 8);
        for (size_t i = 0; i < g.size(); ++i) g[i] = f2bf((float)(1 + i % 251));
        uint16_t *dg;
        CK(cudaMalloc(&dg, g.size() * 2));
        CK(cudaMemcpy(dg, g.data(), g.size() * 2, cudaMemcpyHostToDevice));
        sq_encode_tiled_fn enc = sq_get_encode_tiled();
        if (!enc) { printf("PROBE tma: no cuTensorMapEncodeTiled\n"); return 1; }
        CUtensorMap map;
        cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)CB, (cuuint64_t)NB};
        cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)CB * H * W * 16};
        cuuint32_t box[4] = {PW * 8, PH, CBB, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dg, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("PROBE tma: encode failed %d\n", (int)r); return 1; }
        const int bytes = PW * 8 * PH * CBB * 2;
        uint8_t *dout;
        CK(cudaMalloc(&dout, bytes));
        int fails = 0;
        const int cases[4][4] = {{-1, -1, 0, 0}, {8, 4, 1, 1}, {15, 3, 2, 1}, {0, 15, 0, 1}};
        for (int t = 0; t < 4; ++t) {
            const int x0 = cases[t][0], y0 = cases[t][1], cb0 = cases[t][2], n = cases[t][3];
            CK(cudaMemset(dout, 0xAB, bytes));
            probe_tma<<<1, 128, bytes + 1024>>>(map, x0 * 8, y0, cb0, n, bytes, dout);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            std::vector<uint16_t> o(bytes / 2);
            CK(cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int cb = 0; cb < CBB; ++cb)
                for (int py = 0; py < PH; ++py)
                    for (int px = 0; px < PW; ++px)
                        for (int c = 0; c < 8; ++c) {
                            const int x = x0 + px, y = y0 + py, cc = cb0 + cb;
                            uint16_t want = 0;
                            if (x >= 0 && x < W && y >= 0 && y < H && cc < CB)
                                want = g[((((size_t)n * CB + cc) * H + y) * W + x) * 8 + c];
                            if (o[((cb * PH + py) * PW + px) * 8 + c] != want) ++bad;
                        }
            printf("PROBE tma case%d (x0=%d,y0=%d,cb0=%d,n=%d) mismatches=%d\n", t, x0, y0, cb0, n, bad);
            fails += bad != 0;
        }
        printf("PROBE tma VERDICT %s\n", fails ? "MISMATCH" : "OK");
        return fails;
    }
    if (!strcmp(which, "rate")) {
        long long *dout;
        CK(cudaMalloc(&dout, 1024 * sizeof(long long)));
        const int reps = 6 * 700;
        CK(cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024));
        struct { const char *name; uint32_t lbo, sbo; int ntap; uint32_t tstride; int a_bytes; } lay[3] = {
            {"9tap PW=10 (sbo 160)", 18 * 10 * 16, 160, 9, 16, 2 * 18 * 10 * 16 + 1024},
            {"9tap ILV   (sbo 320)", 34 * 10 * 16, 320, 9, 16, 2 * 34 * 10 * 16 + 1024},
            {"xc 3tap    (sbo 128)", 18 * 16 * 16, 128, 3, 256, 2 * 18 * 16 * 16 + 1024}};
        const int Ns[5] = {16, 32, 48, 64, 96};
        for (int l = 0; l < 3; ++l)
            for (int ni = 0; ni < 5; ++ni)
                for (int ctas = 1; ctas <= 4; ++ctas) {
                    const int N = Ns[ni], nacc = (N <= 64) ? 2 : 1;
                    const int grid = 148 * ctas;
                    const int smem = lay[l].a_bytes + 8192 + 1024;
                    probe_rate<<<grid, 128, smem>>>(N, reps, lay[l].lbo, lay[l].sbo, lay[l].ntap, lay[l].tstride,
                                                    nacc, lay[l].a_bytes, dout);
                    CK(cudaGetLastError());
                    CK(cudaDeviceSynchronize());
                    std::vector<long long> h(grid);
                    CK(cudaMemcpy(h.data(), dout, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                    long long mx = 0;
                    double avg = 0;
                    for (long long v : h) { mx = v > mx ? v : mx; avg += (double)v / grid; }
                    // per-SM cost of one MMA = elapsed / (reps * ctas co-resident)
                    printf("PROBE rate %-22s N=%3d ctas/SM=%d  clk/MMA per CTA %.1f  per SM %.1f  (tensor floor %.1f)\n",
                           lay[l].name, N, ctas, avg / reps, avg / reps / ctas, N / 2.0);
                }
        return 0;
    }
    if (!strcmp(which, "tmarate")) {
        // activation tensor [n=4][cb=4][2048][2048][8] bf16 = 537 MB (streams from HBM)
        const int NB = 4, CB = 4, H = 2048, W = 2048;
        uint16_t *dg;
        CK(cudaMalloc(&dg, (size_t)NB * CB * H * W * 16));
        CK(cudaMemset(dg, 0x3c, (size_t)NB * CB * H * W * 16));
        long long *dout;
        CK(cudaMalloc(&dout, 1024 * sizeof(long long)));
        sq_encode_tiled_fn enc = sq_get_encode_tiled();
        CK(cudaFuncSetAttribute(probe_tma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        struct { const char *name; int pw, ph, step_x, step_y, xoff; } sh[7] = {
            {"9tap S=4: 10 x 66 (x0-1)", 10, 66, 8, 64, -1},
            {"9tap S=2: 10 x 34 (x0-1)", 10, 34, 8, 32, -1},
            {"xc   S=2: 16 x 18 (x0-1)", 16, 18, 14, 16, -1},
            {"xc   S=1: 16 x 10 (x0-1)", 16, 10, 14, 8, -1},
            {"aligned : 16 x 18 (x0)", 16, 18, 16, 16, 0},
            {"wide    : 32 x 18 (x0)", 32, 18, 32, 16, 0},
            {"wide    : 32 x 34 (x0)", 32, 34, 32, 32, 0}};
        for (int i = 0; i < 7; ++i) {
            CUtensorMap map;
            cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)CB, 1, (cuuint64_t)NB};
            cuuint64_t strides[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)CB * H * W * 16,
                                     (cuuint64_t)CB * H * W * 16};
            cuuint32_t box[5] = {(cuuint32_t)sh[i].pw * 8, (cuuint32_t)sh[i].ph, 2, 1, 1};
            cuuint32_t es[5] = {1, 1, 1, 1, 1};
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dg, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            const int box_bytes = sh[i].pw * sh[i].ph * 2 * 16;
            const int tiles_x = (W + sh[i].step_x - 1) / sh[i].step_x, tiles_y = H / sh[i].step_y;
            for (int ctas = 1; ctas <= 2; ++ctas)
                for (int nprod = 1; nprod <= 4; nprod *= 2) {
                    int nst = 100 * 1024 / box_bytes / 4 * 4;          // a stage belongs to ONE producer
                    nst = nst > 16 ? 16 : (nst < 4 ? 4 : nst);
                    if (ctas == 2 && nst * box_bytes > 100 * 1024) continue;
                    const int grid = 148 * ctas, ksteps = 2;
                    probe_tma_rate<<<grid, 160, nst * box_bytes + 1024>>>(map, box_bytes, nst, tiles_x, tiles_y, NB,
                                                                          sh[i].step_x, sh[i].step_y, sh[i].xoff, ksteps,
                                                                          nprod, dout);
                    CK(cudaGetLastError());
                    CK(cudaDeviceSynchronize());
                    std::vector<long long> h(grid);
                    CK(cudaMemcpy(h.data(), dout, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                    long long mx = 0;
                    for (long long v : h) mx = v > mx ? v : mx;
                    const double bytes = (double)NB * tiles_x * tiles_y * ksteps * box_bytes;
                    printf("PROBE tmarate %-26s box %5d B  ctas/SM=%d stages=%2d producers=%d: %.1f B/clk/SM  (%.0f clk, %.2f GB moved)\n",
                           sh[i].name, box_bytes, ctas, nst, nprod, bytes / 148.0 / (double)mx, (double)mx, bytes / 1e9);
                }
        }
        return 0;
    }
    if (!strcmp(which, "ldtm")) {
        long long *dout;
        uint32_t *sink;
        CK(cudaMalloc(&dout, 148 * 16 * sizeof(long long)));
        CK(cudaMalloc(&sink, 4));
        const int reps = 2000;
        for (int depth = 1; depth <= 4; depth *= 2)
            for (int warps = 4; warps <= 16; warps *= 2) {
                if (depth == 1) probe_ldtm<1><<<148, warps * 32>>>(reps, dout, sink);
                else if (depth == 2) probe_ldtm<2><<<148, warps * 32>>>(reps, dout, sink);
                else probe_ldtm<4><<<148, warps * 32>>>(reps, dout, sink);
                CK(cudaGetLastError());
                CK(cudaDeviceSynchronize());
                std::vector<long long> h(148 * 16);
                CK(cudaMemcpy(h.data(), dout, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                long long mx = 0;
                for (int b = 0; b < 148; ++b) for (int w = 0; w < warps; ++w) mx = h[b * 16 + w] > mx ? h[b * 16 + w] : mx;
                const double bytes = (double)reps * depth * warps * 32 * 16 * 4;
                printf("PROBE ldtm 32x32b.x16  warps/SM=%2d loads in flight/warp=%d: %.1f B/clk/SM (%.1f clk per x16 load per warp)\n",
                       warps, depth, bytes / (double)mx, (double)mx / (reps * depth));
            }
        return 0;
    }
    if (!strcmp(which, "overlap")) {
        long long *dout;
        uint32_t *sink;
        CK(cudaMalloc(&dout, 148 * 2 * sizeof(long long)));
        CK(cudaMalloc(&sink, 4));
        CK(cudaFuncSetAttribute(probe_overlap, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
        const int reps = 4000;
        for (int N = 16; N <= 128; N *= 2)
            for (int ldw = 0; ldw <= 8; ldw += 4) {
                CK(cudaMemset(dout, 0, 148 * 2 * sizeof(long long)));
                probe_overlap<<<148, 384, 17 * 1024>>>(N, reps, ldw, 4000, dout, sink);
                CK(cudaGetLastError());
                CK(cudaDeviceSynchronize());
                std::vector<long long> h(148 * 2);
                CK(cudaMemcpy(h.data(), dout, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                double am = 0, al = 0;
                for (int b = 0; b < 148; ++b) { am += h[b * 2] / 148.0; al += h[b * 2 + 1] / 148.0; }
                printf("PROBE overlap N=%3d  ld warps=%d: %.1f clk/MMA; ld warps: %.1f clk per pair of x16 loads\n", N, ldw,
                       am / reps, ldw ? al / 4000 : 0.0);
            }
        return 0;
    }
    printf("unknown case %s\n", which);
    return 1;
}
