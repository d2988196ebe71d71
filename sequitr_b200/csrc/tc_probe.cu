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
    printf("unknown case %s\n", which);
    return 1;
}
