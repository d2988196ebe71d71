// The device-side image of the reference UNet object (networks/unet.py:53-342).
#pragma once
#include "sq_common.cuh"

struct SqHostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
};

// one conv-like layer of the graph in execution order
struct SqLayer {
    enum Kind { CONV = 0, POOL = 1, UPCONV = 2, ELTWISE = 3, HEAD = 4 } kind;
    std::string scope;            // TF variable scope, e.g. "UNet/down0/conv1"
    int level = 0;                // resolution level (0 = full resolution)
    int cin0 = 0, cin1 = 0;       // channels of the first / second (skip) input
    int cout = 0;
    int ksize = 3;
    double flops_per_px = 0;      // algorithmic FLOPs per pixel of this layer's OUTPUT grid
    // fp32-exact device weights
    float *w = nullptr, *scale = nullptr, *shift = nullptr;
    // tensor-core device weights (bf16, re-laid-out; see unet_tc.cu)
    void *w_tc = nullptr;
    void *w_xc = nullptr;         // x-combined layout for Cout <= 32 convs (conv_xc_kernel)
    // level-0 layers on the quad (space-to-depth) layout (conv_qd_kernel): weights re-laid-out as a 64-wide
    // half-resolution conv, scale / shift repeated for the four output parities
    void *w_qd = nullptr;
    void *w_qu = nullptr;         // up0/conv1 only: quad weights with permuted output columns (conv_qu_kernel)
    void *w_qf = nullptr;         // down0/conv2 only: first conv's B matrix + conv2 with permuted output columns (conv_qf_kernel)
    float *scale_q = nullptr, *shift_q = nullptr;
};

struct SqLayerTimer {
    std::vector<cudaEvent_t> ev;          // ev[i], ev[i+1] bracket layer i
    std::vector<const char *> names;
    std::vector<double> flops;
    bool enabled = false;
};

struct sq_unet_s {
    sq_handle_s *h = nullptr;
    int ndim = 2, cin = 1, nout = 2, nlev = 5, bridge = SQ_BRIDGE_CONCAT, mode = SQ_MODE_FP32_EXACT;
    std::vector<int> filters;
    std::map<std::string, SqHostTensor> host;   // as loaded (TF names / layouts)
    std::vector<SqLayer> layers;                // CONV / UPCONV / HEAD layers with weights
    bool finalized = false;
    int last_launches = 0;
    std::vector<void *> dev_allocs;             // freed in sq_unet_destroy
    void *tc_state = nullptr;                   // owned by unet_tc.cu
    SqLayerTimer timer;
    std::map<std::string, std::string> aux_names;   // stable storage for timer names of fused launches
};

// ---- tensor-core path (unet_tc.cu)
int sq_tc_finalize(sq_unet_s *u);
int sq_tc_destroy(sq_unet_s *u);
int sq_tc_workspace_bytes(sq_unet_s *u, int n, int d, int hgt, int wid, size_t *bytes);
int sq_tc_forward(sq_unet_s *u, const float *in, int n, int d, int hgt, int wid, float *probs,
                  uint8_t *mask, float *logits, void *ws, size_t ws_bytes, cudaStream_t st);

// 16-bit frames straight into the fused first pair (no float32 copy of the frames): stats == NULL: `frames16` holds
// bf16 values (the normalised frames, sq_image_norm_u16_to_bf16); stats != NULL: raw uint16 camera values that the
// loader normalises itself with the per-frame (mean, std) (SQ_QNORM=2, measured slower)
bool sq_tc_can_take_raw_u16(sq_unet_s *u, int hgt, int wid);
int sq_tc_forward_raw_u16(sq_unet_s *u, const uint16_t *frames16, const float2 *stats, int n, int hgt, int wid,
                          uint8_t *mask, void *ws, size_t ws_bytes, cudaStream_t st);
int sq_image_norm_stats_u16(sq_handle_t h, const uint16_t *in, int n, int hgt, int wid, void *ws, size_t ws_bytes,
                            cudaStream_t st, const float2 **stats_out);
int sq_image_norm_u16_to_bf16(sq_handle_t h, const uint16_t *in, void *out_bf16, int n, int hgt, int wid, void *ws,
                              size_t ws_bytes, cudaStream_t st);

// ---- training step of the fp32 path (train.cu): the forward keeps every activation and says where
struct SqTape {
    std::vector<float *> down, tmp, pooled, up, merged, upt, upo;   // per level, as in fp32_run
    float *logits = nullptr;
    float drop_rate = 0.0f;       // tf.layers.dropout after conv2 of every block (networks/unet.py:274-276)
    unsigned long long seed = 0;
};
int sq_fp32_workspace(sq_unet_s *u, int n, int d, int hgt, int wid, size_t *need);
int sq_fp32_forward_tape(sq_unet_s *u, const float *in, int n, int d, int hgt, int wid, void *ws, size_t ws_bytes,
                         cudaStream_t st, SqTape *tape);

// timer helpers (unet.cu)
void sq_timer_mark(sq_unet_s *u, cudaStream_t st, const char *name, double flops);
