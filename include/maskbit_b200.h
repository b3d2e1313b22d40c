/* maskbit_b200 -- C ABI of the B200-native MaskBit sampling hot path.
 *
 * The reference (markweberdev/maskbit) has no FFI layer: its seam for this path is three Python callables,
 *   modeling/modules/sampling.py:12-31   sample(model, vqgan_model, ...)
 *   modeling/bert.py:456-508             LFQBert.forward(img_tokens, class_labels, drop_label_mask)
 *   modeling/conv_vqgan.py:98-112        ConvVQModel.decode_tokens(tokens)
 * plus the checkpoint convention modeling/modules/base_model.py:87-142 (load_pretrained of a state_dict).
 * The Python package maskbit_b200 mirrors those callables and binds the entry points below with ctypes
 * (maskbit_b200/_lib.py); INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: every pointer marked "device" is a CUDA device pointer owned by the caller (a torch tensor's
 * data_ptr); the library owns only its packed weights and workspace.  All calls are asynchronous on `stream`
 * unless stated, return 0 on success or a negative mb_status, never throw, and are not thread-safe per handle
 * (one handle per device, one host thread).  mb_last_error() gives the message of the last failure on the
 * calling thread.
 */
#ifndef MASKBIT_B200_H
#define MASKBIT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mb_handle mb_handle;
typedef void* mb_stream; /* cudaStream_t */

enum mb_status {
    MB_OK = 0,
    MB_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
    MB_ERR_CUDA = -2,      /* a CUDA runtime or driver call failed */
    MB_ERR_STATE = -3,     /* call order: weights missing, model not finalized */
    MB_ERR_MISSING = -4,   /* strict loading: a required tensor was not provided */
    MB_ERR_UNEXPECTED = -5 /* strict loading: an unknown tensor name was provided */
};

enum mb_model { MB_GENERATOR = 0, MB_TOKENIZER = 1 };

/* Architecture, from the reference YAML (configs/generator/*.yaml: model.mlm_model.*, model.vq_model.*) and the
 * LFQBert / ConvDecoder constructor arguments (bert.py:345-357, autoencoder.py:358-397). */
typedef struct mb_config {
    int hidden_dim;        /* mlm_model.hidden_dim (1024) */
    int depth;             /* mlm_model.depth (24) */
    int heads;             /* mlm_model.heads (16); head dim must be 64 */
    int mlp_dim;           /* mlm_model.mlp_dim (4096) */
    int token_bits;        /* vq_model.token_size = log2(codebook_size) */
    int codebook_splits;   /* mlm_model.codebook_splits (2) */
    int nclass;            /* 1000; index nclass is the dropped-label embedding (bert.py:371-372) */
    int seq_len;           /* (img_size / input_stride)^2 = 256 */
    int use_prenorm;       /* bert.py use_prenorm: 0 post-norm (every shipped config), 1 pre-norm + norm_after_transformer */
    int dec_hidden_channels;   /* vq_model.hidden_channels (128) */
    int dec_channel_mult[8];   /* vq_model.channel_mult */
    int dec_num_resolutions;   /* vq_model.num_resolutions (5) */
    int dec_num_res_blocks;    /* vq_model.num_res_blocks (2) */
    int num_channels;          /* vq_model.num_channels (3) */
    int generator_cls;         /* mlm_model.model_cls: 0 "lfq_bert" (bert.py:344-508), 1 "bert" (embedding tables, bert.py:184-340) */
    int enc_num_res_blocks;    /* vq_model.num_res_blocks for the ENCODER (autoencoder.py:230-286); dec_num_res_blocks carries the
                                  decoder's num_res_blocks_decoder override (autoencoder.py:371).  0 = same as the decoder */
} mb_config;

int mb_create(const mb_config* cfg, mb_handle** out);
void mb_destroy(mb_handle* h);
const char* mb_last_error(void);
/* ABI / build information: "maskbit_b200 <version> sm_100a" */
const char* mb_version(void);

/* Checkpoint loading -- replaces BaseModel.load_pretrained -> load_state_dict (base_model.py:87-142).
 * Call once per state_dict entry (fp32, contiguous; host or device pointer), then mb_finalize(model), which
 * enforces strict loading (every required name present, with the reference's shapes), repacks into the
 * library's own layouts (bf16 K-major GEMM operands + TMA descriptors, split-bf16 conv weights) and frees the
 * staging copies.  Synchronous. */
int mb_set_tensor(mb_handle* h, int model, const char* name, const float* data, const int64_t* shape, int ndim,
                  int on_device);
int mb_finalize(mb_handle* h, int model);

/* LFQBert.forward (bert.py:456-508).
 *   tokens  device int64 [n_token_rows, seq_len, splits]; sequence i reads row i % n_token_rows
 *   labels  device int64 [n_label_rows]; sequence i reads labels[i % n_label_rows]
 *   drop    device uint8 [n_seq] (1 = replace label by the drop class) or NULL = drop all (drop_label_mask=None)
 *   logits  device fp32 [n_seq, seq_len, splits, V]  (class-token row already removed, bert.py:503) */
/* mb_generator_forward_attn: the same with return_attn=True (bert.py:505-506): additionally writes
 *   attn    device fp32 [depth, n_seq, seq_len + 1, seq_len + 1], layer l = the head-averaged attention weights
 *           nn.MultiheadAttention(need_weights=True) returns in that layer (bert.py:119,137) */
int mb_generator_forward(mb_handle* h, const int64_t* tokens, int n_token_rows, const int64_t* labels, int n_label_rows,
                         const uint8_t* drop, int n_seq, float* logits, mb_stream stream);
int mb_generator_forward_attn(mb_handle* h, const int64_t* tokens, int n_token_rows, const int64_t* labels, int n_label_rows,
                         const uint8_t* drop, int n_seq, float* logits, float* attn, mb_stream stream);

/* One step of the select path (sampling.py:90-131) for B samples.
 *   logits_c / logits_u   device fp32 [B, seq_stride, splits, V]; logits_u NULL = no guidance (sampling.py:100-101)
 *   q       device fp32 [B*n*splits, V] Exp(1) draws, gumbel device fp32 [B, n, splits] raw Gumbel(0,1) draws;
 *           either NULL -> drawn on the device (Philox4x32-10 keyed by seed/step)
 *   tokens_in device int64 [B,n,splits] -> predicted, tokens_out (same shape, distinct buffers) */
typedef struct mb_select_args {
    const float* logits_c;
    const float* logits_u;
    const float* q;
    const float* gumbel;
    const int64_t* tokens_in;
    int64_t* predicted;
    int64_t* tokens_out;
    float scale;                 /* guidance_scale * annealing(i)            (sampling.py:91-98)  */
    float temperature;           /* softmax temperature                      (sampling.py:103-105) */
    float randomize_temperature; /*                                          (sampling.py:117)     */
    float one_minus_progress;    /* 1 - (i+1)/num_steps as fp32              (sampling.py:117)     */
    float mask_len;              /* floor(mask_ratio * n * splits) as fp32   (sampling.py:123)     */
    int B, n, splits, V, seq_stride;
    int64_t mask_token;
    uint64_t seed;
    uint32_t step;
} mb_select_args;
int mb_select_step(mb_handle* h, const mb_select_args* a, mb_stream stream);

/* ConvVQModel.decode_tokens (conv_vqgan.py:98-112): tokens device int64 [B, seq_len] full codebook indices
 * -> images device fp32 [B, 3, H, W] (unclamped, NCHW like the reference). */
int mb_decode_tokens(mb_handle* h, const int64_t* tokens, int B, float* images, mb_stream stream);

/* ConvVQModel.encode (conv_vqgan.py:71-84; ConvEncoder autoencoder.py:230-286 + LookupFreeQuantizer.forward
 * lookup_free.py:46-94): images device fp32 [B,3,H,W] (values as the data pipeline delivers them, [0,1]) ->
 *   z        device fp32 [B, token_bits, P, P] encoder latents before the sign (or NULL)
 *   indices  device int64 [B, P*P] tokens = sum_k [z_k > 0] << k (or NULL)                      (P*P = seq_len) */
int mb_encode(mb_handle* h, const float* images, int B, float* z, int64_t* indices, mb_stream stream);

/* combine_factorized_tokens (factorization.py:7-24): device int64 [B,n,splits] -> device int64 [B,n]. */
int mb_combine_tokens(mb_handle* h, const int64_t* tokens, int B, int64_t* combined, mb_stream stream);

/* clamp(0,1)*255 -> uint8 NHWC (scripts/eval_maskbit.py:134-135): device fp32 [B,3,H,W] -> device u8 [B,H,W,3]. */
int mb_postprocess_u8(mb_handle* h, const float* images, int B, uint8_t* out, mb_stream stream);

/* The whole sampler (sampling.py:57-136) with the step loop resident on the device: no host synchronisation and
 * no host<->device copies between steps.  Per-step host scalars are precomputed by the caller exactly as the
 * reference computes them (tables of length num_steps, host memory).
 *   labels device int64 [B]; images device fp32 [B,3,H,W]; trace device int64 [num_steps,B,n,splits] or NULL
 *   q / gumbel: device fp32 [num_steps, ...] injected noise (parity mode) or NULL (device Philox) */
typedef struct mb_sample_args {
    const int64_t* labels;
    int B;
    int num_steps;
    int use_guidance;             /* 0: guidance_scale == 0 -> single-batch forward (sampling.py:100-101) */
    int skip_zero_scale_uncond;   /* 1: skip the unconditional half on steps whose scale is exactly 0.0 (bit-identical) */
    const float* scale;           /* host [num_steps] */
    const float* temperature;     /* host [num_steps] */
    const float* one_minus_progress; /* host [num_steps] */
    const float* mask_len;        /* host [num_steps] */
    float randomize_temperature;
    const float* q;               /* device [num_steps, B*n*splits, V] or NULL */
    const float* gumbel;          /* device [num_steps, B, n, splits] or NULL */
    uint64_t seed;
    float* images;                /* device fp32 [B,3,H,W] or NULL (skip decode) */
    int64_t* trace;               /* device int64 [num_steps,B,n,splits] or NULL */
    int64_t* final_tokens;        /* device int64 [B,n] combined indices or NULL */
} mb_sample_args;
/* Batches of at most 16 images replay each forward as a CUDA graph on a stream owned by the handle, fenced by events to
 * `stream` at entry and exit (ordering seen from `stream` is unchanged); MASKBIT_B200_GRAPH_MAX_BATCH=0 disables it. */
int mb_sample(mb_handle* h, const mb_sample_args* a, mb_stream stream);

/* ---- forward half of the generator's training step (reference scripts/train_maskbit.py:362-380; no backward) ----
 * Stateless: device pointers in, device pointers out, caller's stream.
 *
 * mb_split_tokens   modeling/modules/factorization.py:27-46: tokens int64 [n] -> out int64 [n, splits],
 *                   out[i, g] = (tokens[i] >> (g * bits_per_split)) & (2^bits_per_split - 1)
 * mb_mask_tokens    modeling/modules/masking.py:7-38 given the random draws: u fp32 [B, slots] uniform [0,1) per slot and
 *                   val_to_mask fp32 [B] per sample (the host mirror maskbit_b200.masking.get_mask_tokens computes it with the
 *                   reference's torch ops and draws u in the reference's order); masked = u < val ? mask_token : token,
 *                   mask uint8 [B, slots] = the predicate
 * mb_mlm_loss       modeling/modules/losses.py:289-339 MLMLoss.forward: logits fp32 [rows, V] (rows = B*n*m), targets int64 [rows],
 *                   masks uint8 [rows] -> out4 fp32 {mlm_loss, correct_tokens, masked_token_loss, masked_correct_tokens};
 *                   scratch: mb_mlm_loss_scratch_bytes() bytes of device memory.  Deterministic (fixed summation order). */
int mb_split_tokens(const int64_t* tokens, int64_t n, int splits, int bits_per_split, int64_t* out, mb_stream stream);
int mb_mask_tokens(const int64_t* tokens, const float* u, const float* val_to_mask, int64_t mask_token, int64_t* masked,
                   uint8_t* mask, int B, int slots, mb_stream stream);
int mb_mlm_loss_scratch_bytes(void);
int mb_mlm_loss(const float* logits, const int64_t* targets, const uint8_t* masks, int64_t rows, int V, int splits,
                float label_smoothing, int sum_splits, void* scratch, float* out4, mb_stream stream);

/* Number of kernels the library has launched on this handle since creation (bench.py "gpu_launches"). */
int64_t mb_launch_count(mb_handle* h);

/* Per-kernel-class timing with CUDA events recorded on the launch stream around every launch of the class
 * (bench.py's live roofline measurement).  mb_profile_enable(h, 1) starts a fresh collection;
 * mb_profile_read synchronises the device, writes the summed milliseconds and launch counts per class into
 * ms[MB_PROF_NUM_KINDS] / counts[MB_PROF_NUM_KINDS] and resets the collection. */
enum mb_prof_kind {
    MB_PROF_EMBED = 0,      /* bit unpack + input_proj + cls/pos + first LayerNorm */
    MB_PROF_GEMM_QKV = 1,   /* [M,1024] x [3072,1024]^T */
    MB_PROF_ATTENTION = 2,
    MB_PROF_GEMM_OUT = 3,   /* [M,1024] x [1024,1024]^T + residual */
    MB_PROF_LAYERNORM = 4,
    MB_PROF_GEMM_UP = 5,    /* [M,1024] x [4096,1024]^T + GELU */
    MB_PROF_GEMM_DOWN = 6,  /* [M,4096] x [1024,4096]^T + residual */
    MB_PROF_GEMM_HEAD = 7,  /* last_layer + prediction_layer */
    MB_PROF_SELECT = 8,
    MB_PROF_DEC_CONV = 9,   /* decoder 3x3 / 1x1 implicit-GEMM convs */
    MB_PROF_DEC_GN = 10,    /* GroupNorm statistics */
    MB_PROF_DEC_IO = 11,    /* conv_in (token unpack) + conv_out */
    MB_PROF_NUM_KINDS = 12
};
int mb_profile_enable(mb_handle* h, int on);
int mb_profile_read(mb_handle* h, double* ms, int64_t* counts, int n_kinds);

/* ---- unit-test hooks on the individual kernels (device pointers, bf16 = uint16 storage) ---- */
/* out = epilogue(A[M,K] W[N,K]^T + bias); epi: 0 bias->bf16, 1 bias+gelu->bf16, 2 bias+residual->f32,
 * 3 bias->f32 with class-row drop (seq_in/seq_out), 4 bias+gelu->f32 */
int mb_test_gemm(const uint16_t* A, const uint16_t* W, const float* bias, const uint16_t* residual, void* out, int M, int N,
                 int K, int epi, int seq_in, int seq_out, mb_stream stream);
/* LayerNorm-folded epilogues (csrc/gemm_tcgen05.cuh): stats_in / stats_out float [M][8][2] partial (sum, sumsq) per row;
 * epi 5: rstd*(acc - mean*vec2) + bias -> bf16; 6: gelu of that -> bf16; 7: acc + bias + ((res - mean)*rstd)*vec2 -> bf16 + stats_out;
 * 8: like 6 + stats_out; 9: like 5 -> f32 with class-row drop.  mean / rstd come from stats_in with width 1/inv_d and eps. */
int mb_test_gemm_ex(const uint16_t* A, const uint16_t* W, const float* bias, const float* vec2, const uint16_t* residual,
                    const float* stats_in, float* stats_out, void* out, int M, int N, int K, int epi, int seq_in, int seq_out,
                    float inv_d, float eps, mb_stream stream);
/* The production-mode (device Philox) noise transforms of the select kernel applied to raw 32-bit draws r[n]:
 * u = uniform in (0,1), q = Exp(1) = -log u, g = Gumbel(0,1) = -log(-log u).  All three finite for every r. */
int mb_test_noise_transform(const uint32_t* r, float* u, float* q, float* g, int n, mb_stream stream);
/* qkv bf16 [n_seq*S, 3*D] -> out bf16 [n_seq*S, D] */
int mb_test_attention(const uint16_t* qkv, uint16_t* out, int n_seq, int S, int D, int H, mb_stream stream);
/* builds with -DATC_TRACE=1 only: device buffer int64 [8 roles][8 events][12 items] receiving block 0's clock64 stamps */
int mb_test_attention_trace(long long* device_buf);
/* GEMM_TRACE builds: device buffer [4 roles][4 events][32 tiles] of clock64 stamps written by the leader CTA of pair 0 */
int mb_test_gemm_trace(long long* device_buf);
#ifdef __cplusplus
}
#endif
#endif /* MASKBIT_B200_H */
