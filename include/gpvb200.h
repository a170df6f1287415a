/*
 * gpvb200 -- C ABI of the B200 (sm_100a) kernels behind the GPV-1 forward/backward hot path.
 *
 * The reference (allenai/gpv-1) has no FFI / operator registry: its seam is the Python module surface of
 * exp/gpv/models/ and utils/matcher.py, utils/set_criterion.py, utils/box_ops.py, every arithmetic step of which is a
 * torch / torchvision / scipy library call.  Each entry point below replaces one family of those library
 * calls (the reference call sites are cited per function) and is bound from Python with ctypes
 * (gpv-1_b200/_C.py).  See INTEGRATION.md for the reference-side stubs.
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch's caching allocator); the library
 *     never allocates, frees or retains them past the call;
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it;
 *   - return 0 on success, a negative code otherwise (see gpvb200_last_error); no exceptions, no exit();
 *   - bf16 tensors are raw uint16 storage (torch.bfloat16), fp32 tensors are float;
 *   - there is no CPU fallback: on a device that is not compute capability 10.x every call fails (-3).
 */
#ifndef GPVB200_H
#define GPVB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPVB200_OK 0
#define GPVB200_ERR_ARG (-1)
#define GPVB200_ERR_CUDA (-2)
#define GPVB200_ERR_ARCH (-3)
#define GPVB200_ERR_WORKSPACE (-4)

int gpvb200_version(void);
/* Copies the last error message of the calling process into buf (NUL-terminated). Returns its length. */
int gpvb200_last_error(char* buf, size_t n);

/* ------------------------------------------------------------------------------------------------
 * Dense contractions on tcgen05 tensor cores (TMA -> shared memory -> tcgen05.mma -> TMEM -> epilogue).
 * One descriptor drives nn.Linear / 1x1 conv / 3x3 conv forward, data-gradient and weight-gradient:
 *   reference call sites: every nn.Linear / nn.Conv2d on the path -- backbone.py:72 (torchvision resnet50),
 *   detr_roi_head.py:79,83-84, transformer.py:153-160,218-231, vilbert.py:748-761,847-851,894-898,
 *   gpv.py:140,145,162,55, answer_head.py:31-33, and their autograd backward.
 *
 * mode 0 (plain)      D[b][m][n] = epi( alpha * sum_k A[b][m][k] * B[b][n][k] )
 *                     A: a_mn=0 -> stored [M][lda] (k contiguous); a_mn=1 -> stored [K][lda] (m contiguous)
 *                     B: b_mn=0 -> stored [N][ldb] (k contiguous); b_mn=1 -> stored [K][ldb] (n contiguous)
 * mode 1 (conv)       A is an NHWC activation [n_img][Hi][Wi][lda]; output pixel (ho,wo) of image i is
 *                     D[i][ho*out_stride+out_off_h][wo*out_stride+out_off_w][n] =
 *                        epi( sum_t sum_c A[i][ho*stride+tap_dh[t]][wo*stride+tap_dw[t]][c] * Bt[tap_w[t]][..] )
 *                     out-of-image taps read zeros.  B: b_mn=0 -> [tap][N][ldb]; b_mn=1 -> [tap][K][ldb].
 * mode 2 (conv wgrad) D[t][m][n] += rowscale[m] * sum_{i,ho,wo} A[i][ho][wo][m] * B[i][ho*stride+dh_t][wo*stride+dw_t][n]
 *                     A = dY NHWC [n_img][Ho][Wo][lda], B = X NHWC [n_img][Hi][Wi][ldb]; D is fp32, atomically
 *                     accumulated (d_atomic must be 1), one [M][ldd] slab per tap (d_batch_stride apart).
 * epilogue            v = alpha*acc; v *= rowscale[m]; v += bias[n]; v += residual[m][n]; D2 = v (optional,
 *                     pre-activation); v = act(v); aux_mode 1: v *= (aux[m][n] > 0); aux_mode 2:
 *                     v *= gelu'(aux[m][n]); store D (bf16 or fp32, plain or atomic add).
 * ------------------------------------------------------------------------------------------------ */
enum { GPVB200_ACT_NONE = 0, GPVB200_ACT_RELU = 1, GPVB200_ACT_GELU = 2, GPVB200_ACT_SIGMOID = 3 };
enum { GPVB200_AUX_NONE = 0, GPVB200_AUX_RELU_MASK = 1, GPVB200_AUX_GELU_GRAD = 2 };

typedef struct gpvb200_gemm_desc {
  int32_t mode;
  int32_t M, N, K;
  int32_t batch;          /* mode 0: number of independent problems (grid y) */
  int32_t a_mn, b_mn;
  int32_t act;
  int32_t aux_mode;
  int32_t d_fp32;         /* D element type: 0 bf16, 1 fp32 */
  int32_t d_atomic;       /* 1: fp32 atomic accumulate into D (required for splits > 1) */
  int32_t splits;         /* split the contraction over this many CTAs (0/1 = none) */
  /* conv geometry (modes 1, 2) */
  int32_t n_img, Hi, Wi, Ho, Wo, stride;
  int32_t ntaps;
  int32_t tap_dh[9], tap_dw[9], tap_w[9];
  int32_t OH, OW, out_stride, out_off_h, out_off_w;
  float alpha;
  int32_t res_fp32;       /* residual element type: 0 bf16, 1 fp32 */
  const void* A;
  const void* B;
  void* D;
  void* D2;               /* optional bf16 pre-activation copy, same indexing as D */
  const float* bias;      /* [N] fp32 or NULL */
  const float* rowscale;  /* [M] fp32 or NULL */
  const void* residual;   /* bf16, indexed like D with leading dimension ldr, or NULL */
  const void* aux;        /* bf16, indexed like D with leading dimension ldaux, or NULL */
  int64_t lda, ldb, ldd, ldr, ldaux;
  int64_t a_batch_stride, b_batch_stride, d_batch_stride; /* elements; mode 0 batches / mode 2 taps */
  /* train-mode dropout fused into the epilogue (nn.Dropout sites of transformer.py:155,158-160, vilbert.py:848-851,
   * 471,514 and nn.TransformerDecoderLayer): drop_mode 1 = on alpha*acc*rowscale + bias, before the residual is added
   * (y = res + dropout(xW + b)); 2 = on the activation output (h = dropout(relu(xW + b))).  The mask of element
   * (row, col) is a counter-based function of (*drop_seed, drop_site, row, col), see gpvb200_dropout_mask. */
  const void* drop_seed;  /* device uint64 scalar (training-step counter) or NULL */
  int32_t drop_mode;      /* 0 none */
  uint32_t drop_site;
  float drop_p;
  int32_t max_ctas;       /* > 0: the persistent grid uses at most this many CTAs (a launch that runs on a lane beside a dependent chain
                           * of kernels leaves the other SMs to the chain); 0: one CTA per SM */
} gpvb200_gemm_desc;

size_t gpvb200_gemm_desc_size(void);
int gpvb200_gemm(const gpvb200_gemm_desc* d, void* stream);
/* Number of gpvb200_gemm calls of this process that ran the CTA-pair variant (`cta_group::2`: two SMs share the B tile;
 * chosen for deep contractions with K-major A and N > 128; GPVB200_PAIR=<min 64-deep k-blocks>, 0 = off). */
int64_t gpvb200_gemm_pair_launches(void);

/* ------------------------------------------------------------------------------------------------
 * Hungarian matcher (utils/matcher.py:32-77, utils/box_ops.py:9-59, scipy.optimize.linear_sum_assignment)
 * ------------------------------------------------------------------------------------------------ */
/* Cost blocks C[b][q][t] = w_bbox*L1 + w_class*(-softmax(logits[b,q])[label]) + w_giou*(-GIoU), fp32, in the
 * reference's op order (matcher.py:53-72).  Targets are ragged: image b owns rows tgt_offsets[b]..tgt_offsets[b+1]
 * of tgt_boxes/tgt_labels.  cost is written as B dense blocks [Q][Tmax] (row stride Tmax). */
int gpvb200_matcher_cost(const float* logits /*[B*Q] rows of C, row stride ldl*/, int64_t ldl,
                         const float* boxes /*[B*Q] rows of 4 (cxcywh), row stride ldb*/, int64_t ldb, const float* tgt_boxes /*[sumT,4]*/, const int64_t* tgt_labels /*[sumT]*/,
                         const int32_t* tgt_offsets /*[B+1]*/, int32_t B, int32_t Q, int32_t C, int32_t Tmax,
                         float w_class, float w_bbox, float w_giou, float* cost /*[B,Q,Tmax]*/, void* stream);
/* Rectangular linear sum assignment per image, shortest augmenting path in float64 with scipy's visiting
 * order and tie rules (scipy/optimize/rectangular_lsap).  out_q/out_t are [B][min(Q,Tmax)] int64, rows sorted
 * by query index like scipy; entries beyond min(Q,T_b) are -1. */
int gpvb200_lsap(const float* cost /*[B,Q,Tmax]*/, const int32_t* tgt_offsets /*[B+1]*/, int32_t B, int32_t Q,
                 int32_t Tmax, int64_t* out_q, int64_t* out_t, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused attention core: O = softmax(scale * Q K^T + mask) V, one (batch, head) per CTA, scores stay on chip.
 * Replaces nn.MultiheadAttention's bmm/softmax/bmm (transformer.py:153-155, 218-226; gpv.py:38-43 via
 * nn.TransformerDecoderLayer) and BertBiAttention's matmul/softmax/matmul (vilbert.py:766-815).
 * Tensors are token-major: element (b, s, h, d) of X lives at X[(b*S + s)*ldx + h*dh + d] (bf16), so the packed
 * QKV projection output can be read in place.  key_mask [B,Sk] (1 = ignore key) may be NULL; causal masks keys j > i.
 * lse [B,H,Sq] fp32 (log2 domain) is written by fwd (may be NULL for inference) and consumed by bwd.
 * ------------------------------------------------------------------------------------------------ */
int gpvb200_attention_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const uint8_t* key_mask,
                          int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int32_t B, int32_t H, int32_t Sq,
                          int32_t Sk, int32_t dh, int32_t causal, float scale, void* stream);
/* same with explicit batch strides (elements; 0 = S*ld): lets a decode step read a preallocated KV cache
 * [B][S_max][H*dh] in place with Sk = tokens written so far (the KV-cached text decoder, gpv.py:178-196, 256-328) */
int gpvb200_attention_fwd_bs(const void* q, const void* k, const void* v, void* o, float* lse, const uint8_t* key_mask,
                             int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t bsq, int64_t bsk, int64_t bsv,
                             int64_t bso, int32_t B, int32_t H, int32_t Sq, int32_t Sk, int32_t dh, int32_t causal, float scale,
                             void* stream);
/* train-mode variants: dropout with probability drop_p on the attention probabilities (nn.MultiheadAttention dropout,
 * vilbert.py:443,782,805); mask row = (b*H + h)*Sq + q, col = key, site drop_site, seed *drop_seed */
int gpvb200_attention_fwd_drop(const void* q, const void* k, const void* v, void* o, float* lse, const uint8_t* key_mask,
                               int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int32_t B, int32_t H, int32_t Sq, int32_t Sk,
                               int32_t dh, int32_t causal, float scale, const void* drop_seed, uint32_t drop_site, float drop_p,
                               void* stream);
int gpvb200_attention_bwd_drop(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                               const uint8_t* key_mask, void* dq, void* dk, void* dv, int64_t ldq, int64_t ldk, int64_t ldv,
                               int64_t ldo, int64_t lddo, int64_t lddq, int64_t lddk, int64_t lddv, int32_t B, int32_t H,
                               int32_t Sq, int32_t Sk, int32_t dh, int32_t causal, float scale, const void* drop_seed,
                               uint32_t drop_site, float drop_p, void* stream);
int gpvb200_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                          const uint8_t* key_mask, void* dq, void* dk, void* dv, int64_t ldq, int64_t ldk, int64_t ldv,
                          int64_t ldo, int64_t lddo, int64_t lddq, int64_t lddk, int64_t lddv, int32_t B, int32_t H,
                          int32_t Sq, int32_t Sk, int32_t dh, int32_t causal, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the last dimension (nn.LayerNorm eps 1e-5: transformer.py:139-140,199-201, gpv.py:38-43;
 * BertLayerNorm eps 1e-12: vilbert.py:296-316; F.layer_norm without affine: detr_roi_head.py:91).
 * x, y, dy, dx bf16 with row strides; gamma/beta fp32 or both NULL; stats [M][2] = (mean, rstd) fp32.
 * bwd accumulates dgamma/dbeta atomically (fp32).
 * ------------------------------------------------------------------------------------------------ */
int gpvb200_layernorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* y, int64_t ldy,
                          float* stats, int32_t M, int32_t D, void* stream);
int gpvb200_layernorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const float* stats, const float* gamma,
                          void* dx, int64_t lddx, float* dgamma, float* dbeta, int32_t M, int32_t D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row-tile-resident transformer sub-layers (layer_umma.cu): one CTA keeps a 128-token tile of the d_model = 256
 * stream in shared memory across the whole sub-layer (tcgen05.mma + TMA; bias / ReLU / dropout / residual /
 * LayerNorm in the epilogues).  SURVEY 8b: `mlp_block`, `attn_block`.
 * ------------------------------------------------------------------------------------------------ */
/* y = LayerNorm(x + drop_o(W2 drop_h(relu(W1 x + b1)) + b2)): TransformerEncoderLayer.forward_post transformer.py:157-160,
 * TransformerDecoderLayer.forward_post transformer.py:228-231 (linear1 -> ReLU -> dropout -> linear2 -> dropout ->
 * residual -> norm).  x [M, 256] bf16, w1 [d_ff, 256] bf16, w2 [256, d_ff] bf16 (nn.Linear layouts), b1 / b2 / gamma / beta fp32.
 * Outputs: y [M, 256] bf16; optional (NULL to skip) h [M, d_ff] bf16 = the hidden activation after ReLU / dropout,
 * pre [M, 256] bf16 = the pre-norm sum, stats [M, 2] fp32 = (mean, rstd) -- what the backward pass reads.
 * seq_len > 0: rows are tiled per sequence of seq_len rows (128-row tiles never straddle two sequences); 0: flat tiles.
 * drop_seed: device uint64 step counter or NULL (eval); site_h / p_h: hidden dropout; site_o / p_o: output dropout. */
int gpvb200_mlp_block_fwd(const void* x, int64_t ldx, const void* w1, int64_t ldw1, const float* b1, const void* w2, int64_t ldw2,
                          const float* b2, const float* gamma, const float* beta, float eps, void* y, int64_t ldy, void* h,
                          int64_t ldh, void* pre, int64_t ldpre, float* stats, int64_t M, int32_t d_model, int32_t d_ff,
                          int32_t seq_len, const void* drop_seed, uint32_t site_h, float p_h, uint32_t site_o, float p_o,
                          void* stream);

/* Data gradients of the same sub-layer in one launch: dh = alpha * (dy W2) (*) [h > 0];  dx = dh W1 + dres  (autograd of
 * transformer.py:157-160 between the LayerNorm backward and the weight gradients).  dy [M, 256] bf16 (row stride lddy), w2t = W2^T
 * [d_ff, 256] and w1t = W1^T [256, d_ff] bf16 (the transposed weights, so both contractions read K-major operands), h [M, d_ff] the
 * forward's hidden activation (post ReLU / dropout: its sign is relu' and the dropout mask at once; alpha = 1 / (1 - p_hidden)), dres
 * [M, 256] the gradient of the residual branch.  Writes dh [M, d_ff] (for dW2 = dy^T h, dW1 = dh^T x) and dx [M, 256]. */
int gpvb200_mlp_block_bwd(const void* dy, int64_t lddy, const void* w2t, int64_t ldw2t, const void* w1t, int64_t ldw1t, const void* h,
                          int64_t ldh, float alpha, const void* dres, int64_t lddres, void* dh, int64_t lddh, void* dx, int64_t lddx,
                          int64_t M, int32_t d_model, int32_t d_ff, int32_t seq_len, void* stream);

/* Multi-head attention for d_model = 256 (8 heads x 32) on tcgen05, optionally with the output projection, residual and LayerNorm
 * in the same kernel: o = concat_h softmax(scale Q_h K_h^T + key mask) V_h;  y = LayerNorm(x + drop_o(o Wo^T + bo)).
 * nn.MultiheadAttention core + out_proj + dropout + residual + norm of TransformerEncoderLayer.forward_post transformer.py:153-157
 * and TransformerDecoderLayer.forward_post transformer.py:216-227.  q [B*Sq, >=256], k / v [B*Sk, >=256] bf16 token-major (slices
 * of a packed projection buffer: row strides ldq / ldk / ldv), Sk <= 304; key_mask [B, Sk] uint8 (1 = padded key) or NULL.
 * o [B*Sq, 256] bf16 and lse [B, H, Sq] fp32 (log2 domain) are what gpvb200_attention_bwd reads (either may be NULL).
 * wo = NULL: attention core only.  Otherwise wo [256, 256] bf16, bo / gamma / beta fp32, x = residual [B*Sq, 256] bf16, outputs
 * y (required), pre, stats (NULL to skip).  site_p / p_p: dropout on the probabilities (same mask as gpvb200_attention_fwd_drop),
 * site_o / p_o: dropout on the projected output before the residual. */
int gpvb200_attn_block_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const uint8_t* key_mask,
                           int32_t B, int32_t H, int32_t Sq, int32_t Sk, int32_t dh, float scale, void* o, int64_t ldo, float* lse,
                           const void* wo, int64_t ldwo, const float* bo, const void* x, int64_t ldx, const float* gamma,
                           const float* beta, float eps, void* pre, int64_t ldpre, void* y, int64_t ldy, float* stats,
                           const void* drop_seed, uint32_t site_p, float p_p, uint32_t site_o, float p_o, void* stream);

/* Developer hook: a device buffer of int64 that CTA 0 of the layer kernels fills with clock64() stamps of its producer /
 * MMA / epilogue roles ([3][chunks + 1][8]); NULL switches it off (the default).  tools/trace_layer.py. */
int gpvb200_layer_trace(void* buf);

/* ------------------------------------------------------------------------------------------------
 * HBM-bound helpers (elementwise.cu)
 * ------------------------------------------------------------------------------------------------ */
/* out[m] = x[m] + p[m % P]  (x may be NULL): q = k = src + pos, transformer.py:153,218,223-224 */
int gpvb200_add_rowbcast(const void* x, int64_t ldx, const void* p, int64_t ldp, void* out, int64_t ldo, int64_t M, int32_t D,
                         int32_t P, void* stream);
/* out[n] += sum_m dy[m][n]  (bias gradients of every nn.Linear / conv bias) */
int gpvb200_colsum(const void* dy, int64_t ld, float* out, int64_t M, int32_t N, void* stream);
/* out[s][d] += sum_b x[b*S+s][d]  (gradient of batch-broadcast parameters, e.g. detr.query_embed) */
int gpvb200_batch_reduce(const void* x, int64_t ld, float* out, int32_t B, int32_t S, int32_t D, void* stream);
/* multi-tensor fp32 -> bf16 weight packing with FrozenBN fold (backbone.py:44-54); items is a device array of
 * {const float* src; bf16* dst; const float* scale; int32 O, I, taps, mode} */
size_t gpvb200_pack_item_size(void);
int gpvb200_pack_chunk(void);
int gpvb200_pack_weights(const void* items, const int32_t* blk_item, const int32_t* blk_chunk, int32_t n_blocks, void* stream);
int gpvb200_bn_fold(const float* w, const float* b, const float* rm, const float* rv, float* scale, float* bias, int32_t n,
                    void* stream);
/* torchvision resnet stem pieces (backbone.py:72): 3x3/s2 max-pool on NHWC bf16; 7x7/s2 im2col from NCHW fp32 */
int gpvb200_maxpool3x3s2(const void* x, void* y, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
int gpvb200_stem_im2col(const float* img, void* col, int32_t B, int32_t H, int32_t W, void* stream);
/* stem space-to-depth: NCHW fp32 -> zero-bordered bf16 [B][Ho+4][Wo+4][16] (Ho = ceil(H/2)), channel = dy*6 + dx*3 + c;
 * with it the 7x7/s2 stem (backbone.py:72) is a 4-tap K=64 implicit GEMM of gpvb200_gemm mode 1 (pixel stride 16) */
int gpvb200_stem_s2d(const float* img, void* out, int32_t B, int32_t H, int32_t W, void* stream);
/* same map from uint8 NHWC [B][H][W][3] with x = (u8/255 - mean[c]) / std[c] folded in (ToTensor + Normalize of
 * datasets/coco_generic_dataset.py:31-32); mean3 / std3 are HOST arrays of three floats */
int gpvb200_stem_s2d_u8(const uint8_t* img, void* out, int32_t B, int32_t H, int32_t W, const float* mean3, const float* std3,
                        void* stream);
/* ROI-align(7x7, aligned, adaptive sampling)+mean as separable weights (detr_roi_head.py:44-56): wroi[bq][y*W+x] */
int gpvb200_roi_weights(const float* boxes, int64_t ldb, void* wroi, int64_t ldw, int32_t BQ, int32_t H, int32_t W, void* stream);
/* relevance conditioning gpv.py:364-375 (+ the memory concat gpv.py:175 through the output row remap) */
int gpvb200_relevance_mix_fwd(const void* x, int64_t ldx, const float* logits, int64_t ldl, const float* tok, void* out,
                              int64_t ldo, int32_t M, int32_t D, int32_t G, int32_t out_gstride, int32_t out_off, void* stream);
int gpvb200_relevance_mix_bwd(const void* dy, int64_t lddy, const float* logits, int64_t ldl, const float* tok, float* dlogits,
                              int64_t lddl, float* dtok, int32_t M, int32_t D, int32_t G, int32_t gstride, int32_t off, void* stream);
/* out[m] = table[ids[m]] (+ pos[m % T]) (+ cst): AnswerInputEmbedding gpv.py:53, BERT embeddings */
int gpvb200_gather_rows(const float* table, const int64_t* ids, const float* pos, const float* cst, void* out, int64_t ldo,
                        int64_t M, int32_t D, int32_t T, void* stream);
/* Same, and also writes pad_mask[m] = (ids[m] == pad_id): the key-padding mask HF's BertTokenizer(padding=True) hands to
 * BertModel as attention_mask (bert.py:12-21; [PAD] = 0), produced by the embedding gather that reads the ids anyway. */
int gpvb200_gather_rows_mask(const float* table, const int64_t* ids, const float* pos, const float* cst, void* out, int64_t ldo,
                             int64_t M, int32_t D, int32_t T, uint8_t* pad_mask, int64_t pad_id, void* stream);
/* row-remapped bf16 copy: row(m) = (m / G) * gstride + off + m % G on either side */
int gpvb200_copy_rows(const void* src, int64_t lds, int32_t sG, int32_t sgs, int32_t soff, void* dst, int64_t ldd, int32_t dG,
                      int32_t dgs, int32_t doff, int64_t M, int32_t D, void* stream);
int gpvb200_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);
/* conv weight gradient: packed fp32 [taps][O][I] -> Conv2d master layout [O][I][taps]; dst = src or dst += src */
int gpvb200_unpack_conv_grad(const float* src, float* dst, int32_t O, int32_t I, int32_t taps, int32_t accumulate, void* stream);
/* out = a + b on bf16 rows (gradient joins of residual branches) */
int gpvb200_add_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo, int64_t M, int32_t D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Criterion (losses.py:20-26,155-176; utils/set_criterion.py:44-62,78-97)
 * ------------------------------------------------------------------------------------------------ */
/* loss_sum += sum_rows w[row] * CE(logits[row], targets[row]);  dlogits = w[row] * (softmax - onehot) as bf16 */
int gpvb200_ce_fwd_bwd(const float* logits, int64_t ldl, const int64_t* targets, const float* row_weight, float* loss_sum,
                       float* row_loss, void* dlogits, int64_t ldd, int32_t rows, int32_t V, void* stream);
/* out3 += (loss_ce, loss_bbox, loss_giou); dlogits [B*Q][ldl] fp32 and dbox_pre [B*Q][lddb] bf16 (gradient w.r.t.
 * the pre-sigmoid box head output), both already multiplied by the loss weights wt_*.  weight_sum = sum of the
 * class weights over all (valid image, query) pairs and num_boxes = max(sum T_b, 1); pass both <= 0 to have the
 * kernel derive them from tgt_offsets / loc_valid on the device (CUDA-graph replay with changing targets). */
int gpvb200_set_criterion(const float* logits, int64_t ldl, const float* boxes, int64_t ldb, const float* tgt_boxes,
                          const int32_t* tgt_offsets, const int64_t* idx_q, const int64_t* idx_t, int32_t Kmax,
                          const uint8_t* loc_valid, int32_t B, int32_t Q, float eos_coef, float weight_sum, float num_boxes,
                          float wt_ce, float wt_bbox, float wt_giou, float* out3, float* dlogits, void* dbox_pre, int64_t lddb,
                          void* stream);

/* ---- train-mode dropout helpers.  layernorm_fwd_drop: y = dropout(LN(x)) (BERT embeddings, vilbert.py:364 / HF
 * BertEmbeddings).  layernorm_bwd_drop: additionally writes dx_masked = dx (*) mask / (1 - p), the gradient of the
 * dropped-out sub-layer output in y = LN(res + dropout(f)).  dropout_mask: out[row*N + col] = 1 kept / 0 dropped for
 * the logical [rows, N] tensor of (seed, site) -- what the tests feed to autograd to check forward and backward. */
int gpvb200_layernorm_fwd_drop(const void* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* y, int64_t ldy,
                               float* stats, int32_t M, int32_t D, const void* drop_seed, uint32_t drop_site, float drop_p,
                               void* stream);
int gpvb200_layernorm_bwd_drop(const void* dy, int64_t lddy, const void* x, int64_t ldx, const float* stats, const float* gamma,
                               void* dx, int64_t lddx, float* dgamma, float* dbeta, int32_t M, int32_t D, void* dx_masked,
                               int64_t lddxm, const void* drop_seed, uint32_t drop_site, float drop_p, void* stream);
int gpvb200_dropout_mask(uint8_t* out, int64_t rows, int32_t N, const void* drop_seed, uint32_t drop_site, float drop_p,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Decode-loop bookkeeping (exp/gpv/models/gpv.py:178-196 greedy, 256-362 beam search)
 * ------------------------------------------------------------------------------------------------ */
/* Greedy step, gpv.py:186-190: out[row, :V] = logits[row, :V] + vocab_mask (NULL: no mask; out NULL: not written) and
 * ids[row] = arg-max of it, the FIRST maximal index on ties (torch.max semantics).  logits fp32 [rows, >= V]. */
int gpvb200_argmax(const float* logits, int64_t ld, int32_t rows, int32_t V, const float* vocab_mask, float* out, int64_t ldo,
                   int64_t* ids, void* stream);
/* One beam-search step, gpv.py:280-328: logits fp32 [B*K, >= V] are the next-token logits of hypothesis (b, k1).  Per
 * hypothesis log_softmax and its K best tokens, candidates score_in[b, k1] + logp in (k1 major, k2 minor) order, the K
 * best candidates in stable descending order (first occurrence wins ties); at t = 0 only hypothesis 0 of each image is
 * live.  Writes score_out [B, K], ids_out [B, K, L] (ids_out[b, k, :t+1] = ids_in[b, k1, :t+1], ids_out[b, k, t+1] =
 * the new token), parent [B*K] = b K + k1 (the row whose KV cache hypothesis (b, k) continues) and tok [B*K] = the new
 * token.  K <= 8; ids / scores are double-buffered by the caller (no in-place update); workspace: >= 8 B K K bytes of device memory,
 * 8-byte aligned (candidate scores and tokens between the per-row and the per-image launch). */
int gpvb200_beam_update(const float* logits, int64_t ld, int32_t B, int32_t K, int32_t V, int32_t t, int32_t L, const float* score_in,
                        const int64_t* ids_in, float* score_out, int64_t* ids_out, int64_t* parent, int64_t* tok, void* workspace,
                        void* stream);
/* Attention of ONE query row per hypothesis against cached keys / values: the self- and cross-attention cores of a KV-cached decode step
 * (nn.TransformerDecoderLayer inside GPV.decode_text, gpv.py:449-466, called with a one-token query).  q [Bq, >= H*dh] bf16 (row stride
 * ldq); hypothesis r reads the K / V batch r / rep (rep hypotheses share one batch: the beams of an image attend to the same encoder
 * memory; rep = 1 for per-hypothesis caches): key s of batch b at k + b*bsk + s*ldk, Sk keys; o [Bq, H*dh] bf16.  K and V of a
 * (batch, head) are staged on chip once and serve its rep hypotheses. */
int gpvb200_decode_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, int64_t bsk, const void* v, int64_t ldv, int64_t bsv,
                             void* o, int64_t ldo, int32_t Bq, int32_t rep, int32_t H, int32_t Sk, int32_t dh, float scale, void* stream);
/* KV-cache permutation after a beam step (gpv.py:318-326 re-decodes the re-ordered prefixes; the cached K/V follow the
 * hypotheses instead): dst[r, :n_elems] = src[parent[r], :n_elems] for `rows` rows of row_elems bf16. */
int gpvb200_reorder_rows(const void* src, void* dst, const int64_t* parent, int32_t rows, int64_t row_elems, int64_t n_elems,
                         void* stream);

/* ---- optimizer: clip_grad_norm_ + AdamW over the flat gradient arena (exp/gpv/train_distr.py:414-428, 228-253).
 * items: device array of {float* p; int64 goff; int32 n, group, clip, step0} (gpvb200_optim_item_size() bytes each);
 * blk_item / blk_chunk: one entry per CTA = (tensor index, chunk of gpvb200_optim_chunk() elements).
 * grad_sqnorm: out_sq[0] = sum of squares over the listed chunks.  clip_adamw: gradients of items with clip = 1 are
 * scaled by min(1, max_norm / (sqrt(total_sq[0]) + 1e-6)) in place, then p, m, v follow torch.optim.AdamW at step `step`
 * (>= 1) with the learning rate of the item's group; a tensor's own Adam step (bias correction) is step - step0, as torch keeps
 * state['step'] per parameter. */
size_t gpvb200_optim_item_size(void);
int gpvb200_optim_chunk(void);
int gpvb200_grad_sqnorm(const void* items, const int32_t* blk_item, const int32_t* blk_chunk, int32_t n_blocks, const float* grads,
                        float* out_sq, void* stream);
int gpvb200_clip_adamw(const void* items, const int32_t* blk_item, const int32_t* blk_chunk, int32_t n_blocks, float* grads, float* m,
                       float* v, const float* total_sq, float max_norm, float lr0, float lr1, float lr2, float lr3, float beta1,
                       float beta2, float eps, float weight_decay, int64_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPVB200_H */
