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
  int32_t _pad0;
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
} gpvb200_gemm_desc;

size_t gpvb200_gemm_desc_size(void);
int gpvb200_gemm(const gpvb200_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Hungarian matcher (utils/matcher.py:32-77, utils/box_ops.py:9-59, scipy.optimize.linear_sum_assignment)
 * ------------------------------------------------------------------------------------------------ */
/* Cost blocks C[b][q][t] = w_bbox*L1 + w_class*(-softmax(logits[b,q])[label]) + w_giou*(-GIoU), fp32, in the
 * reference's op order (matcher.py:53-72).  Targets are ragged: image b owns rows tgt_offsets[b]..tgt_offsets[b+1]
 * of tgt_boxes/tgt_labels.  cost is written as B dense blocks [Q][Tmax] (row stride Tmax). */
int gpvb200_matcher_cost(const float* logits /*[B,Q,C]*/, const float* boxes /*[B,Q,4] cxcywh*/,
                         const float* tgt_boxes /*[sumT,4]*/, const int64_t* tgt_labels /*[sumT]*/,
                         const int32_t* tgt_offsets /*[B+1]*/, int32_t B, int32_t Q, int32_t C, int32_t Tmax,
                         float w_class, float w_bbox, float w_giou, float* cost /*[B,Q,Tmax]*/, void* stream);
/* Rectangular linear sum assignment per image, shortest augmenting path in float64 with scipy's visiting
 * order and tie rules (scipy/optimize/rectangular_lsap).  out_q/out_t are [B][min(Q,Tmax)] int64, rows sorted
 * by query index like scipy; entries beyond min(Q,T_b) are -1. */
int gpvb200_lsap(const float* cost /*[B,Q,Tmax]*/, const int32_t* tgt_offsets /*[B+1]*/, int32_t B, int32_t Q,
                 int32_t Tmax, int64_t* out_q, int64_t* out_t, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPVB200_H */
