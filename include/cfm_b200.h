/*
 * cfm_b200.h -- C-ABI of the B200-native (sm_100a) Conformer-encoder kernels.
 *
 * The reference (Lingeng56/conformer-pytorch-lightning) is pure Python/PyTorch and has
 * no FFI of its own: its hot path is a chain of ATen calls made from
 * src/encoder_layer.py:49-71.  This header is therefore the boundary *we* define; each
 * entry point names the reference statements it replaces.  The host side that mirrors
 * the reference's nn.Module surface (the .py files of conformer_pytorch_lightning_b200) binds these
 * with ctypes -- see INTEGRATION.md for the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; no torch types;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - no hidden allocations, no hidden synchronisation: callers own all buffers;
 *   - return value 0 = success, negative = error; cfm_last_error() gives the text of the
 *     last failure on the calling thread.  Nothing throws across the ABI;
 *   - "act" tensors are row-major (rows = B*T tokens, cols = channels) in `dtype`
 *     CFM_F32 or CFM_BF16; the residual stream, biases, LayerNorm/BatchNorm parameters and
 *     all accumulation are always fp32.
 */
#ifndef CFM_B200_H_
#define CFM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFM_ABI_VERSION 1

/* activation dtypes */
#define CFM_F32  0
#define CFM_BF16 1

/* GEMM epilogues (cfm_gemm) */
#define CFM_EPI_BIAS      0  /* C[act]  = A W^T + b                        nn.Linear (attention.py:62-64,78,99)      */
#define CFM_EPI_BIAS_SILU 1  /* C[act]  = silu(A W^T + b)                  w_1 + SiLU (feedforward.py:17-18)          */
#define CFM_EPI_BIAS_GLU  2  /* C[act]  = (A Wa^T + ba) * sigmoid(A Wb^T + bb), W = [Wa;Wb]
                                                                           pointwise_conv1 + GLU (convolution.py:41-42) */
#define CFM_EPI_RESIDUAL  3  /* X[f32]  = R + alpha * rowmask(A W^T + b)   (R may be null: plain fp32 output)
                                                                           w_2 / linear_out / pointwise_conv2 + the
                                                                           residual adds of encoder_layer.py:58,62,66,69
                                                                           and the masked_fill of convolution.py:47-48 */

#define CFM_EPI_BIAS_RELU 4  /* C[act]  = relu(A W^T + b)                  w_1 + ReLU (feedforward.py:10-11, activation='relu') */

/* GEMM engines */
#define CFM_ENGINE_AUTO 0    /* tcgen05 when dtype==BF16 and the shape is supported, else SIMT */
#define CFM_ENGINE_SIMT 1    /* fp32-accumulating CUDA-core kernel (exact-order reference engine) */
#define CFM_ENGINE_TC   2    /* tcgen05/TMEM/TMA kernel; error if unsupported */

int         cfm_abi_version(void);
const char* cfm_last_error(void);
/* one-time per-process/device initialisation (function attributes, driver entry points). Host only. */
int         cfm_init(int device);
/* number of kernel launches issued through this library by the calling process (for bench `gpu_launches`) */
int64_t     cfm_launch_count(void);
/* launches of one kernel family by name ("ffn_fused", "mhsa_fused", "conv_fused", "gemm_tc", "attention_tc", ...):
 * lets a caller (and the parity tests) verify which engine served its calls; 0 for unknown names */
int64_t     cfm_kernel_launches(const char* name);

/*
 * LayerNorm (+ optional second LayerNorm, + optional row mask).  Replaces nn.LayerNorm calls
 * encoder_layer.py:56,59,63,67,70 and encoder.py:74, and the first masked_fill of
 * convolution.py:36-37 when `row_valid` is given.
 *   t = LN(x; g1,b1);  if x_out: x_out = t (fp32);  if g2: t = LN(t; g2,b2);
 *   if row_valid: t = row_valid[row] ? t : 0;  if y: y = cast<y_dtype>(t)
 * x, x_out: (rows, d) fp32.  y: (rows, d) in y_dtype.  d % 128 == 0, d <= 1024.
 */
int cfm_layernorm(const float* x, int rows, int d,
                  const float* g1, const float* b1, float* x_out,
                  const float* g2, const float* b2,
                  void* y, int y_dtype, const uint8_t* row_valid, float eps, void* stream);

/*
 * C = epilogue(A W^T + bias).  A: (M,K) act dtype, row stride lda elements.  W: (N,K) act dtype
 * row-major (nn.Linear / 1x1-conv weight layout), for CFM_EPI_BIAS_GLU W is (2N,K) = [Wa;Wb]
 * and bias has 2N entries.  bias fp32 or NULL.
 *   EPI_BIAS / EPI_BIAS_SILU / EPI_BIAS_GLU : C (M,N) act dtype, row stride ldc.
 *   EPI_RESIDUAL : C (M,N) fp32 = residual (M,N) fp32 (may alias C) + alpha * v, where
 *                  v = (A W^T + bias), zeroed on rows with row_valid[row]==0 when row_valid != NULL.
 */
int cfm_gemm(const void* A, int lda, const void* W, const float* bias,
             void* C, int ldc, int M, int N, int K, int dtype, int epilogue,
             const float* residual, float alpha, const uint8_t* row_valid,
             int engine, void* stream);

/*
 * Residual GEMM with the following LayerNorm(s) fused into its epilogue (in place on the residual stream X):
 *   v = X + alpha * rowmask(A W^T + bias)                      (rowmask: rows with row_valid[row]==0 contribute 0)
 *   g2 == NULL :  X = v            ; Y = ymask(LN(v; g1,b1))   (encoder_layer.py:58-59, 62-63, 66-67)
 *   g2 != NULL :  X = LN(v; g1,b1) ; Y = ymask(LN(X; g2,b2))   (norm_final chained with the next layer's
 *                                                               norm_ff_macaron, encoder_layer.py:69-70,56)
 * X: (M,N) fp32 row stride ldx.  Y: (M,N) act dtype row stride ldy, rows with y_row_valid[row]==0 are zeroed
 * (the masked_fill of convolution.py:36-37).  On the tcgen05 engine (bf16, N == 256) this is ONE kernel: the
 * row statistics are thread-local in the epilogue; otherwise the library runs cfm_gemm + cfm_layernorm.
 */
int cfm_gemm_ln(const void* A, int lda, const void* W, const float* bias, float* X, int ldx,
                int M, int N, int K, int dtype, float alpha, const uint8_t* row_valid,
                const float* g1, const float* b1, const float* g2, const float* b2,
                void* Y, int ldy, const uint8_t* y_row_valid, float eps, int engine, void* stream);

/*
 * Whole macaron feed-forward on the residual stream, in place (feedforward.py:16-21 + encoder_layer.py:56-59,67-70):
 *   X += alpha * (W2 silu(W1 y + b1) + b2), followed by the optional LayerNorm(s) exactly as in cfm_gemm_ln
 *   (g1 == NULL: none; g2 == NULL: Y = ymask(LN(X;g1,be1)); else X = LN(.;g1,be1), Y = ymask(LN(X;g2,be2))).
 * y: (M,d) act dtype (row stride ld_in), W1: (F,d), W2: (d,F) act dtype, b1 (F), b2 (d) fp32, X: (M,d) fp32.
 * Y may alias y.  On the tcgen05 engine (bf16, d == 256, F % 128 == 0) this is ONE kernel and the (M,F) hidden
 * activation stays in TMEM / shared memory; otherwise the library runs cfm_gemm (SiLU) + cfm_gemm / cfm_gemm_ln
 * through `hidden_ws`, a caller-provided (M,F) act-dtype scratch buffer (may be NULL only if the fused path applies).
 */
int cfm_ffn(const void* y, int ld_in, const void* W1, const float* b1, const void* W2, const float* b2,
            float* X, int ldx, int M, int d, int F, int dtype, float alpha,
            const float* g1, const float* be1, const float* g2, const float* be2,
            void* Y, int ld_out, const uint8_t* y_row_valid, float eps, void* hidden_ws, int engine, void* stream);

/*
 * Up to two feed-forward modules applied back to back to the same rows, optionally followed by a projection of the
 * final LayerNorm output, i.e. exactly
 *   [cfm_ffn(y, a..., X, g1a/be1a/g2a/be2a -> Y)]           (skipped when W1a == NULL)
 *   cfm_ffn(Y or y, b..., X, g1b/be1b/g2b/be2b -> Y, y_row_valid)
 *   [cfm_gemm(Y, Wp, bp -> P (M, Np), CFM_EPI_BIAS)]         (skipped when Wp == NULL)
 * This is the layer boundary of the encoder: the second half-step feed-forward of layer i with norm_final +
 * norm_ff_macaron of layer i+1 (encoder_layer.py:68-70, :56), the first feed-forward of layer i+1 with its attention
 * LayerNorm (:57-59) and that attention's Q/K/V projections (attention.py:62-64, Wp = [Wq;Wk;Wv]).
 * Module a must have a LayerNorm (g1a != NULL) since its output feeds module b.  y, Y: (M,d) act dtype contiguous
 * (Y may alias y), X: (M,d) fp32 contiguous, both modules with the same F.
 * On the tcgen05 engine (bf16, d == 256, F % 256 == 0, alpha_b a power of two, Np % 256 == 0 with g2b == NULL) this is
 * ONE kernel: the residual stream and the LayerNorm output between the two modules stay in TMEM / shared memory, and
 * with a projection the final LayerNorm output is consumed on chip and Y IS NOT WRITTEN.  Otherwise the library issues
 * the calls above (hidden_ws as in cfm_ffn).
 */
int cfm_ffn_chain(const void* y, int M, int d, int F, int dtype,
                  const void* W1a, const float* b1a, const void* W2a, const float* b2a, float alpha_a,
                  const float* g1a, const float* be1a, const float* g2a, const float* be2a,
                  const void* W1b, const float* b1b, const void* W2b, const float* b2b, float alpha_b,
                  const float* g1b, const float* be1b, const float* g2b, const float* be2b,
                  float* X, void* Y, const uint8_t* y_row_valid,
                  const void* Wp, const float* bp, void* P, int Np,
                  float eps, void* hidden_ws, int engine, void* stream);

/*
 * Scaled-dot-product attention with the reference's mask semantics (attention.py:84-97,
 * 160-174): scores = (q . k'_j + key_bias_j) * scale; positions whose mask byte is 0 get -inf,
 * softmax over keys, masked probabilities forced to 0 (a fully masked row yields 0), then . v.
 *   q: (B,Tq,H,64) act dtype with strides (q_bs, q_ts) in elements, head h at offset h*64
 *   k,v: (B,Tk,H,64) likewise.   out: (B,Tq,H*64) act dtype, contiguous rows of H*64.
 *   mask: uint8, NULL = no mask; element (b,i,j) at mask[b*mask_bs + i*mask_rs + j]
 *         (mask_rs = 0 broadcasts one key row over all queries, i.e. the (B,1,Tk) pad mask).
 *   key_bias: fp32 (B,H,Tk) or NULL  (streaming position term, see cfm_relpos_keys).
 * Head dim is 64 (d/H for both Conformer-M and -L).
 */
int cfm_attention(const void* q, int64_t q_bs, int64_t q_ts,
                  const void* k, int64_t k_bs, int64_t k_ts,
                  const void* v, int64_t v_bs, int64_t v_ts,
                  void* out, int B, int H, int Tq, int Tk,
                  const uint8_t* mask, int64_t mask_bs, int64_t mask_rs,
                  const float* key_bias, float scale, int dtype, int engine, void* stream);

/*
 * Streaming relative-position fold (attention.py:78-88 with P == Tk, B == 1).
 * With q' = q + pos_bias_u folded into the q bias:  ac + bd = q'.(k_j + p_j) + (v - u).p_j
 *   k_out[b,j,h,:] = k[b,j,h,:] + p[b,j,h,:]        key_bias[b,h,j] = sum_c (vbias-ubias)[h,c] * p[b,j,h,c]
 * k: (B,Tk,H,64) act dtype strides (k_bs,k_ts); p: (B,Tk,H*64) act dtype, rows contiguous, batch stride
 * p_bs elements (0 = one table shared by the batch); k_out contiguous (B,Tk,H,64); u,vb fp32 (H,64).
 */
int cfm_relpos_keys(const void* k, int64_t k_bs, int64_t k_ts, const void* p, int64_t p_bs,
                    const float* u, const float* vb, void* k_out, float* key_bias,
                    int B, int H, int Tk, int dtype, void* stream);

/*
 * Depthwise conv1d along time + per-channel affine + SiLU on channel-last activations:
 *   y[b,t,c] = silu( sum_j w[j,c] * x[b,t+j-(k-1)/2,c] + bias[c] ),  zero outside [0,T)
 * Replaces depthwise_conv -> BatchNorm1d(eval) -> SiLU (convolution.py:43-45): the host folds the
 * BatchNorm running statistics and the conv bias into (w, bias).  x,y: (B,T,d) act dtype;
 * w: (k,d) fp32 (tap-major); bias: (d) fp32.  With apply_silu == 0 the raw conv+bias is written as
 * fp32 (training path, BatchNorm batch statistics follow).  d % 64 == 0, k odd <= 31.
 */
int cfm_dwconv(const void* x, const float* w, const float* bias, void* y,
               int B, int T, int d, int k, int dtype, int apply_silu, void* stream);

/*
 * Self-attention core + output projection on the residual stream, in place (attention.py:84-99 + encoder_layer.py:60-63):
 *   ctx = cfm_attention(q, k, v, mask, key_bias, scale)   (all H heads, head dim 64, d = 64 H)
 *   X  += Wo ctx + bo,  then, if g1 != NULL, Y = ymask(LN(X; g1, be1))
 * q, k, v, mask, key_bias, scale: exactly as in cfm_attention (q: (B,Tq,H,64) view, k/v: (B,Tk,H,64) views);
 * Wo: (d,d) act dtype, bo (d) fp32; X: (B*Tq, d) fp32 contiguous; Y: (B*Tq, d) act dtype contiguous.
 * On the tcgen05 engine (bf16, H == 4, Tq == Tk <= 256, no key_bias) this is ONE kernel per 128 query rows and the
 * context never touches HBM; otherwise the library runs cfm_attention into ctx_ws ((B*Tq, d) act dtype, may be NULL
 * only if the fused path applies) followed by cfm_gemm / cfm_gemm_ln.
 */
int cfm_mhsa_out(const void* q, int64_t q_bs, int64_t q_ts, const void* k, int64_t k_bs, int64_t k_ts,
                 const void* v, int64_t v_bs, int64_t v_ts, int B, int H, int Tq, int Tk,
                 const uint8_t* mask, int64_t mask_bs, int64_t mask_rs, const float* key_bias, float scale,
                 const void* Wo, const float* bo, float* X, int dtype,
                 const float* g1, const float* be1, void* Y, const uint8_t* y_row_valid, float eps,
                 void* ctx_ws, int engine, void* stream);

/*
 * Whole convolution module (inference: BatchNorm running statistics folded) on the residual stream, in place
 * (convolution.py:34-49 + encoder_layer.py:64-67):
 *   X += rowmask( W2 silu( dw( glu( W1 y + b1 ) ) ) + b2 ),  then, if g1 != NULL, Y = LN(X; g1, be1)
 * y: (B*T,d) act dtype, already zero on padded rows; W1: (2d,d) [value rows; gate rows], W2: (d,d) act dtype;
 * b1 (2d), b2 (d) fp32; dw_w (k,d) / dw_b (d): folded depthwise taps as in cfm_dwconv; X: (B*T,d) fp32;
 * row_valid (B*T) or NULL; Y (B*T,d) act dtype, must NOT alias y (tiles re-read a halo of neighbouring rows of y).
 * On the tcgen05 engine (bf16, d == 256, k == 15, T >= 15) this is ONE kernel: the GLU output and the depthwise
 * output live in shared memory only.  Otherwise the library runs cfm_gemm (GLU) + cfm_dwconv + cfm_gemm / cfm_gemm_ln
 * through glu_ws / dw_ws, two caller-provided (B*T,d) act-dtype scratch buffers (may be NULL only if the fused path
 * applies).
 */
int cfm_conv_module(const void* y, const void* W1, const float* b1, const float* dw_w, const float* dw_b,
                    const void* W2, const float* b2, float* X, int B, int T, int d, int k, int dtype,
                    const uint8_t* row_valid, const float* g1, const float* be1, void* Y, float eps,
                    void* glu_ws, void* dw_ws, int engine, void* stream);

/*
 * CTC head, greedy part (scope row f2): ids[m] = argmax_v (x[m,:] . W[v,:] + bias[v]), the first index of the maximum;
 * best[m] (optional) = that logit.  W: (V,d) act dtype = ctc_lo.weight (decoder.py:14), bias (V) fp32 or NULL,
 * x: (M,d) act dtype with row stride ldx.  Greedy decoding is this argmax followed by collapsing repeats and dropping
 * blank 0 on the host.  On the tcgen05 engine (bf16, d % 64 == 0, M >= 64) the (M,V) logits are never written: the
 * GEMM epilogue reduces each tile to per-row (max, argmax) pairs and combines tiles with a 64-bit atomicMax.  The
 * CUDA-core engine goes through the logit workspace in chunks of 1024 rows.  ws: cfm_ctc_ws_bytes(M, V, dtype) bytes.
 */
int64_t cfm_ctc_ws_bytes(int M, int V, int dtype);
int cfm_ctc_argmax(const void* x, int ldx, const void* W, const float* bias, int M, int V, int d, int dtype,
                   int32_t* ids, float* best, void* ws, int engine, void* stream);

/*
 * BatchNorm1d training-mode pieces (convolution.py:44): statistics over all B*T rows, unmasked.
 *   cfm_bn_stats : sum[c] = sum_r x[r,c], sumsq[c] = sum_r (x[r,c])^2   (x fp32 (rows,d); outputs must be zeroed)
 *   cfm_bn_apply_silu : y = silu((x-mean)*rstd*gamma+beta) in act dtype
 */
int cfm_bn_stats(const float* x, int rows, int d, float* sum, float* sumsq, void* stream);
int cfm_bn_apply_silu(const float* x, int rows, int d, const float* mean, const float* rstd,
                      const float* gamma, const float* beta, void* y, int dtype, void* stream);

/*
 * Conv2d sub-sampling front-end on the bf16 path (scope row f1; convolution.py:52-76 without the final Linear, which
 * is a cfm_gemm):  out = relu(conv2(relu(conv1(x)))), both convolutions 3x3 stride 2, no padding.
 *   x   : (B, Tin, idim) fp32                     w1 : (C, 9) fp32 = conv.0.weight (C,1,3,3), b1 (C) fp32
 *   w2  : (C, 9*C) bf16 with k = (i*3+j)*C + ci   (= conv.2.weight.permute(0,2,3,1)), b2 (C) fp32
 *   ws  : scratch of cfm_subsample_ws_bytes() bytes (the parity-split channels-last conv1 activation)
 *   out : (B, T2, F2, C) bf16, i.e. the (B*T2, F2*C) row-major input of the Linear with columns ordered f*C + c
 *         (the reference orders them c*F2 + f: permute the Linear weight accordingly).
 * T1 = (Tin-3)/2+1, T2 = (T1-3)/2+1, F1 = (idim-3)/2+1, F2 = (F1-3)/2+1.  C % 256 == 0.
 */
int64_t cfm_subsample_ws_bytes(int B, int Tin, int idim, int C);
int cfm_subsample_conv(const float* x, int B, int Tin, int idim, const float* w1, const float* b1,
                       const void* w2, const float* b2, int C, void* ws, void* out, void* stream);

/* Pull `bytes` at p into L2 (prefetch.global.L2, no registers / shared memory: the CTAs run on SMs the concurrently
 * running single-wave kernels leave idle).  Used to warm the NEXT layer's weights from a side stream. */
int cfm_l2_prefetch(const void* p, int64_t bytes, int blocks, void* stream);
/* the same for up to 8 regions in one launch (host arrays of n pointers / sizes) */
int cfm_l2_prefetch_multi(const void* const* ptrs, const int64_t* bytes, int n, int blocks, void* stream);

/*
 * General GEMM of the training path (backward of every nn.Linear / 1x1 Conv1d on the path, and the batched products of
 * attention forward/backward, attention.py:84,96 and their autograd):
 *     C[b][h] (M x N)  (+)=  alpha * opA(A[b][h]) (M x K)  opB(B[b][h])^T (N x K)
 * a_mn_major / b_mn_major = 0: operand is K-major, element (mn, k) at base + mn*ld + k (nn.Linear forward layout);
 *                         = 1: MN-major, element (mn, k) at base + k*ld + mn (the transposed view: no copy is made).
 *   dgrad  dA = dC W   : A = dC (0), B = W (1, ldb = row stride of W)
 *   wgrad  dW += dC^T X: A = dC (1), B = X (1), C = fp32 gradient, accumulate = 1 (split over K = tokens; the K slices
 *                        are combined with TMA reduce-add stores)
 * *_hs / *_bs: element strides of the head / batch dimension (nH, nB >= 1).  in_dtype CFM_F32 | CFM_BF16 (A and B);
 * c_dtype CFM_F32 | CFM_BF16; accumulate needs an fp32 C.  splits: 0 = automatic.  tcgen05 engine: bf16 inputs,
 * 16-byte aligned bases and strides; everything else runs on the CUDA-core engine.
 */
int cfm_gemm_ex(const void* A, int a_mn_major, int64_t lda, int64_t a_hs, int64_t a_bs,
                const void* B, int b_mn_major, int64_t ldb, int64_t b_hs, int64_t b_bs,
                void* C, int c_dtype, int64_t ldc, int64_t c_hs, int64_t c_bs, int accumulate,
                int M, int N, int K, int nH, int nB, int in_dtype, float alpha, int splits, int engine, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * TRAINING path (BASELINE.json configs[4]): the pieces of forward that must save state and every non-GEMM backward.
 * They implement the autograd of the reference's modules (module.py:49-69 drives fwd+bwd through encoder.py:54-75).
 * `dtype` is the activation dtype (CFM_F32 | CFM_BF16); statistics, the residual stream x and all parameter
 * gradients are fp32; gradient outputs named d<param> are ACCUMULATED (+=, atomics).  Dropout is counter based
 * (Philox4x32-10 keyed by (seed, site), counter = element index): forward and backward pass the same (p, seed, site)
 * and no mask is stored.  `seed` is a DEVICE pointer to one uint64 (so that a captured CUDA graph of a training step
 * replays with a new seed); p = 0 (or a null seed) disables dropout.  The reference's dropout sites: feedforward.py:19, attention.py:95,
 * encoder_layer.py:58,62,66,69.
 */
/* y = rowmask(LN(x; g, b)) and the row statistics needed by the backward (encoder_layer.py:56,59,63,67,70). */
int cfm_ln_fwd_train(const float* x, int rows, int d, const float* g, const float* b, void* y, int y_dtype,
                     float* mean, float* rstd, const uint8_t* row_valid, float eps, void* stream);
/* dx_out = (dx_in ? dx_in : 0) + LayerNorm-backward(rowmask(dy));  dg += sum dy*xhat;  db += sum dy. */
int cfm_ln_bwd(const void* dy, int dy_dtype, const float* x, const float* mean, const float* rstd, const float* g,
               const uint8_t* row_valid, const float* dx_in, float* dx_out, float* dg, float* db, int rows, int d,
               void* stream);
/* a = dropout(SiLU(h))  (feedforward.py:18-19) and its backward dh = da * mask * SiLU'(h), dbias += colsum(dh). */
int cfm_silu_dropout_fwd(const void* h, void* a, int rows, int cols, int dtype, float p, const uint64_t* seed, int site,
                         void* stream);
int cfm_silu_dropout_bwd(const void* da, const void* h, void* dh, float* dbias, int rows, int cols, int dtype, float p,
                         const uint64_t* seed, int site, void* stream);
/* x = x_in + alpha * rowmask * dropout(f)  (the residual adds of encoder_layer.py:58,62,66,69 with their dropout and the
 * masked_fill of convolution.py:47-48; x_in null or == x: in place) and its backward df = alpha * rowmask * mask * dx,
 * dbias += colsum(df). */
int cfm_resid_dropout_add(const float* x_in, float* x, const void* f, int rows, int cols, int dtype, float alpha,
                          const uint8_t* row_valid, float p, const uint64_t* seed, int site, void* stream);
int cfm_scale_dropout_bwd(const float* dx, void* df, float* dbias, int rows, int cols, int dtype, float alpha,
                          const uint8_t* row_valid, float p, const uint64_t* seed, int site, void* stream);
/* GLU over the channel halves of g (rows, 2d) (convolution.py:42) and its backward (dbias: 2d bias gradient of
 * pointwise_conv1). */
int cfm_glu_fwd(const void* g, void* u, int rows, int d, int dtype, void* stream);
int cfm_glu_bwd(const float* du, const void* g, void* dg, float* dbias, int rows, int d, int dtype, void* stream);
/* Backward of SiLU(BatchNorm1d(raw)) (convolution.py:44-45): sums[0..d) = dgamma, sums[d..2d) = dbeta of THIS call
 * (overwritten), draw = gradient w.r.t. the depthwise-conv output.  batch_stats = 1: mean / rstd are the batch
 * statistics of `raw` (training mode: the statistics depend on the input); 0: running statistics (eval mode). */
int cfm_bn_silu_bwd(const void* dc, const float* raw, const float* mean, const float* rstd, const float* gamma,
                    const float* beta, float* sums, void* draw, int rows, int d, int dtype, int batch_stats, void* stream);
/* Weight / bias gradient of the depthwise Conv1d (convolution.py:43): dw (k, d) += ..., dbias (d) += ...; the input
 * gradient is cfm_dwconv with the tap-reversed filter. */
int cfm_dwconv_wgrad(const void* dy, const void* u, float* dw, float* dbias, int B, int T, int d, int k, int dtype,
                     void* stream);
/* Masked softmax (+ dropout) over materialised scores S (B,H,Tq,Tp) fp32 -> P (and the dropped copy Pd when p > 0), mask
 * semantics of attention.py:89-92; and its backward dS = P * (dP - rowsum(dP * P)), dP = dPd * dropout multiplier. */
int cfm_softmax_fwd(const float* S, void* P, void* Pd, const uint8_t* mask, int64_t mask_bs, int64_t mask_rs, int B, int H,
                    int Tq, int Tk, int Tp, int dtype, float p, const uint64_t* seed, int site, void* stream);
int cfm_softmax_bwd(const void* P, const float* dPd, void* dS, int B, int H, int Tq, int Tk, int Tp, int dtype, float p,
                    const uint64_t* seed, int site, void* stream);
/* out (cols) += column sums of x (rows, ld) -- bias gradients. */
int cfm_colsum(const void* x, int64_t ld, float* out, int rows, int cols, int dtype, void* stream);
/* Optimizer step of the training loop (module.py:140-143: torch.optim.Adam(self.parameters(), lr), default betas / eps, no
 * weight decay): one Adam update over a flat fp32 segment of n elements, torch.optim.Adam arithmetic,
 *   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr / (1-b1^step) * m / (sqrt(v) / sqrt(1-b2^step) + eps)
 * with g = grad_scale * grad.  `step` counts from 1.  p_bf16 (may be NULL) receives the bf16 copy of the updated values. */
int cfm_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, double lr, double beta1, double beta2,
                  double eps, int step, float grad_scale, void* stream);

/* CTC loss (decoder.py:18-23: log_softmax + nn.CTCLoss(reduction='sum'), blank = 0) on logits (B*T, ld), V valid
 * columns.  fwd: nll[b]; bwd: dlogits = scale * (softmax - occupancy) for valid frames, 0 elsewhere (may alias logits).
 * `ws` (cfm_ctc_loss_ws_bytes) carries lse / alpha / beta from fwd to bwd. */
int64_t cfm_ctc_loss_ws_bytes(int B, int T, int Lmax);
int cfm_ctc_loss_fwd(const void* logits, int64_t ld, int B, int T, int V, const int* labels, int Lmax, const int* in_len,
                     const int* lab_len, float* nll, void* ws, int dtype, void* stream);
int cfm_ctc_loss_bwd(const void* logits, int64_t ld, int B, int T, int V, int Vp, const int* labels, int Lmax,
                     const int* in_len, const int* lab_len, const float* nll, const void* ws, float scale, void* dlogits,
                     int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Feature front-end (scope row f1): Kaldi-compatible log-mel filterbank + global CMVN, replacing
 * torchaudio.compliance.kaldi.fbank as called at processor.py:185-191 and GlobalCMVN.forward (cmvn.py:22-33).  fp32.
 * Pipeline: cfm_fbank_frames -> cfm_gemm_ex (DFT as a GEMM with a (514, 400) [cos | -sin] basis) -> cfm_fbank_power ->
 * cfm_gemm_ex (mel filterbank) -> cfm_fbank_log_cmvn.
 */
/* wave (B, wave_bs) already in int16 range; n_samples (B); window (400) = povey; frames (B*m_max, 400): framing (25 ms /
 * 10 ms, snip_edges), DC removal, pre-emphasis, window; frames past an utterance's last frame are zero. */
int cfm_fbank_frames(const float* wave, int64_t wave_bs, const int* n_samples, const float* window, float* frames, int B,
                     int m_max, float preemph, void* stream);
/* spec (rows, ld_spec) = [re(bins) | im(bins)] -> power (rows, ld_pow), columns >= bins zeroed. */
int cfm_fbank_power(const float* spec, int ld_spec, float* power, int ld_pow, int64_t rows, int bins, void* stream);
/* out = (log(max(mel, eps)) - mean) * istd for valid frames, (0 - mean) * istd for padding frames (mean / istd may be
 * null: plain log-mel with zero padding). */
int cfm_fbank_log_cmvn(const float* mel, float* out, const int* n_samples, const float* mean, const float* istd, int B,
                       int m_max, int nmel, void* stream);
/* y = (x - mean[c]) * istd[c] over n elements with innermost size d (istd may be null: norm_var = False). */
int cfm_cmvn(const float* x, float* y, const float* mean, const float* istd, int64_t n, int d, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * RNN-T joint + loss (scope row f4): joint.py:20-38 and torchaudio.functional.rnnt_loss as called at model.py:95-113.
 * The three Linear layers of the joint run on cfm_gemm / cfm_gemm_ex; these are the pieces in between.
 */
/* z[b,t,u,:] = tanh(e[b,t,:] + p[b,u,:]);  e (B*T, J), p (B*U1, J), z (B*T*U1, J). */
int cfm_joint_add_tanh(const void* e, const void* p, void* z, int B, int T, int U1, int J, int dtype, void* stream);
/* dz <- dz * (1 - z^2) in place, de (B*T, J) = sum over u, dp (B*U1, J) = sum over t. */
int cfm_joint_tanh_bwd(void* dz, const void* z, void* de, void* dp, int B, int T, int U1, int J, int dtype, void* stream);
/* Transducer loss on logits (B*T*U1, ld), V valid columns, fused log-softmax: fwd -> nll[b]; bwd -> dlogits (may alias
 * logits) = scale * d nll_b / d logits, zeros outside each utterance's (t_len, u_len) lattice and in columns [V, Vp).
 * targets (B, Umax) int32, U1 = Umax + 1.  ws (cfm_rnnt_loss_ws_bytes) carries lse / alpha / beta from fwd to bwd. */
int64_t cfm_rnnt_loss_ws_bytes(int B, int T, int U1);
int cfm_rnnt_loss_fwd(const void* logits, int64_t ld, int B, int T, int U1, int V, int blank, const int* targets, int Umax,
                      const int* t_len, const int* u_len, float* nll, void* ws, int dtype, void* stream);
int cfm_rnnt_loss_bwd(const void* logits, int64_t ld, int B, int T, int U1, int V, int Vp, int blank, const int* targets,
                      int Umax, const int* t_len, const int* u_len, const float* nll, const void* ws, float scale,
                      void* dlogits, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CFM_B200_H_ */
