/*
 * movii_b200 — C ABI of the B200-native (sm_100a) DiT / WanVAE hot path.
 *
 * The reference (ZulutionAI/MoviiGen1.1) is pure Python and has no FFI; each entry point below
 * replaces the library call the reference makes at the cited call site (paths relative to the
 * reference repo).  Conventions:
 *   - plain pointers + sizes only; every pointer is DEVICE memory owned by the caller (PyTorch);
 *     the library never allocates device memory and never synchronises the device;
 *   - every launch goes to the caller-supplied stream (a cudaStream_t passed as void*);
 *   - return value: MV_OK (0) or a negative MV_E_* code; the message is in mv_last_error()
 *     (thread-local);  no exceptions, no exit();
 *   - the library refuses to run on anything that is not compute capability 10.x
 *     (MV_E_ARCH) — there is no fallback path of any kind.
 *   - bf16 tensors are row-major with an explicit leading dimension in ELEMENTS.
 */
#ifndef MOVII_B200_H_
#define MOVII_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MV_OK 0
#define MV_E_SHAPE (-1)
#define MV_E_ARCH (-2)
#define MV_E_CUDA (-3)
#define MV_E_NCCL (-4)
#define MV_E_WORKSPACE (-5)

typedef void* mv_stream_t; /* cudaStream_t */

/* GEMM epilogues (mv_gemm_bf16). acc = fp32 accumulator, y = bf16(acc + bias) as nn.Linear under
 * bf16 autocast returns it. */
#define MV_EPI_BF16 0       /* out_bf16[m,n]  = y                                   (model.py:139-141,171-173) */
#define MV_EPI_BF16_GELU 1  /* out_bf16[m,n]  = bf16(gelu_tanh(float(y)))           (model.py:267-268)         */
#define MV_EPI_RESID_F32 2  /* out_f32[m,n]  += float(y) * (gate ? gate[n] : 1)     (model.py:302,306,309)     */
#define MV_EPI_F32_ROUND 3  /* out_f32[m,n]   = float(y)                            (model.py:529)             */
#define MV_EPI_F32 4        /* out_f32[m,n]   = acc + bias (no bf16 rounding; VAE attention scores, vae.py:252) */

const char* mv_last_error(void);
int mv_version(void);
/* MV_OK when the current device is sm_100 (B200); MV_E_ARCH otherwise; MV_E_CUDA if no device. */
int mv_device_check(void);

/* ---- dense contractions ---------------------------------------------------------------------- */

/* out = epilogue(A[M,K] . W[N,K]^T + bias[N]); A, W bf16 (K contiguous, lda/ldw in elements,
 * multiples of 8), fp32 accumulate on tcgen05 tensor cores.  Replaces nn.Linear under bf16 autocast
 * (cuBLASLt) at wan/modules/model.py:139-141,155,171-173,180,267-269,451-453 and the patch-embedding
 * Conv3d at :445-450,529.  bias may be NULL; gate (fp32[N]) only for MV_EPI_RESID_F32, may be NULL. */
int mv_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* out,
                 int64_t ldo, const float* gate, int M, int N, int K, int epilogue, mv_stream_t stream);

/* mv_gemm_bf16 with IEEE fp16 operands and fp16 (MV_EPI_BF16 / _GELU slots) or fp32 outputs: the two contractions of the
 * WanVAE AttentionBlock (vae.py:246-257), whose activations are fp16 here. */
int mv_gemm_f16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* out, int64_t ldo,
                const float* gate, int M, int N, int K, int epilogue, mv_stream_t stream);

/* Same contraction with A stored as K/a_kblock slabs: A[m, j*a_kblock + c] = A_base[j*a_block_stride + m*lda + c].
 * This is the layout the Ulysses attention-output all-to-all delivers ([src rank][local token][local heads*128],
 * xdit_context_parallel.py:185-197), so the o projection consumes it without a transpose pass.
 * a_kblock must be a multiple of 64. */
int mv_gemm_bf16_ksplit(const void* A, int64_t lda, int64_t a_block_stride, int a_kblock, const void* W, int64_t ldw,
                        const float* bias, void* out, int64_t ldo, const float* gate, int M, int N, int K,
                        int epilogue, mv_stream_t stream);

/* o[Lq,H,128] = softmax(q k^T * scale) v, non-causal, bf16 in/out, fp32 softmax and accumulation;
 * q,k,v,o are [L, H, 128] views with row strides ldq/ldk/ldv/ldo (elements) and heads contiguous
 * (head h at column offset h*128).  Keys/values are rows [0, Lk).  Replaces
 * flash_attn.flash_attn_varlen_func at wan/modules/attention.py:113-127 (self- and cross-attention,
 * model.py:146-151,176). */
int mv_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                     void* o, int64_t ldo, int Lq, int Lk, int H, float softmax_scale, mv_stream_t stream);

/* ---- HBM-bound fused row kernels -------------------------------------------------------------- */

/* out_bf16[m,:] = bf16( LN(x[m,:]) [*w + b] [rounded to bf16 if round_ln] * (1 + scale) + shift ).
 * LN is affine-free unless w/b given; fp32 statistics, eps as given.  shift/scale may be NULL (no
 * modulation).  Replaces WanLayerNorm + adaLN modulation + autocast input cast,
 * wan/modules/model.py:89-99,299,306,307. */
int mv_ln_modulate(const float* x, int64_t ldx, const float* shift, const float* scale, const float* w,
                   const float* b, void* out_bf16, int64_t ldo, int M, int C, float eps, int round_ln,
                   mv_stream_t stream);

/* In place on a bf16 [M, C] slab (row stride ld): full-row WanRMSNorm (fp32 stats, bf16 rounding
 * before the weight multiply, model.py:70-86) followed, when cs != NULL, by 3-axis RoPE on
 * interleaved pairs with a per-token cos/sin table cs[M][head_dim/2][2] fp32 (model.py:39-67;
 * xdit_context_parallel.py:24-62 for the rank-offset table), output rounded to bf16
 * (attention.py:59-83). */
int mv_rmsnorm_rope(void* x_bf16, int64_t ld, const float* weight, const float* cs, int M, int C,
                    int head_dim, float eps, mv_stream_t stream);

/* mv_rmsnorm_rope with the Ulysses head scatter fused in (xdit_context_parallel.py:169-190): when out != NULL the
 * result goes to out[dst][row][C/sp_world] (dst = destination rank of the head group), the all-to-all send
 * layout, instead of in place; weight == NULL skips the norm (plain scatter of V); cs == NULL skips RoPE. */
int mv_qkv_prepare(void* x_bf16, int64_t ld, const float* weight, const float* cs, void* out_bf16, int sp_world,
                   int M, int C, int head_dim, float eps, mv_stream_t stream);

/* q, k (and, when scattering, v) of a fused QKV row [M, 3C] (row stride ld) in ONE launch: full-row WanRMSNorm with
 * gain_q / gain_k + RoPE (cs != NULL) on the q and k column slabs [0,C) and [C,2C) — what model.py:139-148 does with
 * five ATen launches per tensor.  n_dst == 0: in place (v untouched).  n_dst > 0 (Ulysses, xdit_context_parallel.py:
 * 169-190): head group d of local row m of q / k / v is stored to dst_x[d] + (src_slot*M + m)*(C/n_dst), i.e. into
 * slab `src_slot` of destination d's [slot][M][C/n_dst] buffer — local send buffers (NCCL mode: dst_x[d] = base +
 * d*M*(C/n_dst), src_slot = 0) or NVLink-mapped peer receive buffers (src_slot = this rank). */
int mv_qkv_norm_rope(void* qkv_bf16, int64_t ld, const float* gain_q, const float* gain_k, const float* cs, int M, int C,
                     int head_dim, float eps, void* const* dst_q, void* const* dst_k, void* const* dst_v, int n_dst,
                     int src_slot, mv_stream_t stream);

/* Classifier-free guidance + one FlowUniPC multistep update in a single pass over the latent
 * (wan/text2video.py:245-254; wan/utils/fm_solvers_unipc.py:319-332,351-485,487-627):
 *   noise = uncond + guide*(cond - uncond);  x0 = sample - sigma*noise;
 *   corrected = UniC(last_sample, history, x0)   (when use_corrector);   prev = UniP(corrected, x0, history)
 * hist0/1/2 = model_outputs[-1/-2/-3] BEFORE this step (NULL when unused).  coef: MV_UNIPC_NCOEF host floats
 * {guide, sigma, use_corrector, c_order, c_ratio, c_a, c_b, c_rho_last, c_rk0, c_rk1, c_rho0, c_rho1,
 *  p_order, p_ratio, p_a, p_b, p_rk0, p_rk1, p_rho0, p_rho1} computed by the scheduler on the host exactly as the
 * reference does.  Outputs: x0_out (new model_outputs[-1]), sample_out (corrected sample = next last_sample),
 * prev_out (next latent).  Element-wise fp32 in the reference's operation order (bit-identical to the ATen chain). */
#define MV_UNIPC_NCOEF 20
int mv_unipc_cfg_step(const float* cond, const float* uncond, const float* sample, const float* last_sample,
                      const float* hist0, const float* hist1, const float* hist2, float* x0_out, float* sample_out,
                      float* prev_out, int64_t n, const float* coef, int ncoef, mv_stream_t stream);

/* out[l, :] = mods[l, :] + e0[:], fp32: the per-layer `modulation + e` tables of all blocks (model.py:292-295) or of
 * the head (:341) in one launch. */
int mv_modulation_table(const float* mods, const float* e0, float* out, int layers, int len, mv_stream_t stream);

/* A_bf16[L, C*ph*pw] <- latent fp32 [C, F, H, W], patch (1,ph,pw); column = c*ph*pw + i*pw + j, the
 * flatten(1) order of patch_embedding.weight (model.py:445-450,529-533). */
int mv_patchify(const float* latent, void* a_bf16, int C, int F, int H, int W, int ph, int pw,
                mv_stream_t stream);

/* Head (model.py:333-343) + unpatchify (:581-609), all fp32:
 * out[c, f, h*ph+i, w*pw+j] = sum_k (LN(x[n,:])*(1+scale)+shift)[k] * Wh[(i*pw+j)*Cout + c, k] + bh[..]
 * for token n = (f,h,w) < F*Hp*Wp. */
int mv_head_unpatchify(const float* x, int64_t ldx, const float* shift, const float* scale, const float* Wh,
                       const float* bh, float* out, int F, int Hp, int Wp, int ph, int pw, int Cout, int C,
                       float eps, mv_stream_t stream);

/* Head only, token-major: out_tokens[L, nout] (fp32) — the sequence-parallel path all-gathers these rows
 * (xdit_context_parallel.py:145-148) and then calls mv_unpatchify. */
int mv_head_tokens(const float* x, int64_t ldx, const float* shift, const float* scale, const float* Wh,
                   const float* bh, float* out_tokens, int L, int nout, int C, float eps, mv_stream_t stream);
/* tokens[L, ph*pw*Cout] -> out[Cout, F, Hp*ph, Wp*pw] (model.py:581-609). */
int mv_unpatchify(const float* tokens, float* out, int F, int Hp, int Wp, int ph, int pw, int Cout,
                  mv_stream_t stream);

/* out[N] = W[N,K] . act(x[K]) + b[N], fp32 (M = 1).  act_in: 0 none, 1 SiLU.  Replaces the fp32
 * time_embedding / time_projection Linears (model.py:455-457,541-545). */
int mv_linear_f32_vec(const float* x, const float* W, const float* b, float* out, int N, int K, int act_in,
                      mv_stream_t stream);

/* out[dim] = cat(cos(t*w_i), sin(t*w_i)), w_i = 10000^(-i/(dim/2)), computed in fp64, stored fp32
 * (sinusoidal_embedding_1d, model.py:15-25).  t: device int64/fp32 scalar. */
int mv_sinusoid_embed(const void* t, int t_is_int64, float* out, int dim, mv_stream_t stream);

/* ---- fused Ulysses exchange over NVLink peer memory (wan/distributed/xdit_context_parallel.py:155-198) --------- */
/* The all-to-alls of the reference (xfuser -> NCCL) are fused into the producing kernels: each rank maps every peer's
 * receive buffers (CUDA IPC) and the kernels store into them directly; a flag barrier orders producers and consumers.
 * Receive buffer layout on every rank: [src rank][local token][heads_per_rank*128] (bf16). */

/* Export / map a caller-owned device buffer.  handle64: 64 bytes; offset: of dptr inside its allocation. */
int mv_ipc_export(const void* dptr, void* handle64, int64_t* offset, int64_t* alloc_bytes);
int mv_ipc_open(const void* handle64, void** base_out);
int mv_ipc_close(void* base);

/* Cross-GPU barrier on the caller's stream: publishes `epoch` into slot `rank` of every peer's flag array
 * (peer_flag_ptrs[i] = mapped base of rank i's uint32[8] flags; entry `rank` = own), then waits until all `world`
 * local slots reached `epoch`.  Release/acquire at system scope: everything the previous kernels of this stream
 * stored to peers is visible to the peers' next kernels.  Bounded wait (traps after 20 s instead of hanging). */
int mv_sp_barrier(void* const* peer_flag_ptrs, void* local_flags, int rank, int world, unsigned int epoch,
                  mv_stream_t stream);

/* mv_qkv_prepare with the head scatter going straight to the destination ranks: head group d of local token m is
 * written to dst_ptrs[d][src_rank][m][C/sp_world] (host array of sp_world mapped device pointers). */
int mv_qkv_prepare_p2p(void* x_bf16, int64_t ld, const float* weight, const float* cs, void* const* dst_ptrs,
                       int src_rank, int sp_world, int M, int C, int head_dim, float eps, mv_stream_t stream);

/* mv_attention_fwd whose epilogue returns each query row to its owner rank: row r -> o_dst[r / rows_per_rank]
 * [src_rank][r % rows_per_rank][H*128] (ldo = row stride of the receive buffer, normally H*128). */
int mv_attention_fwd_scatter(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                             void* const* o_dst, int n_dst, int src_rank, int rows_per_rank, int64_t ldo, int Lq,
                             int Lk, int H, float softmax_scale, mv_stream_t stream);

/* ---- WanVAE decoder (wan/modules/vae.py) ------------------------------------------------------- */
/* Activations are channels-last FP16 [T, H, W, C]; weights are packed per conv as fp16 [Cout_pad][tap][Cin]; fp32
 * accumulation.  (fp16 keeps the 10 mantissa bits of the TF32 convolutions the reference runs with autocast disabled.) */

/* Stride-1 "same" convolution as an implicit GEMM on tcgen05 (csrc/vae_conv_sm100.cu):
 *   out[t,h,w,:] = bias + sum_i in[t+dt_i, h+dh_i, w+dw_i, :] . W[:, i, :]^T (+ res[t,h,w,:])
 * with zero fill outside the input grid (spatial zero padding and the causal temporal padding of
 * CausalConv3d, vae.py:17-36).  taps = ntaps (dt,dh,dw) int8 triples (host memory).  The output element offset of
 * voxel (t,h,w) is o_base + t*os_t + h*os_h + w*os_w, which expresses the sub-pixel (nearest-2x + Conv2d,
 * vae.py:74-79) and frame-interleave (time_conv, vae.py:128-137) stores; output channel blocks >= nsplit are
 * written to channel (n - nsplit) at offset + nsplit_off (nsplit = 0 disables).
 * Temporal chunking (the reference's feature cache, vae.py:28-36,205-217): the input holds in_T = t_off + out_T frames,
 * the first t_off being the cached tail of the previous chunk; output frame t reads input frames t + t_off + dt.
 * out_mode 0: fp16 channels-last (+ optional fp16 residual with the same addressing);
 * out_mode 1: fp32 channel-first video clamped to [-1, 1] (decoder head, vae.py:465-471,660-661): element
 * out[c*os_t + o_base + (t*H + h)*W + w] (os_t = channel stride = T_total*H*W, o_base = first frame * H*W; os_t = 0:
 * a dense [cout_real, out_T, H, W] tensor). */
int mv_vae_conv(const void* in_cl, int in_T, int in_H, int in_W, int Cin, const void* w_packed, const float* bias,
                const void* res_cl, void* out, int out_mode, int out_T, int out_H, int out_W, int Cout, int cout_real,
                int ntaps, const int8_t* taps_dt_dh_dw, int64_t o_base, int64_t os_t, int64_t os_h, int64_t os_w,
                int nsplit, int64_t nsplit_off, int t_off, mv_stream_t stream);

/* mv_vae_conv (fp16 channels-last output) with the CONSUMER's RMS_norm + SiLU fused into the epilogue
 * (vae.py:194-199: every conv of a ResidualBlock is fed silu(rms_norm(.))): norm_out[voxel,:] =
 * silu(rms_norm(out[voxel,:]) * gamma) with the same addressing as out; out may be NULL when only the normalised
 * tensor is consumed.  Requires Cout <= 256 (whole channel row in one tile). */
int mv_vae_conv_fused(const void* in_cl, int in_T, int in_H, int in_W, int Cin, const void* w_packed, const float* bias,
                      const void* res_cl, void* out, int out_T, int out_H, int out_W, int Cout, int ntaps,
                      const int8_t* taps_dt_dh_dw, int64_t o_base, int64_t os_t, int64_t os_h, int64_t os_w,
                      const float* norm_gamma, void* norm_out, int t_off, mv_stream_t stream);

/* Diagnostics only: 1 = route the tensor-bound convolutions to the CTA-pair kernel (cta_group::2, NT stacked tiles per
 * CTA, one A box per (dt, dw)), 0 = the single-CTA kernel, -1 = keep, -2 = back to the default (MV_CONV_PAIR or built-in). */
int mv_vae_conv_config(int pair, int tiles_per_cta /* 0 auto, 1 | 2 | 4 */,
                       int epi_regs /* fused norm epilogue: 1 = one TMEM pass, row in registers; 0 = two passes */);

/* y = [silu]( x / max(||x||_2, 1e-12) * sqrt(C) * gamma ) per voxel over channels (RMS_norm + nn.SiLU,
 * vae.py:39-54,194-199); channels-last fp16, in place allowed. */
int mv_vae_rmsnorm_silu(const void* x_cl, void* y_cl, const float* gamma, int64_t nvox, int C, int silu,
                        mv_stream_t stream);

/* x_cl[v, o] = sum_c W2[o,c] * (z[c, v] * std[c] + mean[c]) + b2[o]: latent de-normalisation + conv2 (1x1x1) +
 * fp32 channel-first -> fp16 channels-last (vae.py:547-553,629-639). */
int mv_vae_latent_in(const float* z, const float* W2, const float* b2, const float* mean, const float* stdv,
                     void* out_cl, int Z, int64_t nvox, mv_stream_t stream);

/* ---- WanVAE encoder (vae.py:265-366,516-542) --------------------------------------------------------
 * Strided convolution on the same implicit GEMM: out[t,h,w,:] = bias + sum_taps in[st*t + t_off + dt, sh*h + dh,
 * sw*w + dw, :] . W_tap^T, strides 1 or 2; the A operand is a TMA box with traversal strides, taps running past the far
 * edge read zeros.  Serves Resample 'downsample2d/3d': ZeroPad2d((0,1,0,1)) + Conv2d(3x3, stride 2) (taps dh, dw in
 * 0..2; vae.py:92-104) and the time_conv (3,1,1) stride (2,1,1) on [last cached frame | chunk] (t_off = 2, taps
 * dt = -2..0; vae.py:143-159).  in_T = st*(out_T-1) + t_off + 1.  fp16 channels-last in and out. */
int mv_vae_conv_strided(const void* in_cl, int in_T, int in_H, int in_W, int Cin, const void* w_packed, const float* bias,
                        void* out_cl, int out_T, int out_H, int out_W, int Cout, int ntaps, const int8_t* taps_dt_dh_dw,
                        int t_off, int stride_t, int stride_h, int stride_w, mv_stream_t stream);

/* Frames [t0, t0+n) of a channel-first fp32 video [3, T_total, H, W] -> fp16 channels-last [n, H, W, 16] (channels
 * 3..15 zero), the stem conv's operand (vae.py:286,519-530). */
int mv_vae_video_in(const float* video, int T_total, int t0, int n, int H, int W, void* out_cl, mv_stream_t stream);

/* mu[o, mu_off + v] = ((sum_c W1[o,c] * head[v,c] + b1[o]) - mean[o]) * inv_std[o] for o < Z: conv1 (1x1x1, 2Z -> 2Z,
 * mu half only), `.chunk(2)` and the latent normalisation (vae.py:531-537); head fp16 channels-last [nvox, 2Z], mu fp32
 * channel-first with mu_plane elements per channel. */
int mv_vae_latent_out(const void* head_cl, const float* W1, const float* b1, const float* mean, const float* inv_std,
                      float* mu, int Z, int64_t nvox, int64_t mu_plane, int64_t mu_off, mv_stream_t stream);

/* Second half of the decoder's head conv (96 -> 3, 3x3x3, vae.py:420-421 + the clamp of :660-661) computed as
 *   out[co, v] = clamp(bias[co] + sum_tap D[v + tap][4*tap + co], -1, 1),
 * where D [frames, H, W, 112] fp16 = one 1x1x1 mv_vae_conv of the normalised input with the 27 x (3 + 1 pad) tap-major weight
 * matrix.  d_cur: D of this chunk's n frames; d_prev: D of the kprev (<= 2) frames before it (NULL / 0 at the start of the
 * sequence); bias3_host: the three biases (HOST pointer); video fp32 [3][plane], frame t of the chunk lands at
 * frame_off + t*H*W. */
int mv_vae_head_gather(const void* d_cur, const void* d_prev, int kprev, int n, int H, int W, const float* bias3_host,
                       float* video, int64_t plane, int64_t frame_off, mv_stream_t stream);

/* P_f16[m, :N] = softmax(S[m, :N] * scale) (fp32 in, fp16 out): the softmax of the VAE's single-head attention between the
 * two tcgen05 GEMMs (vae.py:246-257). */
int mv_softmax_rows(const float* S, int64_t lds, void* P_f16, int64_t ldp, int M, int N, float scale,
                    mv_stream_t stream);

/* Diagnostics only: mv_attention_fwd (128-key-step kernel) that also writes clock64 stamps of CTA (0, head 0) to
 * trace[2 tiles][trace_steps][8] (uint64, device memory): 0 scores visible, 1 scores in registers, 2 row max done,
 * 3 exponentials done, 4 P handed over, 5 P seen by the MMA warp, 6 P.V + next Q.K^T issued.  tools/attn_trace.py. */
int mv_attention_fwd_trace(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                           int64_t ldo, int Lq, int Lk, int H, float softmax_scale, unsigned long long* trace,
                           int trace_steps, mv_stream_t stream);

/* Diagnostics only: 1 = route the large GEMMs to the CTA-pair (cta_group::2, 256 x 256 tile) kernel, 0 = single-CTA
 * 128 x 256 tiles, -1 = keep, -2 = back to the default (MV_GEMM_PAIR or built-in).  A/B timing inside one process. */
int mv_gemm_config(int pair);

/* Diagnostics only: 1 = kernels run their TMA-producer / MMA-issuer warps at the highest warp ids of the CTA (the SM
 * sub-partition arbiter favours the highest warp id), 0 = at the lowest, -1 keep, -2 default (MV_ROLES_HI or built-in). */
int mv_roles_config(int hi);

/* Diagnostics only: overrides the attention kernel variant chosen from the environment (MV_ATTN_KSTEP / _EMU / _STALE /
 * _PINGPONG / _SKEW) for A/B timing inside one process; a negative argument keeps the current value.  kstep 64 | 128,
 * emu 0..2 (fraction of exponentials on the FMA pipe: none, 1/4, 1/2), stale 1 = fixed-reference softmax (128-key
 * kernel), skew = one-time start offset of the second Q tile (clocks), wait_spin 1 = the waits on the per-tile chain
 * poll instead of parking the warp, pack 1 = P packed to bf16 by a truncating PRMT of pre-scaled exponentials instead of
 * F2FP conversions.  tools/ab_step.py. */
int mv_attention_config(int kstep, int emu, int stale, int pingpong, int skew, int wait_spin, int pack);

/* ---- umT5 text encoder (caller side of the hot path) --------------------------- */

/* o[Lq,H,64] = softmax(q k^T + bias[h, (j - i) + bias_center] + key mask) v, no 1/sqrt(d) scaling, keys >= kv_len
 * masked out; bf16 in/out, fp32 scores / softmax / accumulation.  q,k,v,o are [L, H, 64] views with row strides
 * in elements; Lk <= 512 (the 128 x 512 score block of a CTA lives in TMEM).  bias is the per-head relative-position
 * table (fp32 [H, bias_ld]) or NULL.  Replaces T5Attention.forward's einsum/softmax/einsum, wan/modules/t5.py:98-115,
 * with the bias that T5RelativeEmbedding.forward (:233-243) materialises as [1,H,Lq,Lk]. */
int mv_t5_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                    int64_t ldo, const float* bias, int64_t bias_ld, int bias_center, int Lq, int Lk, int kv_len,
                    int H, mv_stream_t stream);

/* out_bf16[r,:] = bf16( bf16(x[r,:] * rsqrt(mean(x[r,:]^2) + eps)) * weight ), x fp32: T5LayerNorm,
 * wan/modules/t5.py:61-66 (the encoder keeps its residual stream in fp32 here). */
int mv_t5_rmsnorm(const float* x, int64_t ldx, const float* weight, void* out, int64_t ldo, int rows, int C,
                  float eps, mv_stream_t stream);

/* out_f32[n,:] = float(table_bf16[clamp(ids[n]),:]): nn.Embedding lookup, wan/modules/t5.py:304. */
int mv_embed_gather(const void* table, int64_t ldt, int64_t vocab, const int64_t* ids, float* out, int64_t ldo,
                    int n, int C, mv_stream_t stream);

/* out = bf16(a * b), bf16 [rows, C] each: the fc1(x) * gelu(gate(x)) product, wan/modules/t5.py:137. */
int mv_mul_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo, int rows, int C,
                mv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MOVII_B200_H_ */
