/*
 * humanliff_b200.h -- C ABI of libhumanliff_b200.so (sm_100a only).
 *
 * The reference (skhu101/HumanLiff) has no FFI / operator layer: every FLOP of
 * its two hot paths is a PyTorch ATen call.  Each entry point below names the
 * reference call site(s) (file:line under /root/reference) whose arithmetic it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer unless it
 *     is documented as host memory; the caller owns every buffer;
 *   - `stream` is a cudaStream_t passed as void*; kernels are launched on it,
 *     nothing synchronises, nothing allocates (safe under CUDA-graph capture);
 *   - return value 0 = ok, negative = HL_E_*; hl_last_error() gives the text;
 *   - activations are NHWC ("pixel-major"): element (b, y, x, c) lives at
 *     ptr[((b*H + y)*W + x)*ld + c], `ld` (elements) >= C is the pixel pitch so a
 *     tensor may be a channel slice of a wider (concat) buffer.  The residual
 *     stream and every conv result are fp32; conv OPERANDS (the activated /
 *     normalised tensors a conv reads, and the packed weights) are HL_DT_F16 or
 *     HL_DT_F32 buffers -- the rounding to the operand type is the only place
 *     precision is given up (fp32 accumulation everywhere);
 *   - process-wide settings (hl_conv_set_*, hl_set_pdl, the workspace registry, hl_launch_count) are
 *     plain globals: one host thread per process drives the library, as in the reference's
 *     one-process-per-GPU model.
 */
#ifndef HUMANLIFF_B200_H
#define HUMANLIFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HL_OK 0
#define HL_E_INVALID (-1)   /* bad argument / unsupported shape            */
#define HL_E_CUDA (-2)      /* a CUDA runtime / driver call failed          */
#define HL_E_UNSUPPORTED (-3)

/* element types of activation / weight OPERAND buffers (results and the residual stream are fp32) */
#define HL_DT_F32 0   /* fp32; the tensor-core path reads it as TF32 (caller rounds with cvt.rna)   */
#define HL_DT_F16 1   /* IEEE fp16: 11-bit significand like TF32, half the bytes, 2x the MMA rate   */

/* hl_conv2d flags */
/* Operand mode word of the kernels that WRITE conv operands (the `round_tf32` argument of hl_nchw_to_nhwc,
 * hl_cast_operand, hl_upsample2x, hl_gn_apply): */
#define HL_OP_TF32 1      /* fp32 operand: round to TF32 (cvt.rna)                                            */
#define HL_OP_SCALE 0.0625f /* = 2^-4: |x| up to 1.0e6 stays finite in fp16, and the lo half of an O(1) value is still a
                              normal fp16 number (2^-8 would push it into the subnormals: 15 instead of 22 bits)        */
#define HL_OP_SCALED 2    /* fp16 operand: store value * 2^-4 (the conv's packed weights carry 2^4): a raw
                             residual-stream operand keeps fp16 range up to 1.0e6 instead of 65504           */
#define HL_OP_SPLIT 4     /* fp16 operand: store an fp16 hi | lo pair (lo = fp16(v - hi), ~22 significant bits),
                             lo at + (mode >> 8) elements -- the operand of HL_CONV_SPLIT3 / SPLIT2P convs   */
#define HL_OP_RAW_SHIFT 4 /* hl_gn_apply: bits 4-6 = the HL_OP_* flags of the raw copy (split: lo at channel C) */
#define HL_OP_X_F16 8     /* hl_gn_apply: the INPUT x is an fp16 tensor (pitch ldx in halves) -- the result of a conv
                             run with HL_CONV_OUT_F16 whose only reader is this GroupNorm (a ResBlock's in_layers
                             output, unet.py:201-206); no raw copy in this mode                                */

#define HL_CONV_FORCE_SIMT 1   /* use the fp32 CUDA-core kernel even where the tcgen05 path applies */
#define HL_CONV_UPSAMPLE2X 2   /* input is read through a nearest x2 upsample (unet.py:77)          */
#define HL_CONV_TF32 4         /* fp32 operands may go through tcgen05 kind::tf32                   */
#define HL_CONV_OUT_F16 8      /* y is an fp16 buffer (pitch ldy in halves); statistics, if asked for, are those
                                  of the ROUNDED values (what the reader of y sees)                  */
/* High-precision operand passes of the fp16 plan (tcgen05 kernel only).  The convs that read the RAW
 * residual stream (1x1 skip, ControlNet projection, Downsample, stem) and the output conv carry each operand as an
 * fp16 hi + lo pair (~22 significant bits) and run three (two) tensor-core passes into one accumulator:          */
#define HL_CONV_SPLIT3 16      /* x = [hi(Cin) | lo(Cin)] per pixel (ldx >= 2 Cin), w = {W_hi, W_lo} slabs
                                  [2*taps][Cout_pad][Cin]:  y = x_hi.W_hi + x_lo.W_hi + x_hi.W_lo          */
#define HL_CONV_SPLIT2P 32     /* hi and lo packed INSIDE the Cin channels (stem: [hi(27) 0(5) | lo(27) 0(5)]),
                                  w = {[W_hi | W_hi], [W_lo | 0]}:  two passes                              */
#define HL_CONV_SPLIT2A 128    /* x = [hi(Cin) | lo(Cin)] as SPLIT3, ONE weight slab: y = x_hi.W + x_lo.W (the
                                  activation pair only: the output conv, whose N = 27 makes a third pass dear)  */
#define HL_CONV_OUT_F16_SPLIT 64  /* y = [hi(Cout) | lo(Cout)] fp16 of (result * 2^-4) (ldy >= 2 Cout,
                                  Cout % 32 == 0): the operand of a following HL_CONV_SPLIT3 conv           */

int hl_version(void);
const char *hl_last_error(void);
/* 1 if the tcgen05 (tensor-core) kernel would be used for this conv shape. */
int hl_conv2d_uses_tensor_cores(int x_dtype, int B, int H, int W, int Cin, int Cout, int ksize, int stride,
                                int ldx, int ldy, int flags);
/* Experiment / test hook for the tcgen05 conv: force halves per CTA (mh 1|2), the N tile, HALO
 * operand reuse (0|1), in-epilogue GroupNorm statistics (0|1), descriptor base_offset (0|1);
 * -1 = automatic.  Process-wide; not part of the reference-facing surface.                      */
int hl_conv_set_tuning(int mh, int n_tile, int halo, int epi_stats, int base_off);
/* More experiment knobs: cap on the smem pipeline depth, epilogue staging buffers per group (2|4),
 * CTA-pair MMA (cta_group::2: 0 off, 1 on where legal); -1 = automatic.                           */
int hl_conv_set_tuning2(int max_stages, int nbuf, int cta2);
/* Split-K for the low-resolution 3x3 layers (8^2, 16^2: a grid of 24-48 CTAs streaming a 100+ stage K loop):
 * with a partial-sum workspace registered for `stream` (caller-owned device memory, >= S * B*H*W*cout_pad*4
 * bytes for the layers that should split; NULL unregisters) hl_conv2d cuts K into S slices over S x the CTAs
 * and a second kernel adds the slices in a fixed order (deterministic) with bias / residual / statistics /
 * rounding.  Streams that run concurrently need distinct workspaces.  hl_conv_set_split: -1 automatic
 * (default), 0 off, n > 1 forces n slices wherever n divides the K chunk count (tests).                     */
int hl_conv_set_workspace(void *ws, int64_t bytes, void *stream);
/* Where the second pass of a split-K convolution runs: 0 (default) = the separate fixed-order reduction launch; 1 = inside
 * the conv kernel -- every slice CTA stores its partial tile and bumps the tile's counter (the last 8 KB of the registered
 * workspace, zeroed at registration and left zero by every launch), the CTA that arrives last adds the slices in slice
 * order and runs the real epilogue.  Same results bit for bit (tests compare the two); the in-kernel form measured
 * 0.9 ms/step SLOWER on B200 (one CTA per tile re-reads all partial tiles) and stays an experiment.               */
int hl_conv_set_split_reduce(int in_kernel);
/* Host-only query of the tiling hl_conv2d would use (no device work; 148 SMs assumed without a GPU).  out[16]:
 * 0 tensor-core path applies, 1 CTAs per MMA (1 | 2 = cta_group::2), 2 halves per CTA tile, 3 N tile,
 * 4 HALO operand path, 5 A slots, 6 B slots, 7 staging buffers per epilogue group, 8 TMEM accumulator stages,
 * 9 TMEM columns, 10 dynamic shared memory bytes, 11 grid, 12 tiles, 13 K slices (split-K with a workspace of
 * ws_bytes), 14 channel chunks per slice, 15 statistics in the epilogue.                                        has_res = 2 asks for the plan of the
 * dual-output launch (hl_conv2d_dual: 12 KB of shared memory reserved for the second statistics row, never split). */
int hl_conv2d_plan_info(int x_dtype, int B, int H, int W, int Cin, int Cout, int ksize, int stride, int has_res,
                        int want_stats, int64_t ws_bytes, int *out);
int hl_conv_set_split(int ksplit);

/* Experiment hook: device array of >= 16 uint64 that CTA 0 of every following tcgen05 conv adds its
 * blocked-cycle counters to (0 total, 1 A-full wait, 2 TMEM-empty wait, 3 B-full wait, 4 TMEM-full
 * wait, 5 residual wait, 6 epilogue barrier, 7 A-empty wait, 8 B-empty wait, 9 store-drain wait,
 * 10 tiles per CTA); NULL switches it off.                                                       */
int hl_conv_set_profile(void *dev_counters);

/* ---- layout / elementwise ------------------------------------------------------------------ */

/* NCHW fp32 -> NHWC (dst_dtype) with zero channel padding up to ld; optional second addend
 * (h_cond = x + x_cond, unet.py:596); for HL_DT_F32 optional TF32 round-to-nearest.              */
int hl_nchw_to_nhwc(const float *src, const float *src2 /*nullable*/, void *dst, int dst_dtype, int B,
                    int C, int HW, int ld, int round_tf32, void *stream);
/* NHWC fp32 (pitch ld) -> NCHW fp32.                                                             */
int hl_nhwc_to_nchw(const float *src, int ld, float *dst, int B, int C, int HW, void *stream);
/* dst[b, c, p] = src[b, p, c] + src[b, p, off2 + c]: the output conv of the fp16 plan carries its weights as an fp16
 * hi + lo pair stacked along Cout (rows [0, C) = W_hi, rows [off2, off2 + C) = W_lo; unet.py:475,612), so the two
 * halves of its NHWC result are summed on the way to NCHW.                                                      */
int hl_nhwc_to_nchw_sum2(const float *src, int ld, int off2, float *dst, int B, int C, int HW, void *stream);
/* nearest x2 upsample of the fp32 residual stream into an operand buffer,
 * F.interpolate(scale_factor=2, mode="nearest"), unet.py:77                                     */
int hl_upsample2x(const float *src, int lds, void *dst, int dst_dtype, int ldd, int B, int H, int W,
                  int C, int round_tf32, void *stream);
/* dst = operand copy of the fp32 tensor src: fp16 (round-to-nearest-even) or TF32-rounded fp32 --
 * staging for convs that read the raw residual stream (1x1 skips, projections, down-sampling)   */
int hl_cast_operand(const float *src, int lds, void *dst, int dst_dtype, int ldd, int C, int64_t npix,
                    int round_tf32, void *stream);

/* cudaMemsetAsync(ptr, 0, bytes): one call per step zeroes every GroupNorm statistics buffer.     */
int hl_zero(void *ptr, int64_t bytes, void *stream);

/* ---- embeddings (nn.py:103-121, unet.py:366-373,564,584-586,151-157,200) --------------------- */

/* out[b, 0:half] = cos(t_b f_k), out[b, half:] = sin(t_b f_k); freqs [dim/2] is the device copy of
 * f_k = exp(-ln(1e4) k / half), which the reference evaluates on the HOST in fp32 (nn.py:114-116).      */
int hl_timestep_embedding(const float *t, const float *freqs, int B, int dim, float *out, void *stream);
/* y[b, o] = bias[o] + sum_i W[o, i] * act(x[b, i]) (+ add[idx[b], o]);  act = SiLU if silu_in.
 * W is row-major [out, in] exactly as nn.Linear stores it.  Used for time_embed, and once per
 * step for ALL ResBlock emb_layers stacked into one [sum 2Cout, 768] matrix.                    */
int hl_linear_small(const float *x, const float *W, const float *bias, float *y, int B, int in_f,
                    int out_f, int silu_in, const float *add_table /*nullable*/,
                    const int64_t *add_idx /*nullable*/, void *stream);

/* ---- GroupNorm32 + SiLU + FiLM (nn.py:12-19,93-100; unet.py:204-206) ------------------------- */

/* Per-channel statistics: stats[(b*stats_ld + c)*2 + {0,1}] (double) += sum / sum of squares of
 * channel c of sample b over the HW pixels.  ACCUMULATES: the caller zeroes `stats` (one memset per
 * step covers every tensor).  Channel sums fold into any group layout, including groups that
 * straddle the two sources of a concatenation (unet.py:606).  hl_conv2d can produce the same
 * numbers from its epilogue.                                                                     */
int hl_gn_stats(const float *x, int ldx, int B, int HW, int C, double *stats, int stats_ld, void *stream);
/* y = act( GN(x; gamma, beta, eps) * (1 + scale_b) + shift_b ) written as an OPERAND (y_dtype),
 * scale/shift optional (film + b*film_ld points at [scale(C) | shift(C)] of sample b); act = SiLU if
 * silu; `raw` (nullable) additionally receives the plain operand copy of x (same dtype, pitch ldraw)
 * for the ResBlock's 1x1 skip convolution, saving a second pass over x.                          */
/* GroupNorm-1 + SiLU operand pass FUSED with the ResBlock's 1x1 skip convolution (unet.py:198-201,208-219): x is read
 * once (fp32, TMA); act[b, p, :Cin] = fp16(SiLU(GN(x))) (pitch ld_act halves) and skip[b, p, :Cout] = W . x + bias (fp32,
 * pitch ld_skip) come out of one kernel -- the conv on tcgen05 with the hi + lo operand pair of x * 2^-4 built in shared
 * memory (weights: the HL_CONV_SPLIT3 packing {W_hi, W_lo} of w * 2^4).  Same arithmetic as hl_gn_apply (+ raw copy)
 * followed by hl_conv2d(HL_CONV_SPLIT3): act bit-identical, skip up to fp32 accumulation order.  Half the HBM bytes.
 * hl_gn_skip_supported: H*W % 128 == 0, Cin % 64 == 0, Cout <= 256 (or 2 x 192 / 2 x 256).                            */
int hl_gn_skip_supported(int B, int HW, int Cin, int Cout);
/* diagnostics: 10 cycle counters of (CTA 0, thread 0) -- bookkeeping, wait x, act, wait A buffer, split, barrier, MMA issue,
 * wait last MMAs, epilogue, tiles -- accumulated while non-null */
int hl_gn_skip_set_profile(void *dev_counters);
int hl_gn_skip(const float *x, int ldx, const double *stats, int stats_ld, const float *gamma, const float *beta,
               void *act, int ld_act, const void *wpk, const float *bias, float *skip, int ld_skip, int B, int HW,
               int Cin, int Cout, int groups, float eps, void *stream);
/* experiment hook: blocks of hl_gn_apply per SM and launch (<= 0 restores the default) */
int hl_gn_set_tuning(int blocks_per_sm);
int hl_gn_apply(const float *x, int ldx, const double *stats, int stats_ld, const float *gamma,
                const float *beta, const float *film /*nullable*/, int film_ld, void *y, int y_dtype,
                int ldy, void *raw /*nullable*/, int ldraw, int B, int HW, int C, int groups, float eps,
                int silu, int round_tf32, void *stream);

/* ---- convolution / GEMM (unet.py:68,100,149,164-184,237-239,378,474,481-518) ------------------ */

/* y[b,oy,ox,co] = bias[co] + sum_{ky,kx,ci} x[b, oy*stride+ky-pad, ox*stride+kx-pad, ci] *
 *                 w[co,ci,ky,kx]  (+ residual[b,oy,ox,co]),   zero padding pad = ksize/2.
 * x: NHWC operand buffer of x_dtype, B x H x W x Cin with pixel pitch ldx (elements); H, W are the
 * INPUT size.  wpk is the packed weight in the SAME dtype: [ksize*ksize][Cout_pad][Cin], Cin = the
 * (padded) channel count of x -- a multiple of 64 (fp16) / 32 (tf32) for the tensor-core path --
 * Cout_pad = hl_conv_cout_pad(Cout); bias fp32 [Cout_pad].  ksize in {1,3}; stride in {1,2}.
 * y / residual: fp32 NHWC with pitches ldy / ldr (channel slices of wider buffers are fine: the
 * decoder's concat is never materialised).  stats (nullable): per-channel sum / sum-of-squares of
 * y, layout as hl_gn_stats, ACCUMULATED into the caller-zeroed buffer.  With HL_CONV_OUT_F16 the result
 * (same fp32 arithmetic, rounded once) is written as an fp16 operand buffer: used where the tensor is only
 * ever consumed as an operand (qkv -> attention, ControlNet block output -> its projection conv) or only by a
 * GroupNorm (the tensor between the two convs of a ResBlock: hl_gn_apply with HL_OP_X_F16).
 * A 1x1 conv over [B*T, C] rows is the Conv1d / GEMM of the attention block.                    */
int hl_conv_cout_pad(int Cout);
int hl_conv2d(const void *x, int x_dtype, int ldx, const void *wpk, const float *bias,
              const float *residual /*nullable*/, int ldr, void *y, int ldy, double *stats /*nullable*/,
              int stats_ld, int B, int H, int W, int Cin, int Cout, int ksize, int stride, int flags,
              void *stream);
/* Two results from ONE pass over the operands:  y = conv(x) + bias + residual  and  y2 = conv(x) + bias  (both fp32,
 * each with its own optional statistics row; stats and stats2 are given together or not at all).  Replaces the two
 * launches of the ControlNet projection: y2 = h_cond feeds the next ControlNet block (unet.py:599-601), y = hs + h_cond
 * is the decoder's skip slice (unet.py:606).  On the tcgen05 kernel one epilogue warpgroup drains every accumulator
 * chunk with the residual into y while the other writes the plain chunk into y2; other shapes run as two launches. */
int hl_conv2d_dual(const void *x, int x_dtype, int ldx, const void *wpk, const float *bias, const float *residual,
                   int ldr, float *y, int ldy, double *stats /*nullable*/, int stats_ld, float *y2, int ldy2,
                   double *stats2 /*nullable*/, int stats2_ld, int B, int H, int W, int Cin, int Cout, int ksize,
                   int stride, int flags, void *stream);

/* ---- launch mode ------------------------------------------------------------------------------ */

/* Programmatic dependent launch for every UNet-step kernel of this library (default off): with on = 1 each
 * launch carries cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel's CTAs are scheduled while
 * the previous kernel of the stream drains and block in griddepcontrol.wait until it has completed (stream
 * order semantics are unchanged; works eagerly and under stream capture).  on < 0 only queries.  Returns the
 * previous mode.  hl_pdl_barrier(): the NEXT launch is issued as a normal, fully serialized one (used after
 * cross-stream event waits).                                                                          */
int hl_set_pdl(int on);
void hl_pdl_barrier(void);
/* Kernels launched (or recorded into a stream capture) by the UNet-step entry points since load.    */
int64_t hl_launch_count(void);

/* ---- attention (unet.py:255-274) ------------------------------------------------------------ */

/* qkv: fp32 or fp16 (qkv_dtype; fp16 = written by hl_conv2d with HL_CONV_OUT_F16, tensor-core kernel only)
 * [B, T, 3C] rows (pitch ldq elements) with channel order [head][q(ch) k(ch) v(ch)];
 * out[b, t, head*ch + c] = sum_s softmax_s( q_t.k_s / sqrt(ch) ) v_s[c], written as an operand
 * (out_dtype) for the proj_out GEMM.  With an fp16 output both contractions run on the tensor
 * cores (q, k, v, softmax weights rounded to fp16; scores / softmax / accumulation fp32); an fp32
 * output selects the exact CUDA-core kernel.  fp16 qkv AND output with T % 64 == 0 and a head width of 64 / 96 /
 * 128 / 192 run on the tcgen05 kernel (S and O accumulators and the softmax weights in tensor memory, Q / K / V
 * tiles by TMA; attention_tc5.cu).  round_tf32: bit 0 = TF32-round an fp32 output, bit 1 = force the CUDA-core
 * kernel, bit 2 = force the mma.sync kernel where the tcgen05 one would run (tests).              */
int hl_attention(const void *qkv, int qkv_dtype, int ldq, void *out, int out_dtype, int ldo, int B, int T, int C,
                 int heads, int round_tf32, void *stream);

/* ---- DDPM posterior step (gaussian_diffusion.py:293-314,328-333,383-387) ---------------------- */

/* coef: device [T, 4] fp32 = {sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod,
 * posterior_mean_coef1, posterior_mean_coef2}; sigma: device [T] fp32 = exp(0.5*log_var) with the
 * t == 0 entry already zeroed (nonzero_mask).  t: device int64 [B].  n = C*H*W per sample.
 * x0 = clip(c0 x - c1 eps);  mean = c2 x0 + c3 x;  sample = mean + sigma_t * noise               */
int hl_ddpm_step(const float *x, const float *eps, const float *noise, const float *coef,
                 const float *sigma, const int64_t *t, float *sample, float *pred_xstart, int B,
                 int64_t n, int clip, void *stream);

/* DDIM update (gaussian_diffusion.py:484-529, "next" row of SURVEY 8(f)): coef [T, 4] fp32 =
 * {sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, sqrt(alpha_bar_prev), sqrt(1 - alpha_bar_prev -
 * sigma^2)}; sigma [T] = eta-dependent DDIM sigma with the t == 0 entry zeroed; noise may be NULL (eta = 0).
 * x0 = clip(c0 x - c1 eps);  eps' = (c0 x - x0) / c1;  sample = x0 ca + cb eps' + sigma_t noise           */
int hl_ddim_step(const float *x, const float *eps, const float *noise /*nullable*/, const float *coef,
                 const float *sigma, const int64_t *t, float *sample, float *pred_xstart, int B, int64_t n,
                 int clip, void *stream);

/* ---- sampling loop without library kernels (gaussian_diffusion.py:383-387,390-482; respace.py:117-122) ---- */

/* Counter-based Gaussian generator (Philox4x32-10 + Box-Muller): element e of draw number `draw` under `seed`.
 * rng_state (nullable): device uint64[2] = {seed, draw} read by the kernel instead of the by-value pair, so a
 * captured CUDA graph draws fresh noise on every replay (hl_loop_advance increments the draw counter).
 * Replaces th.randn(*shape) (gaussian_diffusion.py:460).                                                        */
int hl_randn(float *out, int64_t n, const uint64_t *rng_state, uint64_t seed, uint64_t draw,
             int64_t element_offset /* index of out[0] in the global tensor: sharding-invariant noise */, void *stream);

/* hl_ddpm_step with the per-step Gaussian drawn inside the kernel when noise == NULL (replaces the
 * th.randn_like(x) launch of gaussian_diffusion.py:383): 16 B / element of HBM traffic.  T = table length; a
 * timestep outside [0, T) fills that sample's row with NaN instead of reading out of bounds.                     */
int hl_ddpm_step_rng(const float *x, const float *eps, const float *noise /*nullable*/, const float *coef,
                     const float *sigma, const int64_t *t, int T, float *sample, float *pred_xstart /*nullable*/,
                     int B, int64_t n, int clip, const uint64_t *rng_state /*nullable*/, uint64_t seed,
                     uint64_t draw, int64_t sample_offset /* global index of sample 0: the noise of a sample does not
                     depend on how the batch is sharded across ranks */, void *stream);

/* Posterior from a caller-supplied x0 (the denoised_fn route, gaussian_diffusion.py:294-295,312-314):
 * x0c = clip(x0); sample = c2 x0c + c3 x + sigma_t * noise.                                                      */
int hl_ddpm_posterior(const float *x, const float *x0, const float *noise /*nullable*/, const float *coef,
                      const float *sigma, const int64_t *t, int T, float *sample, float *x0_clipped /*nullable*/,
                      int B, int64_t n, int clip, const uint64_t *rng_state /*nullable*/, uint64_t seed,
                      uint64_t draw, int64_t sample_offset, void *stream);

/* End of a loop iteration, on the device: t[b] -= 1; t_model[b] = scale * (timestep_map ? map[t[b]] : t[b])
 * (_WrappedModel.__call__, respace.py:117-122; scale = 1000 / T_original iff rescale_timesteps, else 1);
 * rng_state[1] += 1.  Lets UNet + posterior + advance be ONE CUDA graph replayed per step.                      */
int hl_loop_advance(int64_t *t, float *t_model, const int64_t *timestep_map /*nullable*/, float scale, int B,
                    uint64_t *rng_state /*nullable*/, void *stream);

/* ---- tri-plane volume renderer (recon_NeRF/lib/renderer.py:142-295,504-581;
 *      recon_NeRF/run_nerf_batch.py:29-67; human_diffusion/NeRF/renderer.py:234-281) ----------- */

/* Packed decoder MLP (one device buffer of HL_MLP_PACK_FLOATS floats).  Every matrix is stored
 * TRANSPOSED w.r.t. nn.Linear ([in][out], "k-major") so a thread's 8 output weights are contiguous. */
#define HL_MLP_W0 0                         /* pts_linears.0.weight^T  [27][128]                    */
#define HL_MLP_B0 (HL_MLP_W0 + 27 * 128)    /* pts_linears.0.bias      [128]                        */
#define HL_MLP_W1 (HL_MLP_B0 + 128)         /* pts_linears.1.weight^T  [128][128]                   */
#define HL_MLP_B1 (HL_MLP_W1 + 128 * 128)
#define HL_MLP_W2 (HL_MLP_B1 + 128)         /* pts_linears.2.weight^T  [155][128], in = [x | h1]    */
#define HL_MLP_B2 (HL_MLP_W2 + 155 * 128)
#define HL_MLP_WA (HL_MLP_B2 + 128)         /* alpha_linear.weight     [128]                        */
#define HL_MLP_BA (HL_MLP_WA + 128)         /* alpha_linear.bias       [1] (+3 pad)                 */
#define HL_MLP_WF (HL_MLP_BA + 4)           /* feature_linear.weight^T [128][128]                   */
#define HL_MLP_BF (HL_MLP_WF + 128 * 128)
#define HL_MLP_WV (HL_MLP_BF + 128)         /* views_linear.weight^T   [155][64], in = [feat | pe]  */
#define HL_MLP_BV (HL_MLP_WV + 155 * 64)
#define HL_MLP_WR (HL_MLP_BV + 64)          /* rgb_linear.weight^T     [64][4] (3 used)             */
#define HL_MLP_BR (HL_MLP_WR + 64 * 4)      /* rgb_linear.bias         [4] (3 used)                 */
#define HL_MLP_PACK_FLOATS (HL_MLP_BR + 4)

/* Reference tri-plane [3 planes, 9 channels, R, R] (channel = sub*3 + c) -> texel-major
 * [3 planes][3 sub-planes][R][R][4] (3 channels + 1 zero pad): one bilinear tap = one 16-byte load. */
int hl_triplane_to_texels(const float *planes, float *texels, int R, void *stream);

/* One call renders n_rays rays of one tri-plane: 128 coarse samples -> importance resampling with
 * the caller's uniforms u [n_rays, 128] (the reference draws them with torch.rand on the CPU,
 * renderer.py:563; u == NULL selects an in-kernel counter-based generator keyed by `seed`)
 * -> sort -> 256-sample fine pass -> composite.
 * rays_o / rays_d: [n_rays, 3]; near / far: [n_rays]; bounds: HOST float[6] = min xyz, max xyz.
 * z_coarse: the caller's coarse depths (Renderer.render receives them from render(),
 * run_nerf_batch.py:46-47); NULL = near*(1-t) + far*t with t = linspace(0,1,128) computed in-kernel.
 * Outputs: rgb [n_rays,3], acc [n_rays], depth [n_rays] (normal_map aliases rgb, renderer.py:237).
 * clamp_depth: 1 = human_diffusion/NeRF/renderer.py:273-274 variant, 0 = recon_NeRF variant.    */
int hl_render_rays(const float *texels, int R, const float *mlp_packed, const float *rays_o,
                   const float *rays_d, const float *near, const float *far,
                   const float *z_coarse /*nullable [n_rays,128]*/, const float *u /*nullable*/,
                   uint64_t seed, const float *bounds /*host*/, float *rgb, float *acc, float *depth,
                   int64_t n_rays, int clamp_depth, void *stream);

/* ---- canonical-space rendering (use_canonical_space=True: the TightCap branch of
 * human_diffusion/scripts/triplane_sample_layered.py:73-76) --------------------------------------------------------
 * Replaces human_diffusion/NeRF/renderer.py:52-133 (deform_target2c, deform_target2c_op; == recon_NeRF/lib/renderer.py
 * :60-140) including its pytorch3d knn_points(K=1) call, :354-420 (the per-vertex part of linear blend skinning).
 *
 * hl_smpl_vertex_tables: once per frame.  Folds everything deform_target2c_op indexes by the nearest vertex -- blended
 * joint transforms of the frame's pose and of the canonical big pose, pose / shape blend-shape offsets -- into one 3x4
 * affine per vertex, and writes the vertex positions in the SMPL frame, grouped into spatial clusters with their
 * bounding spheres, for the exact nearest-vertex search.
 *   weights [V,J], posedirs [V,3,9(J-1)], shapedirs [V,3,n_betas_asset]: device fp32 tables of the SMPL asset (J = 24
 *   joints for SMPL, 55 for SMPL-X); vertices [V,3]: tp_input['vertices'] (world space, device fp32)
 *   consts: device fp64[HL_SMPL_CONSTS(J)]: A_pose[J][3][4] | A_big[J][3][4] (get_transform_params_torch of params and of
 *           t_params with zero shape, rows of the 4x4) | pose_feature[9(J-1)] | pose_feature_big[9(J-1)] | betas[16] |
 *           R[9] | Th[3]
 *   slot_vertex: device int32 [n_clusters * cluster_slots]: the vertex of every table slot, -1 = unused (clusters are a
 *           property of the asset: any partition is correct, a spatially compact one is fast)
 *   knn_table: float4 [n_clusters] bounding spheres | float4 [n_clusters * cluster_slots] vertices, both two entries per
 *           pair of float4 -- {x0, x1, y0, y1} {z0, z1, w0, w1}, w = radius | vertex index -- the operand layout of the packed
 *           fp32 instructions, which serve bounds and the scan filter only: decisions are taken on separately rounded
 *           distances (n_clusters even, <= 128; cluster_slots a multiple of 4, <= 96);
 *           affine_table: [V][3][4] floats (rows M | c)                                                             */
#define HL_SMPL_CONSTS(J) (24 * (J) + 18 * ((J) - 1) + 28)
int hl_smpl_vertex_tables(const float *weights, const float *posedirs, const float *shapedirs, int n_betas_asset,
                          int n_betas, const float *vertices, const double *consts, int n_verts, int n_joints,
                          const int *slot_vertex, int n_clusters, int cluster_slots, float *knn_table,
                          float *affine_table, void *stream);

/* hl_render_rays with every sample point (and, in the fine pass, its view direction) deformed to the canonical space:
 * q = (p - trans) rot; nearest vertex of q (exact; lowest index on ties); p_canonical = M q + c; viewdir_canonical =
 * M ((viewdir - trans) rot) (the reference subtracts Th from the direction as well, renderer.py:125).  t_bounds =
 * tp_input['t_world_bounds'] (HOST float[6]); rot = params['R'] row-major (HOST float[9]); trans = params['Th'] (HOST).
 * Exact fp32 MLP on the CUDA cores. */
int hl_render_rays_canon(const float *texels, int R, const float *mlp_packed, const float *rays_o,
                         const float *rays_d, const float *near, const float *far,
                         const float *z_coarse /*nullable*/, const float *u /*nullable*/, uint64_t seed,
                         const float *t_bounds /*host*/, const float *knn_table, const float *affine_table,
                         int n_clusters, int cluster_slots, const float *rot /*host*/, const float *trans /*host*/,
                         float *rgb, float *acc, float *depth, int64_t n_rays, int clamp_depth, void *stream);

/* The same on the tcgen05 render kernel (hl_render_rays_tc5: fp16 operands, activations in tensor memory): the sample's
 * encoded canonical direction is written per thread into its row of the constant tile.  n_importance: 128 or 0. */
int hl_render_rays_tc5_canon(const void *quads, int R, const void *mlp_tc5, const float *rays_o, const float *rays_d,
                             const float *near, const float *far, const float *z_coarse /*nullable*/,
                             const float *u /*nullable*/, uint64_t seed, const float *t_bounds /*host*/,
                             const float *knn_table, const float *affine_table, int n_clusters, int cluster_slots,
                             const float *rot /*host*/, const float *trans /*host*/, float *rgb, float *acc, float *depth,
                             int64_t n_rays, int n_importance, int clamp_depth, void *stream);

/* deform_target2c on caller-supplied points [n,3] (and optionally directions [n,3]): Renderer.deform_target2c and the
 * parity tests. */
int hl_canonical_points(const float *pts, const float *dirs /*nullable*/, int64_t n, const float *knn_table,
                        const float *affine_table, int n_clusters, int cluster_slots, const float *rot /*host*/,
                        const float *trans /*host*/, float *out_pts, float *out_dirs /*nullable*/, void *stream);

/* extract_geometry's field with use_canonical_space=True (human_diffusion/NeRF/renderer.py:290-318): out[xi][yi][zi] =
 * -sigma at linspace(world_bounds)^3, each grid point deformed to the canonical space first and looked up inside
 * t_bounds.  world_bounds / t_bounds / rot / trans: HOST arrays. */
int hl_density_grid_canon(const float *texels, int R, const float *mlp_packed, const float *world_bounds /*host[6]*/,
                          const float *t_bounds /*host[6]*/, const float *knn_table, const float *affine_table,
                          int n_clusters, int cluster_slots, const float *rot /*host*/, const float *trans /*host*/,
                          int resolution, float *out, void *stream);

/* Tensor-core variant of hl_render_rays: every 128-point layer of the decoder MLP runs on mma.sync
 * (fp16 operands, fp32 accumulate; activations stay in registers between layers).  mlp_f16 is the fp16
 * weight image below (HL_MLP16_HALVES halves; rows = output features, row pitch = K + 8 halves so that
 * ldmatrix is bank-conflict free; unused k columns zero); mlp_packed (the fp32 pack above) still supplies
 * biases, the alpha / rgb heads and the view-direction rows of views_linear.  Same outputs as
 * hl_render_rays to ~1e-5 relative (operand rounding averaged over 256 samples per ray).
 *   HL_MLP16_W0: pts_linears.0  [128][40]   k = x(27) | 0(5)
 *   HL_MLP16_W1: pts_linears.1  [128][136]
 *   HL_MLP16_W2: pts_linears.2  [128][168]  k = x(27) | 0(5) | h1(128)
 *   HL_MLP16_WF: feature_linear [128][136]
 *   HL_MLP16_WV: views_linear   [64][136]   k = feature(128); pe(d) columns are folded into a per-ray bias */
#define HL_MLP16_W0 0
#define HL_MLP16_W1 (HL_MLP16_W0 + 128 * 40)
#define HL_MLP16_W2 (HL_MLP16_W1 + 128 * 136)
#define HL_MLP16_WF (HL_MLP16_W2 + 128 * 168)
#define HL_MLP16_WV (HL_MLP16_WF + 128 * 136)
#define HL_MLP16_HALVES (HL_MLP16_WV + 64 * 136)
int hl_render_rays_tc(const float *texels, int R, const float *mlp_packed, const void *mlp_f16,
                      const float *rays_o, const float *rays_d, const float *near, const float *far,
                      const float *z_coarse /*nullable*/, const float *u /*nullable*/, uint64_t seed,
                      const float *bounds /*host*/, float *rgb, float *acc, float *depth, int64_t n_rays,
                      int clamp_depth, void *stream);

/* Density on a regular grid -- the GPU part of Renderer.extract_geometry ("next" row, SURVEY 8(f) rank 1;
 * human_diffusion/NeRF/renderer.py:290-318): out[xi][yi][zi] = -sigma(p), p = (linspace(min_x, max_x, res)[xi],
 * ...), the coarse stage of the renderer (nine-plane gather + density MLP on the tensor cores) over res^3
 * points.  Marching cubes stays on the host (mcubes in the reference).                              */
/* Experiment hook: device array of >= 8 uint64; CTA 0 of following hl_render_rays_tc launches adds the cycles it
 * spends per phase (0 per-ray setup, 1 gather, 2 MLP, 3 resample + sort, 4 composite, 5 total); NULL = off. */
int hl_render_set_profile(void *dev_counters);
int hl_density_grid_tc(const float *texels, int R, const float *mlp_packed, const void *mlp_f16,
                       const float *bounds /*host*/, int resolution, float *out, void *stream);

/* ---- tcgen05 / TMEM renderer (the default of precision="fp16") ---------------------------------------------
 * Same chain as hl_render_rays_tc, but every MLP layer is a tcgen05.mma (M = 128 samples, accumulators AND
 * activations in tensor memory, weights resident in shared memory); two ray groups per SM.
 * mlp_tc5: HL_MLP_TC5_BYTES bytes.  First the fp16 operands as K-major SWIZZLE_128B atoms [rows][64 halves] (the
 * 16-byte chunk c of row r stored at chunk c ^ (r & 7)); L2E = log2(e), LN2 = ln 2 (the softplus runs in the log2
 * domain: the factor L2E is folded into the weights that feed a softplus, LN2 into the weights that read one):
 *   pts_linears.0 * L2E          128 x 64   k 0..26 = weights, k 27 = bias * L2E (x carries a 1.0 in slot 27)
 *   pts_linears.1                128 x 128  (2 atoms)
 *   pts_linears.2 x part * L2E   128 x 64   k 27 = bias * L2E
 *   pts_linears.2 h1 part        128 x 128  (2 atoms)
 *   feature_linear * LN2         128 x 128  (2 atoms)
 *   views_linear[:, :128] * L2E   64 x 128  (2 atoms of 64 rows)
 *   bias atom                    128 x 64   k 0 = pts_linears.1 bias * L2E, k 16 = feature_linear bias
 *   views_linear[:, 128:] * L2E   64 x 64   k 0 = views bias * L2E, k 1..27 = view-direction columns
 * then 392 floats: alpha_linear.weight * LN2 [128], alpha bias [1] + 3 pad, rgb_linear.weight^T * LN2 [64][4],
 * rgb bias [3] + 1 pad.
 * bounds: 6 floats {min xyz, max xyz}, host memory, or device memory when bounds_on_device != 0 (no host sync on
 * tp_input['world_bounds']).  n_importance: 128, or 0 = no coarse pass: the n_samples = 128 coarse depths are
 * composited directly (recon_NeRF/lib/renderer.py:258 `if n_importance > 0`).                                  */
/* texels of the tcgen05 renderer: "quad texels" -- [9 sub-planes][R + 1][R + 1] entries of 32 bytes; entry (yq, xq)
 * holds the 2 x 2 bilinear footprint whose top-left tap is (yq - 1, xq - 1): 4 taps x 3 channels as fp16 (+ 4 pad),
 * out-of-range taps = 0 (grid_sample's zero padding), so one sub-plane of one sample point is ONE 32-byte load.   */
int hl_triplane_to_quads(const float *planes /*[3][9][R][R]*/, void *quads, int R, void *stream);
#define HL_MLP_TC5_BYTES (16384 + 32768 + 16384 + 32768 + 32768 + 16384 + 16384 + 8192 + 392 * 4)
int hl_render_rays_tc5(const void *quads, int R, const void *mlp_tc5,
                       const float *rays_o, const float *rays_d, const float *near, const float *far,
                       const float *z_coarse /*nullable*/, const float *u /*nullable*/, uint64_t seed,
                       const float *bounds, int bounds_on_device, float *rgb, float *acc, float *depth,
                       int64_t n_rays, int n_importance, int clamp_depth, void *stream);
int hl_density_grid_tc5(const void *quads, int R, const void *mlp_tc5,
                        const float *bounds, int bounds_on_device, int resolution, float *out, void *stream);
/* per-phase cycle counters of (CTA 0, group 0) of following tc5 launches, as hl_render_set_profile: 8 x uint64; a
 * canonical-space launch (hl_render_rays_tc5_canon) also fills slots 8..15 with the counters of its nearest-vertex search
 * (tools/canon_probe.py), so the buffer must then hold 16 */
int hl_render5_set_profile(void *dev_counters);


#ifdef __cplusplus
}
#endif
#endif /* HUMANLIFF_B200_H */
