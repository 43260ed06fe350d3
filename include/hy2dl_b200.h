/*
 * hy2dl_b200.h -- C ABI of the B200-native Hy2DL hot path (libhy2dl_b200.so, sm_100a).
 *
 * The reference (KIT-HYD/Hy2DL) is pure Python: it has no FFI / plugin layer, its "operator API"
 * for this path is the nn.Module surface of modelzoo/ and the loss functions of aux_functions/.
 * Each entry point below replaces the arithmetic behind one of those call sites (cited per
 * function, paths relative to /root/reference/Hy2DL); the Python classes in hy2dl_b200/ keep the
 * reference's class names, constructor arguments, state-dict keys and return dicts and reach
 * this library through ctypes (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, int64 sizes, a CUDA stream handle; no C++/torch types.
 *   - every function is an asynchronous launch on `stream`; it never allocates, frees or
 *     synchronises.  Scratch space is passed in (`ws`, size from hy2dl_workspace_bytes()).
 *   - returns 0 on success, a negative HY2DL_ERR_* otherwise; hy2dl_last_error() gives a
 *     thread-local message.  There is no CPU fallback anywhere.
 *   - all floating-point buffers are fp32 and 16-byte aligned.
 *
 * Device layouts ("native", time-major so one step of one batch tile is one contiguous run)
 *   xT      [T][B][IP]      LSTM inputs, IP = I rounded up to a multiple of 8, zero padded
 *   h_seq   [T][B][H]       hidden state tape          c_seq [T][B][H]  cell state tape
 *   gates   [T][B][4][H]    activated gates i,f,g,o (tape); overwritten in place by dG in bwd
 *   forc    [T][3][B]       conceptual forcings P, PET, Tmean
 *   With HY2DL_PREC_TF32X3 the LSTM streams (xT, h_seq, c_seq, gates, dG, dh_seq) use the
 *   row-interleaved layout RI32 instead: [T][ceil(B/32)][F/32][32 rows][32 floats], and in row r the
 *   8-float unit u (= (feature % 32) / 8) is stored at unit position u ^ (r & 3).  A (step, group,
 *   32-feature block) slab is 4 KB contiguous and is exactly the MN-major SWIZZLE_128B_BASE32B UMMA
 *   operand tile of the weight-gradient GEMM, while the row-owning threads of the recurrent kernels
 *   move whole 32-byte sectors / 128-byte lines.  F is a multiple of 32: xT is 32 columns wide
 *   (IP = 32, I <= 31: column 31 carries the bias inside the forward kernel).  gates/dG order their 4H columns as n = 4*unit + gate.
 *   N = B*m series (sample b, member j -> n = b*m + j)
 *   par     planes: dynamic parameter -> [T_sim][N], static parameter -> [N]   (see hy2dl_parmap)
 *   states  [T][5][N]       bucket storages after each step
 *   q       [T][B]          member-mean outflow
 */
#ifndef HY2DL_B200_H
#define HY2DL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HY2DL_ABI_VERSION 1

#define HY2DL_OK 0
#define HY2DL_ERR_SHAPE (-1)   /* unsupported or inconsistent sizes            */
#define HY2DL_ERR_ALIGN (-2)   /* pointer not 16-byte aligned / null           */
#define HY2DL_ERR_ARCH (-3)    /* device is not sm_100                         */
#define HY2DL_ERR_LAUNCH (-4)  /* CUDA launch failed (message has the reason)  */
#define HY2DL_ERR_WORKSPACE (-5) /* workspace too small                        */

typedef void* hy2dl_stream_t; /* cudaStream_t */

/* precision of the recurrent / projection matmuls */
#define HY2DL_PREC_FP32 0   /* fp32-class (parity path, rtol 1e-4), row-major streams.  H = 128 / 256: tcgen05 kind::tf32 with the
                             * 3-term operand split -- weights streamed from L2, or (batches up to ~10 k) the unit-split kernels
                             * with resident weight slices; RB8 tapes.  H = 64: the unit-split kernels up to one round of groups
                             * (4 736 sequences on 148 SMs; RB8 tapes), CUDA-core FMA kernels above (row-major tapes).
                             * HY2DL_FP32_TC=0 in the environment selects the CUDA-core kernels everywhere */
#define HY2DL_PREC_TF32X3 1 /* fp32-class: tcgen05 kind::tf32, 3-term split, weights resident on chip (H = 64, RI32 streams) */
#define HY2DL_PREC_TF32 2   /* throughput mode: the same tensor-core kernels with the correction terms switched off (single-pass
                             * TF32 operands, 10-bit mantissa, fp32 accumulate).  NOT the parity path: measured against fp64,
                             * discharge ~1e-3 and gradients ~1e-2 of their scale (tests/test_gpu_parity.py::test_tf32_mode).  Streams
                             * and tapes as TF32X3 at H = 64 and as FP32 at H = 128 / 256 */

/* layout of the LSTM-side streams handled by the packing / sampler / head functions */
#define HY2DL_LAYOUT_ROWMAJOR 0 /* [T][B][F]                                                      */
#define HY2DL_LAYOUT_RI32 1     /* [T][ceil(B/32)][F/32][32][32] swizzled, the layout of HY2DL_PREC_TF32X3 */

/* conceptual models */
#define HY2DL_MODEL_SHM 0 /* 8 parameters: dd f_thr sumax beta perc kf ki kb     (modelzoo/shm.py:193-204) */
#define HY2DL_MODEL_HBV 1 /* 13 parameters: BETA FC K0 K1 K2 LP PERC UZL TT CFMAX CFR CWH BETAET (hbv.py:199-215) */
#define HY2DL_MODEL_LINRES 2   /* 2 parameters: ki aux_ET; 1 storage si            (modelzoo/linear_reservoir.py:99-108) */
#define HY2DL_MODEL_NONSENSE 3 /* 5 parameters: dd sumax beta ki kb; 4 storages ss su si sb (modelzoo/nonsense.py:170-176) */
#define HY2DL_MAX_PAR 16
#define HY2DL_N_STATES 5 /* planes of the states / init buffers; a model with fewer storages uses planes 0..NS-1 and
                            never reads or writes the others */

/* workspace kinds for hy2dl_workspace_bytes */
#define HY2DL_WS_LSTM_FWD 0
#define HY2DL_WS_LSTM_BWD 1
#define HY2DL_WS_LSTM_WGRAD 2

int hy2dl_abi_version(void);
const char* hy2dl_last_error(void);
/* 0 if device `dev` is compute capability 10.x, HY2DL_ERR_ARCH otherwise */
int hy2dl_check_device(int dev);
int64_t hy2dl_workspace_bytes(int kind, int64_t B, int64_t T, int64_t I, int64_t H, int64_t precision);

/* ---- layout packing --------------------------------------------------------------------------
 * x [B][T][I] (reference batch-first layout, modelzoo/cudalstm.py:34, hybrid.py:72) -> xT [T][B][IP].
 * xc [B][T][nc], nc = 3 (P, PET, T), 4 (P, PET, Tmax, Tmin -> mean; shm.py:74-80, hbv.py:70-77) or 2 (P, PET;
 *    temperature plane zero -- linear_reservoir.py:79-80 reads nothing else)  -> forc [T][3][B].
 * layout = HY2DL_LAYOUT_RI32: xT is written as RI32 (rows >= B of the last group zero filled). */
int hy2dl_pack_x(const float* x, int64_t B, int64_t T, int64_t I, int64_t IP, float* xT, int layout, hy2dl_stream_t stream);
int hy2dl_pack_forcing(const float* xc, int64_t B, int64_t T, int64_t nc, float* forc, hy2dl_stream_t stream);

/* ---- head: Linear(H, C) on the needed steps + sigmoid range mapping ---------------------------
 * replaces modelzoo/hybrid.py:93-97,110-113 and baseconceptualmodel.py:53-76.
 * `layout` is the layout of h_seq (fwd) and of h_seq + dh_seq (bwd); RI32 dh_seq rows >= B are zero.
 * Column c = p*m + j for conceptual parameter p, member j; routing parameters (m = 1) follow. */
typedef struct {
  int32_t n_par;                    /* conceptual parameters P (8 SHM, 13 HBV)                   */
  int32_t n_members;                /* m                                                          */
  int32_t n_routing;                /* 0 or 2 (alpha, beta), static, taken at t = T-1             */
  int32_t warmup;                   /* warm-up steps, >= 1 (hybrid.py:39)                         */
  int32_t dynamic[HY2DL_MAX_PAR];   /* 1 = dynamic, 0 = static (baseconceptualmodel.py:80-104)    */
  float lo[HY2DL_MAX_PAR + 2];      /* range low per parameter, then routing                      */
  float hi[HY2DL_MAX_PAR + 2];      /* range high                                                 */
  int64_t sim_off[HY2DL_MAX_PAR + 2]; /* float offset of each parameter's plane in par_sim        */
} hy2dl_parmap;
/* Optional fusion of the head into hy2dl_lstm_fwd (precision HY2DL_PREC_TF32X3, at most 32 head columns):
 * the forward kernel evaluates Linear(H, C) with its own tensor-core accumulation on the steps the
 * conceptual model needs and writes par_sim / par_warm exactly like hy2dl_head_map_fwd. */
typedef struct {
  const float* w;          /* [C][H] */
  const float* bias;       /* [C]    */
  const hy2dl_parmap* pm;  /* layout filled by hy2dl_parmap_layout */
  float* par_sim;
  float* par_warm;
} hy2dl_head_fwd;
/* Optional fusion of the head adjoint into hy2dl_lstm_bwd (same conditions): the external gradient of the
 * hidden state is dz * W_head with dz formed in the kernel from dpar_sim (gradient of par_sim) and par_sim,
 * added to the recurrent term by the same tensor-core accumulation -- no dh_seq buffer exists.
 * dz (RI32, F = 32: [T][ceil(B/32)][32][32], steps >= warmup written, columns >= C untouched) feeds
 * hy2dl_lstm_wgrad's head part.  It may be NULL, in which case the head's weight and bias gradients come from
 * hy2dl_head_map_bwd with dh_seq = NULL. */
typedef struct {
  const float* w;          /* [C][H] */
  const hy2dl_parmap* pm;
  const float* par_sim;
  const float* dpar_sim;
  float* dz;               /* out, nullable */
} hy2dl_head_bwd;
/* Optional head part of hy2dl_lstm_wgrad: dw [C][H] = sum_t dz_t^T h_t and db [C] = sum_t dz_t over the steps
 * t >= t0, on the h slabs the kernel loads anyway. */
typedef struct {
  const float* dz;         /* as written by hy2dl_lstm_bwd */
  int64_t t0;              /* first step with a head gradient (= warmup) */
  int32_t n_columns;       /* C <= 32 */
  float* dw;               /* [C][H] out */
  float* db;               /* [C] out, nullable */
} hy2dl_head_wgrad;

/* ---- LSTM (replaces torch.nn.LSTM at modelzoo/cudalstm.py:50 and hybrid.py:92; one layer,
 * unidirectional, h0 = c0 = 0, gate order i|f|g|o as torch/nn/modules/rnn.py:935-941) -------------
 * Weights come in PyTorch layout: w_ih [4H][I], w_hh [4H][H], b_ih [4H], b_hh [4H].
 * Tape outputs may be NULL (inference): c_seq, gates; h_seq may be NULL if only h_last is wanted.
 * c_seq, gates and dG are tapes that only hy2dl_lstm_bwd / hy2dl_lstm_wgrad read back: their layout is the forward
 * kernel's choice and the same for the three calls of one (B, T, H, precision).  Row-major precisions: ALWAYS allocate
 * T * ceil(B/32)*32 * F floats for them (F = H or 4H): the tensor-core kernels (H = 128 / 256, and H = 64 at small batches,
 * see HY2DL_PREC_FP32) store them as "RB8" = [T][ceil(B/32)][F/8][32 sequences][8 floats]; the CUDA-core kernels as
 * [T][B][H] and [T][B][4][H] in the first T*B*F floats. */
int hy2dl_lstm_fwd(const float* xT, int64_t B, int64_t T, int64_t I, int64_t IP, int64_t H,
                   const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                   float* h_seq, float* c_seq, float* gates, float* h_last,
                   void* ws, int64_t ws_bytes, int precision, const hy2dl_head_fwd* head, hy2dl_stream_t stream);
/* Back-propagation through time. External gradient: dh_seq [T][B][H] used for steps t >= dh_t0
 * (NULL: none) plus dh_last [B][H] added at t = T-1 (NULL: none).  dG [T][B][4][H] receives the
 * pre-activation gate gradients and MAY alias `gates`. */
int hy2dl_lstm_bwd(int64_t B, int64_t T, int64_t H, const float* w_hh,
                   const float* gates, const float* c_seq, const float* dh_seq, int64_t dh_t0,
                   const float* dh_last, float* dG,
                   void* ws, int64_t ws_bytes, int precision, const hy2dl_head_bwd* head, hy2dl_stream_t stream);
/* Weight gradients from dG: dw_ih [4H][I] = dG^T xT, dw_hh [4H][H] = dG^T h_{t-1}, db [4H] = sum dG
 * (db is the gradient of both b_ih and b_hh).  Outputs are overwritten, not accumulated. */
int hy2dl_lstm_wgrad(const float* dG, const float* h_seq, const float* xT,
                     int64_t B, int64_t T, int64_t I, int64_t IP, int64_t H,
                     float* dw_ih, float* dw_hh, float* db,
                     void* ws, int64_t ws_bytes, int precision, const hy2dl_head_wgrad* head, hy2dl_stream_t stream);

/* ---- head kernels (unfused; any H, any number of columns) --------------------------------------- */
/* fills sim_off for (B, T) and returns the number of floats par_sim needs; par_warm needs
 * n_par * B * m floats (plane p at p*B*m). */
int64_t hy2dl_parmap_layout(hy2dl_parmap* pm, int64_t B, int64_t T);
int hy2dl_head_map_fwd(const float* h_seq, int64_t B, int64_t T, int64_t H,
                       const float* w, const float* bias, const hy2dl_parmap* pm,
                       float* par_sim, float* par_warm, int layout, hy2dl_stream_t stream);
/* par_sim: the forward output; dpar_sim has the same layout.  Writes dh_seq rows t >= warmup
 * (the warm-up runs under no_grad, hybrid.py:99-102, so row warmup-1 and earlier get no gradient
 * and are left untouched), and overwrites dw [C][H], dbias [C].  dh_seq may be NULL: weight and bias
 * gradients only (the hidden-state gradient is then formed inside hy2dl_lstm_bwd, see hy2dl_head_bwd). */
/* ws: hy2dl_head_bwd_workspace_bytes(H) bytes of device scratch: per-thread-block partial sums of dw / dbias, added in
 * block order by a finish kernel (no floating-point atomics: bit-reproducible). */
int64_t hy2dl_head_bwd_workspace_bytes(int64_t H);
int hy2dl_head_map_bwd(const float* h_seq, const float* par_sim, const float* dpar_sim, int64_t B, int64_t T,
                       int64_t H, const float* w, const hy2dl_parmap* pm, float* dh_seq, float* dw, float* dbias,
                       void* ws, int64_t ws_bytes, int layout, hy2dl_stream_t stream);

/* ---- conceptual models (replace SHM.forward shm.py:135-182, HBV.forward hbv.py:127-188) --------
 * par[p] / tstride[p]: device pointer of parameter p's plane and its stride in floats between
 * time steps (N for dynamic, 0 for static).  HBV: par[12] (BETAET) may be NULL (hbv.py:157-162).
 * init [5][N] or NULL (model defaults: 0.001; NonSense ss 0, su 5, si 10, sb 15).  Steps simulated: forc rows t0 .. t0+Tn-1. */
int hy2dl_conceptual_fwd(int model, const float* forc, int64_t B, int64_t t0, int64_t Tn, int64_t m,
                         const float* const* par, const int64_t* tstride, const float* init,
                         float* states, float* q, hy2dl_stream_t stream);
/* Adjoint. dq [Tn][B] (gradient of q), dstates [Tn][5][N] or NULL.  dpar[p] has par[p]'s shape
 * (overwritten); dinit [5][N] or NULL (planes beyond the model's storages are left untouched). */
int hy2dl_conceptual_bwd(int model, const float* forc, int64_t B, int64_t t0, int64_t Tn, int64_t m,
                         const float* const* par, const int64_t* tstride, const float* init,
                         const float* states, const float* dq, const float* dstates,
                         float* const* dpar, float* dinit, hy2dl_stream_t stream);

/* ---- unit-hydrograph routing (replaces UH_routing.forward, modelzoo/uh_routing.py:42-44,64-74,
 * 92-109): alpha, beta [B]; q, y [T][B]; uh [15][B] scratch/tape ---------------------------------- */
int hy2dl_uh_fwd(const float* q, const float* alpha, const float* beta, int64_t B, int64_t T,
                 float* uh, float* y, hy2dl_stream_t stream);
int hy2dl_uh_bwd(const float* q, const float* alpha, const float* beta, const float* uh, const float* dy,
                 int64_t B, int64_t T, float* dq, float* dalpha, float* dbeta, hy2dl_stream_t stream);

/* ---- losses (replace aux_functions/functions_training.py:32-41 and :72-81) ----------------------
 * y_sim, y_obs, basin_std: n elements in the same order; NaN in y_obs = missing.
 * *_reduce writes the masked sums to acc[0..3] (NSE: {sum w*e^2, count}; wRMSE: {sum e^2, sum le^2, count});
 * acc holds hy2dl_loss_acc_floats() floats (the sums, an arrival ticket and one partial sum per thread block:
 * the reduction uses no floating-point atomics, the last block adds the partials in a fixed order, so results
 * are bit-reproducible); it must be zeroed by the caller ONCE (the kernel leaves the ticket at zero) and
 * acc[0..3] can be all-reduced across ranks before *_finish.
 * *_finish writes the scalar loss and, if dy != NULL, dL/dy_sim scaled by *gscale (device scalar
 * or NULL = 1). */
int64_t hy2dl_loss_acc_floats(void);
int hy2dl_loss_nse_reduce(const float* y_sim, const float* y_obs, const float* basin_std, int64_t n,
                          float* acc, hy2dl_stream_t stream);
int hy2dl_loss_nse_finish(const float* y_sim, const float* y_obs, const float* basin_std, int64_t n,
                          const float* acc, const float* gscale, float* loss, float* dy, hy2dl_stream_t stream);
int hy2dl_loss_wrmse_reduce(const float* y_sim, const float* y_obs, int64_t n, float* acc, hy2dl_stream_t stream);
int hy2dl_loss_wrmse_finish(const float* y_sim, const float* y_obs, int64_t n, const float* acc,
                            const float* gscale, float* loss, float* dy, hy2dl_stream_t stream);

/* ---- evaluation metric on the device (replaces the per-basin loop of aux_functions/functions_evaluation.py:28-48
 * for series that are already resident): y_sim, y_obs time-major [T][B], NaN in either = pair skipped;
 * nse [B] float64, NaN where fewer than two valid pairs remain. */
int hy2dl_nse_basins(const float* y_sim, const float* y_obs, int64_t T, int64_t B, double* nse, hy2dl_stream_t stream);

/* ---- GPU-resident sliding-window sampler (replaces BaseDataset.__getitem__ + collate,
 * datasetzoo/basedataset.py:168-195).  Resident series are the concatenation of all basins:
 * x_d [R][nd], x_conc [R][nc] (or NULL), y_obs [R], x_s [nb][ns] (or NULL), basin_std [nb],
 * row0 [nb].  Sample k = (basin[k], end[k]).  Outputs in native layouts:
 * xT [L][B][IP] (or RI32 per `layout`), forc [L][3][B] (or NULL), y_win [n_last][B], std_win [n_last][B]. */
int hy2dl_window_gather(const float* x_d, const float* x_s, const float* x_conc, const float* y_obs,
                        const float* basin_std, const int64_t* row0, const int64_t* basin, const int64_t* end,
                        int64_t B, int64_t L, int64_t n_last, int64_t nd, int64_t ns, int64_t nc, int64_t IP,
                        float* xT, float* forc, float* y_win, float* std_win, int layout, hy2dl_stream_t stream);

/* ---- optimiser tail on one flat buffer (replaces clip_grad_norm_(.., max_norm) + Adam.step,
 * experiments/LSTM_CAMELS_GB.ipynb:298-299).  norm_sq: hy2dl_sqnorm_floats() device floats, zeroed by the
 * caller once: [0] receives the squared norm (overwritten by every hy2dl_grad_sqnorm; can be all-reduced), the
 * rest is the arrival ticket and one partial sum per thread block (fixed-order sum, no floating-point atomics:
 * bit-reproducible);  lr: device scalar (so a scheduler can change it under a
 * captured CUDA graph);  step_count: device int64, incremented by the kernel;  the gradient is
 * first multiplied by grad_scale (e.g. 1/world_size), then clipped to max_norm (<= 0: no clipping). */
int64_t hy2dl_sqnorm_floats(void);
int hy2dl_grad_sqnorm(const float* grad, int64_t n, float* norm_sq, hy2dl_stream_t stream);
int hy2dl_clip_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                    const float* norm_sq, float grad_scale, float max_norm, const float* lr, float beta1,
                    float beta2, float eps, int64_t* step_count, hy2dl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HY2DL_B200_H */
