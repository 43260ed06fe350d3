// extern "C" entry points of the LSTM kernels: argument checks + precision dispatch.
#include "common.cuh"

namespace hy2dl {
int lstm_fp32_fwd(const float* xT, int64_t B, int64_t T, int64_t I, int64_t IP, int64_t H, const float* w_ih,
                  const float* w_hh, const float* b_ih, const float* b_hh, float* h_seq, float* c_seq, float* gates,
                  float* h_last, void* ws, int64_t ws_bytes, int single, cudaStream_t s);
int lstm_fp32_bwd(int64_t B, int64_t T, int64_t H, const float* w_hh, const float* gates, const float* c_seq,
                  const float* dh_seq, int64_t dh_t0, const float* dh_last, float* dG, void* ws, int64_t ws_bytes, int single,
                  cudaStream_t s);
int lstm_fp32_wgrad(const float* dG, const float* h_seq, const float* xT, int64_t B, int64_t T, int64_t I, int64_t IP,
                    int64_t H, float* dw_ih, float* dw_hh, float* db, void* ws, int64_t ws_bytes, int single, cudaStream_t s);
int64_t lstm_fp32_workspace(int kind, int64_t B, int64_t T, int64_t I, int64_t H);
namespace tc {
int lstm_tc_fwd(const float* x_ri, int64_t B, int64_t T, int64_t I, int64_t IP, const float* w_ih, const float* w_hh,
                const float* b_ih, const float* b_hh, float* h_seq, float* c_seq, float* gates, float* h_last, void* ws,
                int64_t ws_bytes, const hy2dl_head_fwd* head, int single, cudaStream_t s);
int lstm_tc_bwd(int64_t B, int64_t T, const float* w_hh, const float* gates, const float* c_seq, const float* dh_seq,
                int64_t dh_t0, const float* dh_last, float* dG, void* ws, int64_t ws_bytes, const hy2dl_head_bwd* head,
                int single, cudaStream_t s);
int lstm_tc_wgrad(const float* dG, const float* h_seq, const float* x_ri, int64_t B, int64_t T, int64_t I, int64_t IP,
                  float* dw_ih, float* dw_hh, float* db, void* ws, int64_t ws_bytes, const hy2dl_head_wgrad* head,
                  int single, cudaStream_t s);
int64_t lstm_tc_workspace(int kind, int64_t B, int64_t T, int64_t I, int64_t H);
bool lstm_tc_supported(int64_t I, int64_t H);
}  // namespace tc

// HY2DL_PREC_TF32 (single pass) runs on the kernels of TF32X3 at H = 64 (RI32 streams) and on those of FP32 at H = 128 / 256
// (row-major streams, RB8 tapes), with the correction MMAs switched off.
static bool ri32_family(int precision, int64_t H) { return precision == HY2DL_PREC_TF32X3 || (precision == HY2DL_PREC_TF32 && H == 64); }
static bool known_precision(int p) { return p == HY2DL_PREC_FP32 || p == HY2DL_PREC_TF32X3 || p == HY2DL_PREC_TF32; }

static int check_lstm_shape(const char* who, int64_t B, int64_t T, int64_t I, int64_t IP, int64_t H, int precision) {
  HY_REQUIRE(B > 0 && T > 0, HY2DL_ERR_SHAPE, "%s: empty batch or sequence (B=%lld T=%lld)", who, (long long)B, (long long)T);
  HY_REQUIRE(H == 64 || H == 128 || H == 256, HY2DL_ERR_SHAPE, "%s: hidden_size=%lld not supported (64, 128, 256)", who, (long long)H);
  HY_REQUIRE(known_precision(precision), HY2DL_ERR_SHAPE, "%s: unknown precision %d", who, precision);
  HY_REQUIRE(precision != HY2DL_PREC_TF32X3 || H == 64, HY2DL_ERR_SHAPE, "%s: precision tf32x3 (weights resident on chip) takes hidden_size 64 (got %lld)",
             who, (long long)H);
  if (ri32_family(precision, H)) {
    HY_REQUIRE(I > 0 && I <= 31 && IP == 32, HY2DL_ERR_SHAPE,
               "%s: the resident-weights tensor-core path takes input_size <= 31 in a 32-column RI32 stream (I=%lld IP=%lld)", who, (long long)I, (long long)IP);
    return HY2DL_OK;
  }
  HY_REQUIRE(I > 0 && I <= 64 && IP == round_up(I, 8), HY2DL_ERR_SHAPE,
             "%s: input_size=%lld (<= 64) with IP=%lld (must be I rounded up to 8)", who, (long long)I, (long long)IP);
  return HY2DL_OK;
}
}  // namespace hy2dl
using namespace hy2dl;

extern "C" int64_t hy2dl_workspace_bytes(int kind, int64_t B, int64_t T, int64_t I, int64_t H, int64_t precision) {
  if (kind == HY2DL_WS_LSTM_FWD || kind == HY2DL_WS_LSTM_BWD || kind == HY2DL_WS_LSTM_WGRAD) {
    if (ri32_family((int)precision, H)) return tc::lstm_tc_workspace(kind, B, T, I, H);
    return lstm_fp32_workspace(kind, B, T, I, H);
  }
  hy2dl::set_error("hy2dl_workspace_bytes: unknown kind %d", kind);
  return -1;
}

extern "C" int hy2dl_lstm_fwd(const float* xT, int64_t B, int64_t T, int64_t I, int64_t IP, int64_t H,
                              const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* h_seq,
                              float* c_seq, float* gates, float* h_last, void* ws, int64_t ws_bytes, int precision,
                              const hy2dl_head_fwd* head, hy2dl_stream_t stream) {
  int rc = check_lstm_shape("hy2dl_lstm_fwd", B, T, I, IP, H, precision);
  if (rc) return rc;
  HY_PTR(xT); HY_PTR(w_ih); HY_PTR(w_hh); HY_PTR(b_ih); HY_PTR(b_hh);
  HY_PTR_OPT(h_seq); HY_PTR_OPT(c_seq); HY_PTR_OPT(gates); HY_PTR_OPT(h_last);
  HY_REQUIRE(h_seq || h_last || head, HY2DL_ERR_ALIGN, "hy2dl_lstm_fwd: h_seq, h_last and head are all null");
  const int single = precision == HY2DL_PREC_TF32;
  HY_REQUIRE(head == nullptr || ri32_family(precision, H), HY2DL_ERR_SHAPE,
             "hy2dl_lstm_fwd: the fused head needs the resident-weights path (H = 64, precision tf32x3 or tf32; use hy2dl_head_map_fwd otherwise)");
  if (ri32_family(precision, H)) {
    HY_REQUIRE(tc::lstm_tc_supported(I, H), HY2DL_ERR_SHAPE,
               "hy2dl_lstm_fwd: the resident-weights tensor-core path supports hidden_size 64 and input_size <= 31 (got H=%lld I=%lld)",
               (long long)H, (long long)I);
    return tc::lstm_tc_fwd(xT, B, T, I, IP, w_ih, w_hh, b_ih, b_hh, h_seq, c_seq, gates, h_last, ws, ws_bytes, head, single,
                           (cudaStream_t)stream);
  }
  return lstm_fp32_fwd(xT, B, T, I, IP, H, w_ih, w_hh, b_ih, b_hh, h_seq, c_seq, gates, h_last, ws, ws_bytes, single,
                       (cudaStream_t)stream);
}

extern "C" int hy2dl_lstm_bwd(int64_t B, int64_t T, int64_t H, const float* w_hh, const float* gates,
                              const float* c_seq, const float* dh_seq, int64_t dh_t0, const float* dh_last, float* dG,
                              void* ws, int64_t ws_bytes, int precision, const hy2dl_head_bwd* head,
                              hy2dl_stream_t stream) {
  int rc = check_lstm_shape("hy2dl_lstm_bwd", B, T, 8, ri32_family(precision, H) ? 32 : 8, H, precision);
  if (rc) return rc;
  HY_PTR(w_hh); HY_PTR(gates); HY_PTR(c_seq); HY_PTR(dG); HY_PTR_OPT(dh_seq); HY_PTR_OPT(dh_last);
  HY_REQUIRE(dh_seq || dh_last || head, HY2DL_ERR_ALIGN, "hy2dl_lstm_bwd: no external gradient given");
  const int single = precision == HY2DL_PREC_TF32;
  HY_REQUIRE(head == nullptr || ri32_family(precision, H), HY2DL_ERR_SHAPE,
             "hy2dl_lstm_bwd: the fused head needs the resident-weights path (H = 64, precision tf32x3 or tf32; use hy2dl_head_map_bwd otherwise)");
  HY_REQUIRE(dh_t0 >= 0 && dh_t0 <= T, HY2DL_ERR_SHAPE, "hy2dl_lstm_bwd: dh_t0=%lld out of range", (long long)dh_t0);
  if (ri32_family(precision, H))
    return tc::lstm_tc_bwd(B, T, w_hh, gates, c_seq, dh_seq, dh_t0, dh_last, dG, ws, ws_bytes, head, single, (cudaStream_t)stream);
  return lstm_fp32_bwd(B, T, H, w_hh, gates, c_seq, dh_seq, dh_t0, dh_last, dG, ws, ws_bytes, single, (cudaStream_t)stream);
}

extern "C" int hy2dl_lstm_wgrad(const float* dG, const float* h_seq, const float* xT, int64_t B, int64_t T, int64_t I,
                                int64_t IP, int64_t H, float* dw_ih, float* dw_hh, float* db, void* ws,
                                int64_t ws_bytes, int precision, const hy2dl_head_wgrad* head, hy2dl_stream_t stream) {
  int rc = check_lstm_shape("hy2dl_lstm_wgrad", B, T, I, IP, H, precision);
  if (rc) return rc;
  HY_PTR(dG); HY_PTR(h_seq); HY_PTR(xT); HY_PTR(dw_ih); HY_PTR(dw_hh); HY_PTR(db);
  const int single = precision == HY2DL_PREC_TF32;
  if (ri32_family(precision, H)) {
    HY_REQUIRE(tc::lstm_tc_supported(I, H), HY2DL_ERR_SHAPE,
               "hy2dl_lstm_wgrad: the resident-weights tensor-core path supports hidden_size 64 and input_size <= 31 (got H=%lld I=%lld)",
               (long long)H, (long long)I);
    return tc::lstm_tc_wgrad(dG, h_seq, xT, B, T, I, IP, dw_ih, dw_hh, db, ws, ws_bytes, head, single, (cudaStream_t)stream);
  }
  HY_REQUIRE(head == nullptr, HY2DL_ERR_SHAPE, "hy2dl_lstm_wgrad: the fused head needs the resident-weights path (use hy2dl_head_map_bwd otherwise)");
  return lstm_fp32_wgrad(dG, h_seq, xT, B, T, I, IP, H, dw_ih, dw_hh, db, ws, ws_bytes, single, (cudaStream_t)stream);
}
