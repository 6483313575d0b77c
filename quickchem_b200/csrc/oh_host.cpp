// oh_host.cpp — host-side mirror of the reference's driver routine `predict_OH_with_XGB`
// (/root/reference/OH_GridComp/OH_GridCompMod.F90:123-398), restated in C++ because this
// build box has no Fortran compiler.  It calls ONLY the eleven XGBoost-named C symbols that
// xgb_fortran_api.F90 binds, in the reference's order, with the reference's ownership rules —
// i.e. it is what the unmodified Fortran does against libqcoh.so, and serves as the executable
// proof that the library is a drop-in behind that interface.  fortran/ holds the Fortran text.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qcoh.h"

namespace {

// SAVE variables of the reference (OH_GridCompMod.F90:182,209): one booster per process
// (thread_local: a host thread is one rank bound to one GPU, see context.hpp)
thread_local BoosterHandle xx_bst = nullptr;
thread_local bool first_time = true;
// opt-in fix of SURVEY.md 0.5 (the file name carries the month but the booster is loaded once)
thread_local bool reload_on_file_change = false;
thread_local std::string loaded_fname;

constexpr int64_t xx_param_count = 27;  // :228
constexpr float xx_miss = -999.0f;      // :213

}  // namespace

// the library's last-error string (XGBGetLastError's), for the messages of the reference's own _ASSERTs;
// internal to libqcoh.so (hidden), defined in capi_xgb.cpp
extern "C" __attribute__((visibility("hidden"))) void qcoh_internal_set_error(const char *msg);

namespace {

template <class F>
void parallel_for(int64_t n, F &&body) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  if (nt > 64) nt = 64;
  if (n < (int64_t)nt * 4096) nt = 1;
  if (nt == 1) {
    body((int64_t)0, n);
    return;
  }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back([=, &body] { body(n * t / nt, n * (t + 1) / nt); });
  for (auto &t : th) t.join();
}

}  // namespace

extern "C" void qcoh_predict_OH_reset(void) {
  if (xx_bst) XGBoosterFree(xx_bst);
  xx_bst = nullptr;
  first_time = true;
  loaded_fname.clear();
}

extern "C" void qcoh_predict_OH_reload_on_file_change(int on) { reload_on_file_change = on != 0; }

// Arrays are Fortran (icount, jcount, kcount) column-major == C [k][j][i]; the (i,j) members of bb
// are [j][i].  Returns 0 (ESMF_SUCCESS) or -1 with XGBGetLastError() / the _ASSERT message.
extern "C" int qcoh_predict_OH_with_XGB(const char *xgb_fname, int icount, int jcount, int kcount, int dynamic_k_range,
                                        float tropp_min, const float *pl, const float *tropp, const float *const bb[27],
                                        const int is2d[27], float *OH_ML) {
  const int64_t ncol = (int64_t)icount * jcount;
  int rc;

  // ---- INIT (:242-271)
  if (first_time) {
    std::vector<float> xx_carr_small((size_t)xx_param_count, 0.0f);  // just 1 prediction
    DMatrixHandle xx_dmtrx = nullptr;
    rc = XGDMatrixCreateFromMat(xx_carr_small.data(), 1, (bst_ulong)xx_param_count, xx_miss, &xx_dmtrx);
    if (rc != 0) return -1;  // _ASSERT 'Failed in XGDMatrixCreateFromMat_f'
    // the reference hands the DMatrix handle itself as `dmats` with len = 0 (:255-256)
    rc = XGBoosterCreate((const DMatrixHandle *)xx_dmtrx, 0, &xx_bst);
    if (rc != 0) {
      XGDMatrixFree(xx_dmtrx);
      return -1;
    }
    rc = XGBoosterLoadModel(xx_bst, xgb_fname);
    if (rc != 0) {  // a retry must not leak this pair (first_time stays true); keep the loader's message
      const std::string why = XGBGetLastError();
      XGDMatrixFree(xx_dmtrx);
      XGBoosterFree(xx_bst);
      xx_bst = nullptr;
      qcoh_internal_set_error(why.c_str());
      return -1;
    }
    rc = XGDMatrixFree(xx_dmtrx);
    if (rc != 0) return -1;
    first_time = false;
    loaded_fname = xgb_fname;
  } else if (reload_on_file_change && loaded_fname != xgb_fname) {
    // not in the reference: its SAVE'd booster keeps the first month's model for the whole run
    rc = XGBoosterLoadModel(xx_bst, xgb_fname);
    if (rc != 0) return -1;
    loaded_fname = xgb_fname;
  }

  // ---- RUN: level slab (:275-301)
  int ksubcount = 0;
  if (!dynamic_k_range) {
    for (int64_t c = 0; c < ncol; ++c)
      if (tropp[c] <= tropp_min) {  // _ASSERT(ALL(tropp > tropp_min), ...) (:287-288)
        qcoh_internal_set_error("OH Prediction: Minimum tropopause pressure is not low enough!");
        return -1;
      }
  }
  for (int64_t c = 0; c < ncol; ++c) {
    const float cmp = dynamic_k_range ? tropp[c] : tropp_min;
    int k = 0;
    for (int kk = 0; kk < kcount; ++kk) k += pl[(size_t)kk * ncol + c] > cmp;
    if (k > ksubcount) ksubcount = k;
  }
  const int k1 = kcount - ksubcount + 1, k2 = kcount;
  (void)k2;

  // ---- pack xx_carr(27, N) (:303-345); m runs over k = k1..k2, j, i (i fastest)
  const int64_t xx_prediction_count = ncol * ksubcount;
  float *xx_carr = (float *)malloc(sizeof(float) * (size_t)(xx_prediction_count ? xx_prediction_count : 1) * xx_param_count);
  if (!xx_carr) return -1;
  parallel_for(xx_prediction_count, [&](int64_t m0, int64_t m1) {
    for (int64_t m = m0; m < m1; ++m) {
      const int64_t c = m % ncol;
      const size_t e = (size_t)(k1 - 1) * ncol + (size_t)m;
      float *row = xx_carr + (size_t)m * xx_param_count;
      for (int f = 0; f < 27; ++f) row[f] = is2d[f] ? bb[f][c] : bb[f][e];
      row[1] = bb[1][e] / 100.0f;  // convert Pa to hPa (:314)
    }
  });

  DMatrixHandle xx_dmtrx = nullptr;
  rc = XGDMatrixCreateFromMat(xx_carr, (bst_ulong)xx_prediction_count, (bst_ulong)xx_param_count, xx_miss, &xx_dmtrx);
  if (rc != 0) {
    free(xx_carr);
    return -1;
  }

  // ---- predict (:356-359): option_mask = 0, ntree_limit = 0, training = 0 (:231-235)
  bst_ulong xx_pred_len = 0;
  const float *xx_pred = nullptr;
  rc = XGBoosterPredict(xx_bst, xx_dmtrx, 0, 0, 0, &xx_pred_len, &xx_pred);
  if (rc != 0 || xx_pred_len != (bst_ulong)xx_prediction_count) {
    const std::string why = rc != 0 ? std::string(XGBGetLastError()) : "XGBoosterPredict_f returned the wrong number of predictions";  // (:359)
    XGDMatrixFree(xx_dmtrx);
    free(xx_carr);
    qcoh_internal_set_error(why.c_str());
    return -1;
  }

  // ---- OH_ML(i,j,k) = 10.0 ** pred(m) (:364-374), levels above k1 untouched
  float *dst = OH_ML + (size_t)(k1 - 1) * ncol;
  parallel_for(xx_prediction_count, [&](int64_t m0, int64_t m1) {
    for (int64_t m = m0; m < m1; ++m) dst[m] = powf(10.0f, xx_pred[m]);
  });

  rc = XGDMatrixFree(xx_dmtrx);  // :377
  free(xx_carr);                 // :383
  return rc == 0 ? 0 : -1;
}
