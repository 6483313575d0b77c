// run1_control.cpp — the host-side decisions of Run1 that surround the fused call (include/qcoh.h,
// "Run1 control"): when to boost, the 24-hour-average spin-up switch, and which import feeds each
// boost-state field (/root/reference/OH_GridComp/OH_GridCompMod.F90:1189-1193, :1307-1320, :1326-1548).
// In a MAPL host these stay in Fortran (fortran/OH_Run1_fused.F90 binds the chosen pointers); they are
// restated here so that a non-Fortran host, and the tests, take the same decisions.  No numerics, no GPU.
#include <cstring>
#include <string>

#include "context.hpp"

using namespace qcoh;

namespace {

// fields selected per OH_data_source (:1326-1435, :1496-1507, :1524-1528); the 4-D ones carry a wavelength axis online
const char *const kSelected[] = {"T", "Q", "PLE", "ZLE", "TAUCLW", "TAUCLI", "CH4", "CO", "FCLD"};
const char *const kSelected4d[] = {"BCSCACOEF", "OCSCACOEF", "BRSCACOEF", "DUSCACOEF", "SUSCACOEF", "SSSCACOEF", "NISCACOEF"};
// always climatological, import name oh_<field> (:1441-1442, :1493-1494, :1510-1517, :1535, :1540, :1548)
const char *const kClimo[] = {"NO2", "O3", "ISOP", "ACET", "C2H6", "C3H8", "PRPE", "ALK4", "MP", "H2O2", "CH2O",
                              "ALBUV", "GMITO3", "GMITTO3", "OH"};

template <size_t N>
bool in(const char *const (&list)[N], const std::string &s) {
  for (const char *e : list)
    if (s == e) return true;
  return false;
}

}  // namespace

extern "C" {

int qcoh_data_source_from_name(const char *token) {
  // :551-553 — an unknown token leaves the reference's field unset; here it is an error
  if (!token) return -1;
  const std::string t(token);
  if (t == "PRECOMPUTED") return QCOH_PRECOMPUTED;
  if (t == "ONLINE_INST") return QCOH_ONLINE_INST;
  if (t == "ONLINE_AVG24") return QCOH_ONLINE_AVG24;
  return -1;
}

int qcoh_need_to_call_boost(int compute_once_per_day, int nhms) {
  // :1189-1193 — with compute_once_per_day the boost runs only on the step whose hhmmss is exactly 0
  return (compute_once_per_day && nhms > 0) ? 0 : 1;
}

int qcoh_use_inst_values(int data_source, float t_avg24_first) {
  // :1307-1320 — within the first 24 hours the coupler hands back all-zero daily means; the reference tests
  // the first element of T_avg24 and assumes every other daily-mean import follows
  return (data_source == QCOH_ONLINE_AVG24 && t_avg24_first == 0.0f) ? 1 : 0;
}

int qcoh_import_name(const char *field, int data_source, int use_inst_values, char *out, size_t cap, int *is_4d) {
  API_BEGIN
  if (!field || !out) throw Error("qcoh_import_name: NULL argument");
  if (data_source != QCOH_PRECOMPUTED && data_source != QCOH_ONLINE_INST && data_source != QCOH_ONLINE_AVG24)
    throw Error("qcoh_import_name: OH_data_source must be 1 (PRECOMPUTED), 2 (ONLINE_INST) or 3 (ONLINE_AVG24)");
  const std::string f(field);
  std::string name;
  int four_d = 0;
  if (in(kClimo, f)) {
    name = "oh_" + f;
  } else if (in(kSelected, f) || in(kSelected4d, f)) {
    const bool sca = in(kSelected4d, f);
    if (data_source == QCOH_PRECOMPUTED)
      name = "oh_" + f;  // 3-D also for the scattering coefficients (:1388)
    else if (data_source == QCOH_ONLINE_INST || use_inst_values)
      name = f, four_d = sca;
    else
      name = f + "_avg24", four_d = sca;
  } else if (f == "T_MOD" || f == "Q_MOD" || f == "PLE_MOD" || f == "TROPP") {
    name = f == "TROPP" ? f : f.substr(0, f.size() - 4);  // the current model state, whatever the source (:1233-1236)
  } else {
    throw Error("qcoh_import_name: '" + f + "' is not a field of the OH boost inputs");
  }
  if (name.size() + 1 > cap) throw Error("qcoh_import_name: result needs " + std::to_string(name.size() + 1) + " bytes");
  memcpy(out, name.c_str(), name.size() + 1);
  if (is_4d) *is_4d = four_d;
  API_END
}

}  // extern "C"
