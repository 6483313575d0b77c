// model_cache.cpp — monthly model files (include/qcoh.h, "monthly model files").
// The reference re-expands `XGBoostFile: ..._M%m2.model` (OH_instance_OH.rc:20) on every Run1
// (fill_grads_template, /root/reference/OH_GridComp/OH_GridCompMod.F90:1187) yet loads its SAVE'd booster
// only once (:182,209,242-271).  libqcoh keeps that default; a host that wants the month's model opts in
// through the calls here.  Template expansion and the cache are host-side bookkeeping — no numerics.
#include <map>

#include "context.hpp"

using namespace qcoh;

namespace {

// keyed by the file name as given (after template expansion); owns the boosters
thread_local std::map<std::string, Booster *> g_cache;  // per host thread = per GPU (context.hpp)

void put2(std::string &o, int v, int width) {
  char buf[16];
  snprintf(buf, sizeof buf, "%0*d", width, v);
  o += buf;
}

std::string expand(const char *pattern, int nymd, int nhms) {
  if (!pattern) throw Error("qcoh_expand_template: pattern is NULL");
  if (nymd < 0 || nhms < 0) throw Error("qcoh_expand_template: negative nymd / nhms");
  const int yy = nymd / 10000, mm = (nymd % 10000) / 100, dd = nymd % 100;
  const int hh = nhms / 10000, nn = (nhms % 10000) / 100;
  static const char *mon[12] = {"jan", "feb", "mar", "apr", "may", "jun", "jul", "aug", "sep", "oct", "nov", "dec"};
  static const int cum[12] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334};
  auto month_ok = [&] {
    if (mm < 1 || mm > 12) throw Error("qcoh_expand_template: month " + std::to_string(mm) + " out of range in nymd");
  };
  std::string o;
  for (const char *p = pattern; *p; ++p) {
    if (*p != '%') {
      o += *p;
      continue;
    }
    const char a = p[1], b = a ? p[2] : 0;
    if (!a) throw Error(std::string("qcoh_expand_template: dangling '%' at the end of '") + pattern + "'");
    if (a == '%') {
      o += '%', ++p;
      continue;
    }
    const std::string tok = std::string(1, a) + (b ? std::string(1, b) : std::string());
    if (tok == "y4") put2(o, yy, 4);
    else if (tok == "y2") put2(o, yy % 100, 2);
    else if (tok == "m1") month_ok(), o += std::to_string(mm);
    else if (tok == "m2") month_ok(), put2(o, mm, 2);
    else if (tok == "mc") month_ok(), o += mon[mm - 1];
    else if (tok == "Mc") month_ok(), o += (char)(mon[mm - 1][0] - 32), o += mon[mm - 1] + 1;
    else if (tok == "MC") {
      month_ok();
      for (const char *q = mon[mm - 1]; *q; ++q) o += (char)(*q - 32);
    } else if (tok == "d1") o += std::to_string(dd);
    else if (tok == "d2") put2(o, dd, 2);
    else if (tok == "h1") o += std::to_string(hh);
    else if (tok == "h2") put2(o, hh, 2);
    else if (tok == "n2") put2(o, nn, 2);
    else if (tok == "j3") {
      month_ok();
      const bool leap = (yy % 4 == 0 && yy % 100 != 0) || yy % 400 == 0;
      put2(o, cum[mm - 1] + dd + ((leap && mm > 2) ? 1 : 0), 3);
    } else
      throw Error("qcoh_expand_template: unknown token '%" + tok + "' in '" + pattern + "'");
    p += 2;
  }
  return o;
}

Booster *cache_get(const std::string &fname) {
  auto it = g_cache.find(fname);
  if (it != g_cache.end()) return it->second;
  std::unique_ptr<Booster> b(new Booster());
  HostForest hf = load_model_file(fname.c_str());
  FlatForest ff = flatten(hf);
  DuoForest df = build_duo(ff, hf.num_feature);
  b->host = std::move(hf), b->flat = std::move(ff), b->duo = std::move(df);
  b->loaded = true, b->uploaded = false;
  b->version = ++g_version_counter;
  b->cache_owned = true;  // uploaded at first use (qcoh_oh_set_booster / predict), like qcoh_booster_parse
  Booster *raw = b.release();
  g_live_handles.insert(raw);
  g_cache[fname] = raw;
  return raw;
}

}  // namespace

extern "C" {

int qcoh_expand_template(const char *pattern, int nymd, int nhms, char *out, size_t cap) {
  API_BEGIN
  if (!out) throw Error("qcoh_expand_template: out is NULL");
  const std::string s = expand(pattern, nymd, nhms);
  if (s.size() + 1 > cap) throw Error("qcoh_expand_template: result needs " + std::to_string(s.size() + 1) + " bytes");
  memcpy(out, s.c_str(), s.size() + 1);
  API_END
}

int qcoh_model_cache_get(const char *fname, BoosterHandle *out) {
  API_BEGIN
  if (!fname || !out) throw Error("qcoh_model_cache_get: NULL argument");
  *out = cache_get(fname);
  API_END
}

int qcoh_model_cache_size(void) { return (int)g_cache.size(); }

int qcoh_model_cache_clear(void) {
  API_BEGIN
  for (auto &kv : g_cache)
    if (kv.second->oh_refs > 0)
      throw Error("qcoh_model_cache_clear: the booster of '" + kv.first + "' is still used by a fused-Run1 handle (qcoh_oh_free first)");
  if (g.ready) CU(cudaStreamSynchronize(g.stream));
  for (auto &kv : g_cache) {
    Booster *b = kv.second;
    if (g_last_booster == b) g_last_booster = nullptr;
    b->magic = 0;
    g_live_handles.erase(b);
    delete b;
  }
  g_cache.clear();
  API_END
}

int qcoh_oh_select_model(qcoh_oh_handle h, const char *pattern, int nymd, int nhms, int *changed) {
  API_BEGIN
  if (changed) *changed = 0;
  const std::string fname = expand(pattern, nymd, nhms);
  Booster *b = cache_get(fname);
  BoosterHandle cur = nullptr;
  if (qcoh_oh_get_booster(h, &cur) != 0) return -1;
  if (cur != (BoosterHandle)b) {
    if (qcoh_oh_set_booster(h, b) != 0) return -1;
    if (changed) *changed = 1;
  }
  API_END
}

}  // extern "C"
