/* qc_oracle.c — see qc_oracle.h.  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (no reference
 * golden vectors exist; libxgboost 1.6.0 is an absent third-party dependency).
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile). */
#include "qc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static _Thread_local char g_err[512];
const char *orc_last_error(void) { return g_err; }
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
#define FAIL(ret, ...)                          \
  do {                                          \
    snprintf(g_err, sizeof g_err, __VA_ARGS__); \
    return ret;                                 \
  } while (0)

/* ======================================================================================
 * Model loading — XGBoost 1.6.0 legacy binary.
 * Follows LearnerIO::LoadModel (src/learner.cc), GBTreeModel::Load (src/gbm/gbtree_model.cc)
 * and RegTree::Load (src/tree/tree_model.cc); strings/vectors are dmlc serializer
 * uint64-length-prefixed.  Reference call site: OH_GridCompMod.F90:261 (XGBoosterLoadModel_f).
 * ==================================================================================== */
typedef struct {
  const unsigned char *p, *end;
} rd_t;

static int rd_bytes(rd_t *r, void *dst, size_t n) {
  if ((size_t)(r->end - r->p) < n) return 0;
  memcpy(dst, r->p, n);
  r->p += n;
  return 1;
}
static int rd_str(rd_t *r, char *dst, size_t cap) {
  uint64_t len;
  if (!rd_bytes(r, &len, 8)) return 0;
  if ((uint64_t)(r->end - r->p) < len) return 0;
  size_t c = len < cap - 1 ? (size_t)len : cap - 1;
  memcpy(dst, r->p, c);
  dst[c] = 0;
  r->p += len;
  return 1;
}

void orc_model_free(orc_model *m) {
  if (!m) return;
  if (m->trees)
    for (int i = 0; i < m->num_trees; ++i) {
      free(m->trees[i].nodes);
      free(m->trees[i].stats);
    }
  free(m->trees);
  free(m->tree_info);
  free(m);
}

orc_model *orc_model_load(const char *path) {
  FILE *fp = fopen(path, "rb");
  if (!fp) FAIL(NULL, "cannot open %s", path);
  fseek(fp, 0, SEEK_END);
  long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  unsigned char *buf = (unsigned char *)malloc((size_t)sz + 1);
  if (!buf || fread(buf, 1, (size_t)sz, fp) != (size_t)sz) {
    fclose(fp);
    free(buf);
    FAIL(NULL, "short read on %s", path);
  }
  fclose(fp);
  rd_t r = {buf, buf + sz};
  if (sz >= 1 && buf[0] == '{') {
    free(buf);
    FAIL(NULL, "JSON/UBJ model: not handled by the C oracle (see oracle/naive.py)");
  }
  if (sz >= 4 && memcmp(buf, "bs64", 4) == 0) {
    free(buf);
    FAIL(NULL, "Base64 format is not supported");
  }
  if (sz >= 4 && memcmp(buf, "binf", 4) == 0) r.p += 4;

  orc_model *m = (orc_model *)calloc(1, sizeof *m);
  /* LearnerModelParamLegacy, 136 bytes */
  struct {
    float base_score;
    uint32_t num_feature;
    int32_t num_class, contain_extra_attrs, contain_eval_metrics;
    uint32_t major_version, minor_version, num_target;
    int32_t reserved[26];
  } mp;
  _Static_assert(sizeof mp == 136, "LearnerModelParamLegacy");
  if (!rd_bytes(&r, &mp, sizeof mp)) goto trunc;
  m->base_score = mp.base_score;
  m->num_feature = mp.num_feature;
  m->major_version = mp.major_version;
  m->minor_version = mp.minor_version;
  if (!rd_str(&r, m->objective, sizeof m->objective)) goto trunc;
  if (!rd_str(&r, m->booster, sizeof m->booster)) goto trunc;
  if (strcmp(m->booster, "gbtree") != 0) {
    snprintf(g_err, sizeof g_err, "unsupported booster '%s'", m->booster);
    goto bad;
  }
  /* identity-transform objectives only (reg:linear is the pre-1.0 alias) */
  if (strcmp(m->objective, "reg:squarederror") != 0 && strcmp(m->objective, "reg:linear") != 0) {
    snprintf(g_err, sizeof g_err, "unsupported objective '%s'", m->objective);
    goto bad;
  }
  /* GBTreeModelParam, 160 bytes */
  struct {
    int32_t num_trees, num_roots, num_feature, pad;
    int64_t num_pbuffer;
    int32_t num_output_group, size_leaf_vector;
    int32_t reserved[32];
  } gp;
  _Static_assert(sizeof gp == 160, "GBTreeModelParam");
  if (!rd_bytes(&r, &gp, sizeof gp)) goto trunc;
  if (gp.num_trees < 0) {
    snprintf(g_err, sizeof g_err, "negative num_trees");
    goto bad;
  }
  m->num_trees = gp.num_trees;
  m->trees = (orc_tree *)calloc((size_t)gp.num_trees + 1, sizeof(orc_tree));
  for (int t = 0; t < gp.num_trees; ++t) {
    struct {
      int32_t num_roots, num_nodes, num_deleted, max_depth, num_feature, size_leaf_vector;
      int32_t reserved[31];
    } tp;
    _Static_assert(sizeof tp == 148, "TreeParam");
    if (!rd_bytes(&r, &tp, sizeof tp)) goto trunc;
    if (tp.num_nodes <= 0) {
      snprintf(g_err, sizeof g_err, "tree %d: num_nodes=%d", t, tp.num_nodes);
      goto bad;
    }
    orc_tree *tr = &m->trees[t];
    tr->num_nodes = tp.num_nodes;
    tr->nodes = (orc_node *)malloc(sizeof(orc_node) * (size_t)tp.num_nodes);
    tr->stats = (orc_stat *)malloc(sizeof(orc_stat) * (size_t)tp.num_nodes);
    if (!rd_bytes(&r, tr->nodes, sizeof(orc_node) * (size_t)tp.num_nodes)) goto trunc;
    if (!rd_bytes(&r, tr->stats, sizeof(orc_stat) * (size_t)tp.num_nodes)) goto trunc;
    for (int n = 0; n < tp.num_nodes; ++n) {
      const orc_node *nd = &tr->nodes[n];
      if (nd->cleft == -1) continue;
      if (nd->cleft < 0 || nd->cleft >= tp.num_nodes || nd->cright < 0 || nd->cright >= tp.num_nodes) {
        snprintf(g_err, sizeof g_err, "tree %d node %d: child out of range", t, n);
        goto bad;
      }
      if ((nd->sindex & 0x7FFFFFFFu) >= m->num_feature) {
        snprintf(g_err, sizeof g_err, "tree %d node %d: split feature out of range", t, n);
        goto bad;
      }
    }
  }
  m->tree_info = (int32_t *)calloc((size_t)gp.num_trees + 1, sizeof(int32_t));
  if (gp.num_trees && !rd_bytes(&r, m->tree_info, sizeof(int32_t) * (size_t)gp.num_trees)) goto trunc;
  /* trailing attributes / metric names do not affect prediction and are not parsed here */
  free(buf);
  return m;
trunc:
  snprintf(g_err, sizeof g_err, "truncated model file %s", path);
bad:
  free(buf);
  orc_model_free(m);
  return NULL;
}

/* ======================================================================================
 * DMatrix — XGDMatrixCreateFromMat (src/c_api/c_api.cc) -> DenseAdapter -> SparsePage::Push
 * (src/data/data.cc).  Reference call site: OH_GridCompMod.F90:251,347 with missing=-999.0.
 * ==================================================================================== */
void orc_dmatrix_free(orc_dmatrix *d) {
  if (!d) return;
  free(d->offset);
  free(d->index);
  free(d->value);
  free(d);
}

orc_dmatrix *orc_dmatrix_from_mat(const float *data, uint64_t nrow, uint64_t ncol, float missing) {
  orc_dmatrix *d = (orc_dmatrix *)calloc(1, sizeof *d);
  d->nrow = nrow;
  d->ncol = ncol;
  d->offset = (uint64_t *)calloc(nrow + 1, sizeof(uint64_t));
  int bad = 0;
  const int miss_is_inf = isinf(missing);
  /* pass 1: count valid entries per row */
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int64_t i = 0; i < (int64_t)nrow; ++i) {
    const float *row = data + (size_t)i * ncol;
    uint64_t c = 0;
    for (uint64_t j = 0; j < ncol; ++j) {
      float v = row[j];
      if (!miss_is_inf && isinf(v)) bad = 1;
      if (!isnan(v) && v != missing) ++c;
    }
    d->offset[i + 1] = c;
  }
  if (bad) {
    orc_dmatrix_free(d);
    FAIL(NULL, "Input data contains `inf` or `nan`");
  }
  for (uint64_t i = 0; i < nrow; ++i) d->offset[i + 1] += d->offset[i];
  uint64_t nnz = d->offset[nrow];
  d->index = (uint32_t *)malloc(sizeof(uint32_t) * (nnz ? nnz : 1));
  d->value = (float *)malloc(sizeof(float) * (nnz ? nnz : 1));
  /* pass 2: fill */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)nrow; ++i) {
    const float *row = data + (size_t)i * ncol;
    uint64_t o = d->offset[i];
    for (uint64_t j = 0; j < ncol; ++j) {
      float v = row[j];
      if (!isnan(v) && v != missing) {
        d->index[o] = (uint32_t)j;
        d->value[o] = v;
        ++o;
      }
    }
  }
  return d;
}

/* ======================================================================================
 * Predict — CPUPredictor::PredictDMatrix / PredictByAllTrees / GetLeafIndex / GetNextNode
 * (src/predictor/cpu_predictor.cc, src/predictor/predict_fn.h), XGBoost 1.6.0:
 * rows in blocks of 64; per block fill one dense FVec per row (absent => missing flag);
 * tree-outer / row-inner; nid = missing ? DefaultChild : cleft + !(fvalue < split_cond);
 * out[row] += leaf_value in float32, starting from base_score.
 * Reference call site: OH_GridCompMod.F90:356 (option_mask=0, ntree_limit=0, training=0).
 * ==================================================================================== */
#define ORC_BLOCK 64

typedef union {
  float fvalue;
  int32_t flag;
} fvec_entry;

static inline int leaf_index(const orc_node *nodes, const fvec_entry *fv, int has_missing) {
  int nid = 0;
  while (nodes[nid].cleft != -1) {
    const orc_node *nd = &nodes[nid];
    uint32_t f = nd->sindex & 0x7FFFFFFFu;
    if (has_missing && fv[f].flag == -1) {
      nid = (nd->sindex >> 31) ? nd->cleft : nd->cright;
    } else {
      nid = nd->cleft + !(fv[f].fvalue < nd->info);
    }
  }
  return nid;
}

uint64_t orc_predict(const orc_model *m, const orc_dmatrix *d, int option_mask, unsigned ntree_limit,
                     float *out) {
  if (!m || !d) FAIL(0, "null handle");
  if (d->ncol > m->num_feature && m->num_feature != 0)
    FAIL(0, "Number of columns does not match number of features in booster.");
  int ntree = m->num_trees;
  if (ntree_limit != 0 && (int)ntree_limit < ntree) ntree = (int)ntree_limit;
  const uint32_t nf = m->num_feature;
  const int pred_leaf = (option_mask & 2) != 0;
  if (option_mask & ~3) FAIL(0, "option_mask %d not supported by the oracle", option_mask);
  const int64_t nblock = (int64_t)((d->nrow + ORC_BLOCK - 1) / ORC_BLOCK);
#pragma omp parallel
  {
    fvec_entry *fv = (fvec_entry *)malloc(sizeof(fvec_entry) * (size_t)nf * ORC_BLOCK);
    int has_missing[ORC_BLOCK];
#pragma omp for schedule(static)
    for (int64_t b = 0; b < nblock; ++b) {
      const uint64_t r0 = (uint64_t)b * ORC_BLOCK;
      const int bs = (int)((d->nrow - r0) < ORC_BLOCK ? (d->nrow - r0) : ORC_BLOCK);
      /* FVec::Init + Fill */
      memset(fv, 0xFF, sizeof(fvec_entry) * (size_t)nf * (size_t)bs);
      for (int i = 0; i < bs; ++i) {
        uint64_t o0 = d->offset[r0 + i], o1 = d->offset[r0 + i + 1];
        for (uint64_t o = o0; o < o1; ++o)
          if (d->index[o] < nf) fv[(size_t)i * nf + d->index[o]].fvalue = d->value[o];
        has_missing[i] = (o1 - o0) != nf;
      }
      if (pred_leaf) {
        for (int i = 0; i < bs; ++i)
          for (int t = 0; t < ntree; ++t)
            out[(r0 + i) * (uint64_t)ntree + t] =
                (float)leaf_index(m->trees[t].nodes, fv + (size_t)i * nf, has_missing[i]);
      } else {
        for (int i = 0; i < bs; ++i) out[r0 + i] = m->base_score;
        for (int t = 0; t < ntree; ++t) {
          const orc_node *nodes = m->trees[t].nodes;
          for (int i = 0; i < bs; ++i) {
            int nid = leaf_index(nodes, fv + (size_t)i * nf, has_missing[i]);
            out[r0 + i] += nodes[nid].info;
          }
        }
        /* reg:squarederror PredTransform is the identity (src/objective/regression_obj.cu) */
      }
    }
    free(fv);
  }
  return pred_leaf ? d->nrow * (uint64_t)ntree : d->nrow;
}

/* ======================================================================================
 * Run1 restatement (OH_GridCompMod.F90:1232-1599).  All arithmetic is float32 in the
 * Fortran evaluation order; this file must be compiled with -ffp-contract=off.
 * ==================================================================================== */
static int leap_year(int ny) { /* OH_GridCompMod.F90:1940-1971 */
  if (ny >= 0) {
    if (ny % 100 == 0 && ny % 400 == 0) return 1;
    if (ny % 4 == 0 && ny % 100 != 0) return 1;
  }
  return 0;
}

int orc_julian_day(int nymd) { /* OH_GridCompMod.F90:1905-1936 */
  static const int days[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
  int ny = nymd / 10000, mm = (nymd % 10000) / 100, dd = nymd % 100;
  int ds = dd;
  for (int mth = 1; mth <= mm - 1; ++mth) ds += (mth == 2 && leap_year(ny)) ? 29 : days[mth - 1];
  return ds;
}

void orc_noon_sza(int jday, const float *lat, const float *lon, int n, float r2d, float d2r,
                  float *sza) { /* OH_GridCompMod.F90:401-466 */
  const float sindec = 0.3978f * sinf(0.9863f * ((float)jday - 80.0f) * d2r);
  const float soldek = asinf(sindec);
  const float cosdec = cosf(soldek);
  for (int i = 0; i < n; ++i) {
    float sinlat = sinf(lat[i]);
    float sollat = asinf(sinlat);
    float coslat = cosf(sollat);
    float mylon = lon[i] * r2d;
    if (mylon > 180.0f) mylon = mylon - 360.0f;
    if (mylon < -180.0f) mylon = mylon + 360.0f;
    float tau = 12.0f + (mylon / -180.0f) * 12.0f;
    float loct = ((tau * 15.0f) - 180.0f) * d2r + lon[i];
    float cosz = cosdec * coslat * cosf(loct) + sindec * sinlat;
    cosz = fminf(1.0f, cosz);
    cosz = fmaxf(-1.0f, cosz);
    sza[i] = acosf(cosz) * r2d;
  }
}

int orc_run1(const orc_model *m, const orc_run1_in *in, orc_run1_out *out) {
  const int nc = in->ncol, km = in->km;
  const size_t n3 = (size_t)nc * km;
  float *PL_MOD = (float *)malloc(n3 * 4), *NDWET = (float *)malloc(n3 * 4);
  float *PL_BST = (float *)malloc(n3 * 4), *aod = (float *)malloc(n3 * 4);
  float *wdn = (float *)malloc(n3 * 4), *idn = (float *)malloc(n3 * 4), *iup = (float *)malloc(n3 * 4);
  float *wup = (float *)malloc(n3 * 4), *aup = (float *)malloc(n3 * 4), *adn = (float *)malloc(n3 * 4);
  float *lat = (float *)malloc((size_t)nc * 4), *so3 = (float *)malloc((size_t)nc * 4);
  float *sza = (float *)malloc((size_t)nc * 4);
  float *OH_ML = (float *)calloc(n3, 4); /* self%OH_ML(:,:,:) = 0.0  (:1559) */
  int rc = 0;

  /* :1246-1257 model-state fields */
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < (int64_t)n3; ++e) {
    float pl = (in->PLE_MOD[e] + in->PLE_MOD[e + nc]) * 0.5f;
    float q = in->Q_MOD[e];
    float tv = in->T_MOD[e] * (1.0f + q / in->mapl_epsilon) / (1.0f + q);
    PL_MOD[e] = pl;
    NDWET[e] = (in->mapl_avogad * pl) / (in->mapl_runiv * tv);
  }
  /* :1441-1466, :1488 derived fields */
  for (int c = 0; c < nc; ++c) {
    lat[c] = in->LATS[c] * in->mapl_radians_to_degrees;
    so3[c] = in->GMITO3[c] - in->GMITTO3[c];
  }
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < (int64_t)n3; ++e) {
    double thick = (double)(in->ZLE_BST[e] - in->ZLE_BST[e + nc]); /* REAL*8 gridBoxThickness */
    float s = in->SCA[0][e] + in->SCA[1][e];
    s = s + in->SCA[2][e];
    s = s + in->SCA[3][e];
    s = s + in->SCA[4][e];
    s = s + in->SCA[5][e];
    s = s + in->SCA[6][e];
    aod[e] = (float)(thick * (double)s);
    PL_BST[e] = (in->PLE_BST[e] + in->PLE_BST[e + nc]) * 0.5f;
  }
  /* :1468-1478 — every SUM(x(:,:,a:b),3) restarts from its first level and adds downward */
#pragma omp parallel for schedule(static)
  for (int c = 0; c < nc; ++c) {
    for (int k = 0; k < km; ++k) {
      float s_wdn = 0.f, s_idn = 0.f, s_adn = 0.f, s_iup = 0.f, s_wup = 0.f, s_aup = 0.f;
      for (int kk = k; kk < km; ++kk) {
        size_t e = (size_t)kk * nc + c;
        s_wdn += in->TAUCLW[e];
        s_idn += in->TAUCLI[e];
        s_adn += aod[e];
      }
      for (int kk = 0; kk <= k; ++kk) {
        size_t e = (size_t)kk * nc + c;
        s_iup += in->TAUCLI[e];
        s_wup += in->TAUCLW[e];
        s_aup += aod[e];
      }
      size_t e = (size_t)k * nc + c;
      wdn[e] = s_wdn, idn[e] = s_idn, iup[e] = s_iup, wup[e] = s_wup, aup[e] = s_aup, adn[e] = s_adn;
    }
  }
  orc_noon_sza(orc_julian_day(in->nymd), in->LATS, in->LONS, nc, in->mapl_radians_to_degrees,
               in->mapl_degrees_to_radians, sza);

  /* predict_OH_with_XGB :275-301 level slab */
  int ksub = 0;
  if (!in->compute_once_per_day) { /* dynamic_k_range */
    for (int c = 0; c < nc; ++c) {
      int k = 0;
      for (int kk = 0; kk < km; ++kk) k += PL_MOD[(size_t)kk * nc + c] > in->TROPP[c];
      if (k > ksub) ksub = k;
    }
  } else {
    for (int c = 0; c < nc; ++c)
      if (in->TROPP[c] <= in->tropp_min) {
        snprintf(g_err, sizeof g_err, "OH Prediction: Minimum tropopause pressure is not low enough!");
        rc = -1;
        goto done;
      }
    for (int c = 0; c < nc; ++c) {
      int k = 0;
      for (int kk = 0; kk < km; ++kk) k += PL_MOD[(size_t)kk * nc + c] > in->tropp_min;
      if (k > ksub) ksub = k;
    }
  }
  const int k1 = km - ksub + 1; /* 1-based */
  out->k1 = k1;
  const uint64_t npred = (uint64_t)nc * ksub;
  {
    /* :303-345 pack xx_carr(27, N); m runs over k=k1..k2, then column (i fastest) */
    const float *src3[27] = {0};
    const float *src2[27] = {0};
    src2[0] = lat, src3[1] = PL_BST, src3[2] = in->T_BST, src3[3] = in->NO2, src3[4] = in->O3;
    src3[5] = in->CH4, src3[6] = in->CO, src3[7] = in->ISOP, src3[8] = in->ACET, src3[9] = in->C2H6;
    src3[10] = in->C3H8, src3[11] = in->PRPE, src3[12] = in->ALK4, src3[13] = in->MP, src3[14] = in->H2O2;
    src3[15] = wdn, src3[16] = idn, src3[17] = iup, src3[18] = wup, src3[19] = in->FCLD;
    src3[20] = in->Q_BST, src2[21] = so3, src2[22] = in->ALBUV, src3[23] = aup, src3[24] = adn;
    src3[25] = in->CH2O, src2[26] = sza;
    float *X = out->X ? out->X : (float *)malloc((npred ? npred : 1) * 27 * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int64_t mm = 0; mm < (int64_t)npred; ++mm) {
      int c = (int)(mm % nc);
      int k = k1 - 1 + (int)(mm / nc);
      size_t e = (size_t)k * nc + c;
      float *row = X + (size_t)mm * 27;
      for (int f = 0; f < 27; ++f) row[f] = src3[f] ? src3[f][e] : src2[f][c];
      row[1] = PL_BST[e] / 100.0f; /* :314 Pa -> hPa, true divide */
    }
    for (int f = 0; f < 27; ++f)
      if (out->feat3d[f]) {
        if (src3[f])
          memcpy(out->feat3d[f], src3[f], n3 * 4);
        else
          memcpy(out->feat3d[f], src2[f], (size_t)nc * 4);
      }
    float *pred = out->pred ? out->pred : (float *)malloc((npred ? npred : 1) * sizeof(float));
    orc_dmatrix *d = orc_dmatrix_from_mat(X, npred, 27, in->missing);
    if (!d) {
      rc = -1;
    } else {
      if (npred && orc_predict(m, d, 0, 0, pred) != npred) rc = -1;
      orc_dmatrix_free(d);
    }
    if (rc == 0) {
      /* :364-374  OH_ML(i,j,k) = 10.0 ** pred(m) */
#pragma omp parallel for schedule(static)
      for (int64_t mm = 0; mm < (int64_t)npred; ++mm)
        OH_ML[(size_t)(k1 - 1) * nc + mm] = powf(10.0f, pred[mm]);
    }
    if (!out->pred) free(pred);
    if (!out->X) free(X);
    if (rc) goto done;
  }
  /* :1569-1595 */
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < (int64_t)n3; ++e) {
    int c = (int)(e % nc);
    float ml = OH_ML[e] * in->ohscale;
    if (out->OH_boost) out->OH_boost[e] = ml;
    float oh = (PL_MOD[e] > in->TROPP[c]) ? ml : in->OH_CLIM[e];
    out->OH[e] = (oh * NDWET[e]) * 1.0e-6f;
    if (out->NDWET) out->NDWET[e] = NDWET[e];
  }
done:
  free(PL_MOD), free(NDWET), free(PL_BST), free(aod), free(wdn), free(idn), free(iup), free(wup);
  free(aup), free(adn), free(lat), free(so3), free(sza), free(OH_ML);
  return rc;
}
