/* qc_oracle — CPU restatement of the reference's OH prediction path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (quickchem_b200/, libqcoh.so) may
 * include, link or call this.  Allowed callers: tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / `--impl reference` legs (as the checker / timed CPU baseline).
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path
 *  and its arithmetic lives in libxgboost 1.6.0 (EXACT pin,
 * /root/reference/Shared/CMakeLists.txt:8), which is neither vendored nor installed here,
 * and there is no Fortran compiler to build the reference itself.  This oracle restates
 *   - XGBoost 1.6.0's published semantics (legacy-binary model format, DenseAdapter ->
 *     SparsePage missing filter, CPUPredictor block-of-64 traversal and float32 summation)
 *   - the reference's own call sites: predict_OH_with_XGB (OH_GridComp/OH_GridCompMod.F90:123-398)
 *     and Run1's feature assembly / export transform (OH_GridCompMod.F90:1232-1599, :401-466,
 *     :1905-1971)
 * and is anchored by hand-derivable known-answer boosters plus an independent numpy
 * traverser (oracle/naive.py); see tests/test_oracle_*.py.
 */
#ifndef QC_ORACLE_H
#define QC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* RegTree::Node, 20 bytes on disk (xgboost include/xgboost/tree_model.h). */
typedef struct {
  int32_t parent;   /* high bit: is-left-child flag; -1 for the root */
  int32_t cleft;    /* -1 => leaf */
  int32_t cright;
  uint32_t sindex;  /* bit31: default_left; low 31 bits: split feature */
  float info;       /* split_cond, or leaf_value at a leaf */
} orc_node;

typedef struct {
  float loss_chg, sum_hess, base_weight;
  int32_t leaf_child_cnt;
} orc_stat;

typedef struct {
  int32_t num_nodes;
  orc_node *nodes;
  orc_stat *stats;
} orc_tree;

typedef struct {
  float base_score;
  uint32_t num_feature;
  uint32_t major_version, minor_version;
  int32_t num_trees;
  orc_tree *trees;
  int32_t *tree_info;
  char objective[64];
  char booster[32];
} orc_model;

typedef struct {
  uint64_t nrow, ncol;
  uint64_t *offset; /* nrow+1 */
  uint32_t *index;  /* nnz */
  float *value;     /* nnz */
} orc_dmatrix;

const char *orc_last_error(void);
int orc_num_threads(void); /* OpenMP threads the predictor uses */
void orc_set_num_threads(int n); /* override OMP_NUM_THREADS (torchrun sets it to 1) */

/* XGBoosterLoadModel restatement, legacy binary ("binf" optional) only. */
orc_model *orc_model_load(const char *path);
void orc_model_free(orc_model *m);

/* XGDMatrixCreateFromMat restatement: row-major data[nrow][ncol]; entries that are NaN or
 * == missing are dropped; +-inf with a finite `missing` is an error (returns NULL). */
orc_dmatrix *orc_dmatrix_from_mat(const float *data, uint64_t nrow, uint64_t ncol, float missing);
void orc_dmatrix_free(orc_dmatrix *d);

/* XGBoosterPredict restatement. option_mask: 0 value, 1 margin, 2 leaf index.
 * out must hold nrow floats (mask 0/1) or nrow*ntree_used floats (mask 2).
 * Returns number of floats written, or 0 on error. */
uint64_t orc_predict(const orc_model *m, const orc_dmatrix *d, int option_mask, unsigned ntree_limit,
                     float *out);

/* ---- Run1 restatement ------------------------------------------------------------- */
typedef struct {
  int ncol, km;            /* columns (im*jm, i fastest) and levels */
  /* MAPL constants (external to the reference; passed in) */
  float mapl_epsilon, mapl_avogad, mapl_runiv, mapl_radians_to_degrees, mapl_degrees_to_radians;
  float ohscale;           /* OH_instance_OH.rc:42 */
  int compute_once_per_day;/* => dynamic_k_range = !compute_once_per_day (:1561) */
  float tropp_min;         /* 4000 Pa (:1563) */
  int nymd;                /* yyyymmdd for JulianDay (:1481) */
  float missing;           /* -999.0 (:213) */
  /* model state (current values), OH_GridCompMod.F90:1233-1257 */
  const float *T_MOD, *Q_MOD, *PLE_MOD /*[km+1][ncol]*/, *TROPP /*[ncol]*/;
  /* values handed to boost (selected per OH_data_source, :1326-1436) */
  const float *T_BST, *Q_BST, *PLE_BST /*[km+1][ncol]*/, *ZLE_BST /*[km+1][ncol]*/;
  const float *TAUCLW, *TAUCLI, *FCLD, *CH4, *CO;
  const float *SCA[7];     /* BC OC BR DU SU SS NI scattering coefficients, [km][ncol] */
  const float *NO2, *O3, *ISOP, *ACET, *C2H6, *C3H8, *PRPE, *ALK4, *MP, *H2O2, *CH2O;
  const float *GMITO3, *GMITTO3, *ALBUV, *LATS, *LONS; /* [ncol] */
  const float *OH_CLIM;    /* oh_OH default, [km][ncol] */
} orc_run1_in;

typedef struct {
  float *OH;        /* [km][ncol] molec/cm3 (internal state OH) */
  float *OH_boost;  /* [km][ncol] mol/mol, OH_ML after scaling */
  float *X;         /* optional [ncol*ksub][27] packed feature matrix (may be NULL) */
  float *feat3d[27];/* optional per-feature [km][ncol] (3-D) or [ncol] (2-D) dumps; NULL to skip */
  float *NDWET;     /* optional [km][ncol] */
  float *pred;      /* optional raw booster output [ncol*ksub] */
  int k1;           /* out: first predicted level, 1-based (k2 = km) */
} orc_run1_out;

int orc_julian_day(int nymd);
void orc_noon_sza(int jday, const float *lat_rad, const float *lon_rad, int n, float r2d, float d2r,
                  float *sza_deg);
/* One alarmed, need_to_call_BOOST=true pass of Run1.  Returns 0, or -1 with orc_last_error(). */
int orc_run1(const orc_model *m, const orc_run1_in *in, orc_run1_out *out);

#ifdef __cplusplus
}
#endif
#endif
