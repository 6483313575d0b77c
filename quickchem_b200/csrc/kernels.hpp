// kernels.hpp — launch interface of the sm_100a kernels (kernels.cu), used by capi_xgb.cpp / capi_oh.cpp.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace qcoh {

// Device copy of a FlatForest (forest.hpp).
struct DeviceForest {
  const uint2 *nodes = nullptr;          // {value bits, meta}
  const uint32_t *tree_offset = nullptr; // [ntree + 1]
  const int32_t *tree_depth = nullptr;   // [ntree]
  const int32_t *orig_id = nullptr;      // [num_nodes]
  cudaTextureObject_t tex = 0;           // the same nodes as a 1-D linear uint2 texture (TEX pipe)
  int const_top_levels = 0;              // levels of every tree currently held in the constant-memory table
  const uint4 *recs = nullptr;           // two-level records (forest.hpp DuoForest), or nullptr
  cudaTextureObject_t tex4 = 0;          // the records as a 1-D linear uint4 texture
  int duo_ready = 0;                     // the constant-memory tables hold trees [const_tree0, const_tree0 + const_ntree) of this booster's records
  int duo_has_dl = 0;                    // the records carry the default-direction bits (DuoForest::has_default_bits)
  uint32_t duo_blk_mul = 0;              // 1 << (32 - DuoForest::blk_shift): blk = mulhi(w3, duo_blk_mul)
  int32_t const_tree0 = 0, const_ntree = 0;
  int32_t ntree = 0;
  int32_t nfeat = 0;
  int32_t max_depth = 0;
  int64_t num_nodes = 0;
  int64_t sum_depth = 0;                 // sum over trees of the deepest leaf's depth
  int duo_shallow = 0;                   // records average < kShallowRecBytesPerTree per tree (picks the launch shape)
  float base_score = 0.f;
};

// How many trees the constant-memory tables hold at once (4 levels x 16 entries per tree in 61 440 B), and how
// many of them one launch walks.  Measured (profiles/README.md "trees per launch"): warps drift apart through a
// forest, and when a launch walks hundreds of trees their tops and shallow records no longer share the constant
// cache and L1 — 500 trees x depth 10 take 75 ms in one 480-tree launch and 25 ms in launches of 120 (the tile
// stream is re-read per launch: 0.23 ms of the 5 ms a 120-tree range takes at C180).
constexpr int kConstTreesMax = 480;
constexpr int kRangeTrees = 120;
// A forest whose records average less than this per tree (depth <= ~7) is "shallow": 4 trees in flight and 6
// resident CTAs instead of 6 and 5 (100 trees x depth 6: 3.2 ms against 8.0 ms).
constexpr int64_t kShallowRecBytesPerTree = 4096;

// The device form of a DMatrix: order-preserving integer KEYS (kernels.cu) in feature-major tiles of 256 rows,
// Xt[tile][1 + col][256] — what a CTA's transposed shared-memory tile holds, so that a tile arrives with one bulk
// async copy (TMA) and needs no staging arithmetic.  Missing entries (NaN or == missing) are key 0xFFFFFFFF.
// Word-row 0 of a tile is its ROW ORDER: within a tile the rows without a missing entry come first (stable), the
// rows with one last; word p of the order row = original row index in the tile | has_missing << 8.  A predict launch
// on a matrix with missing entries reads it, so that a warp whose 32 rows are all clean walks without the
// default-direction test: what missing entries cost is proportional to the rows that have them.  (A clean matrix has
// the identity order and its launches skip that word-row.)
// Built by seal_tiles from the row-major float matrix, like libxgboost builds its SparsePage in
// XGDMatrixCreateFromMat (OH_GridCompMod.F90:347).
constexpr int kTileRows = 256;
constexpr uint64_t kMaxMatrixCols = 192;  // seal stages a 256 x ncol tile in shared memory; boosters take <= 31 features anyway
inline uint64_t tile_count(uint64_t nrow) { return (nrow + kTileRows - 1) / kTileRows; }
inline size_t tile_words(uint64_t nrow, uint64_t ncol) { return (size_t)tile_count(nrow) * (size_t)(ncol + 1) * kTileRows; }
// words from the start of Xt to the tile that holds row `row0` (a multiple of 256)
inline size_t tile_offset_words(uint64_t row0, uint64_t ncol) { return (size_t)(row0 / kTileRows) * (size_t)(ncol + 1) * kTileRows; }

struct PredictArgs {
  const uint32_t *Xt = nullptr;  // key tiles, device
  uint64_t nrow = 0;
  int32_t ncol = 0;
  int has_missing = 1;       // 0: the sealed matrix holds no NaN / == missing entry
  int pred_leaf = 0;         // option_mask & 2
  int32_t tree_begin = 0;    // this launch walks trees [tree_begin, tree_end) ...
  int32_t tree_end = 0;
  int32_t out_stride = 0;    // pred_leaf: floats per row in `out` (= trees used by the whole call)
  int first = 1;             // sums: 1 = start from base_score, 0 = continue from the partial sum in out[row]
  int last = 1;              // sums: 1 = apply the export transform, 0 = store the partial sum
  int exp10 = 0;             // fused export transform, OH_GridCompMod.F90:369,1569
  float scale = 1.f;
  float *out = nullptr;      // device: [nrow] or [nrow][out_stride]
};

struct Tunables {
  int variant = 0;   // 0 = default
  int ilp = 0;       // trees walked concurrently per thread (0 = default)
  int block = 0;     // threads per CTA (0 = default)
  int top_levels = -1;  // tree levels served from constant memory: -1 = default (4), 0 = none
  int park = -1;     // -1 = default (on)
  int minb = 0;      // min resident CTAs per SM the kernel is compiled for (register budget)
  int duo = -1;      // two-level records: -1 = default, 0 = off, 1 = on
  int duo_mask = 0;  // (experiment) texture-pipe tree mask of the two-level kernel, 0 = default
  int persist = -1;  // persistent double-buffered tile loop: 1 = on (default off, see kernels.cu)
  int range_trees = 0;  // trees per launch (<= kConstTreesMax): 0 = default
};

constexpr bool kDuoDefault = true;  // two-level records by default when the booster qualifies

uint64_t launch_count();
// launches per kernel family since load, and the family that served the last predict launch
// ("duo", "duo_missing", "duo_leaf", "duo_missing_leaf", "nodes8", "nodes8_missing", "nodes8_leaf", ...)
uint64_t kernel_launches(const char *family);
const char *last_predict_kernel();

// (experiment) copy levels 0..levels-1 of every tree into the kernel's __constant__ table
cudaError_t upload_const_top(const uint32_t *dev_nodes_xy, const uint32_t *tree_offset, int ntree, int levels, cudaStream_t s);
// two-level layout: DuoForest::top_xy into the tree-top table and DuoForest::tree_slot into its base table
cudaError_t upload_const_duo(const uint32_t *top_xy, const uint32_t *tree_slot, int ntree, cudaStream_t s);

// Row-major float matrix -> key tiles.  flags[0] |= 1 if any entry is NaN or == missing; flags[0] |= 2 if any
// entry is +-inf (and `missing` is finite) — what XGDMatrixCreateFromMat checks (xgboost src/data/data.cc).
// X and Xt cover `nrow` rows starting at a tile boundary.
cudaError_t launch_seal_tiles(const float *X, uint64_t nrow, int ncol, float missing, uint32_t *Xt, int *flags, cudaStream_t s);

cudaError_t launch_predict(const DeviceForest &f, const PredictArgs &a, const Tunables &t, cudaStream_t s);

// Fused Run1 predict: features come straight from the SoA fields (27 sources in feature order)
struct SoaArgs {
  const float *src3[27];     // [km][ncol] source of feature f, or nullptr if it is a 2-D field
  const float *src2[27];     // [ncol] source of a 2-D feature
  const float *ple = nullptr;  // PLE_BST [km+1][ncol] (feature 1 is computed from it)
  int32_t ncol = 0;
  uint64_t e0 = 0;           // cell index of slab row 0: (k1 - 1) * ncol
  uint64_t nrow = 0;         // ncol * ksub
  float missing = -999.f;
  int32_t ntree_used = 0;
  int exp10 = 1;
  float scale = 1.f;
  float *out = nullptr;      // OH_ML slab
  float *pred = nullptr;     // optional raw booster output
  int *flags = nullptr;      // |= 2 on +-inf input
};
cudaError_t launch_predict_soa(const DeviceForest &f, const SoaArgs &a, const Tunables &t, cudaStream_t s);

// ---- fused Run1 pieces (OH_GridCompMod.F90:1232-1599) ---------------------------------
struct Run1Dev {
  int32_t ncol = 0, km = 0;
  float eps = 0, avogad = 0, runiv = 0, r2d = 0;
  float ohscale = 1.f, tropp_min = 4000.f, missing = -999.f;
  int dynamic_k = 0;
  const float *T_MOD, *Q_MOD, *PLE_MOD, *TROPP;
  const float *T_BST, *Q_BST, *PLE_BST, *ZLE_BST;
  const float *TAUCLW, *TAUCLI, *FCLD, *CH4, *CO;
  const float *SCA[7];
  const float *NO2, *O3, *ISOP, *ACET, *C2H6, *C3H8, *PRPE, *ALK4, *MP, *H2O2, *CH2O;
  const float *GMITO3, *GMITTO3, *ALBUV, *LATS;
  const float *SZA;      // [ncol] degrees (host-computed with libm, see oh_host.cpp)
  const float *OH_CLIM;
  const float *AREA;
  // work / outputs (device)
  float *PL_MOD, *NDWET;             // [km][ncol]
  float *sums[6];                    // wdn idn iup wup aup adn, [km][ncol]
  float *lat_deg, *so3;              // [ncol] latarr, stratO3 (written by oh_sums)
  float *aod, *pl_bst;               // [km][ncol] aod (:1451-1466) and PL_BST (:1488), written by oh_sums
  float *OH_ML;                      // persistent [km][ncol]
  float *OH, *OH_boost;              // [km][ncol]
  float *LOSS_CH4, *LOSS_CO;         // optional [km][ncol] 1/s (build-defined)
  int *ctl;                          // [0] ksub (atomicMax) [1] tropp<=tropp_min count [2] matrix flags
  double *diag;                      // [4]
};

cudaError_t launch_oh_state(const Run1Dev &r, cudaStream_t s);
cudaError_t launch_oh_sums(const Run1Dev &r, cudaStream_t s);
cudaError_t launch_oh_pack(const Run1Dev &r, int k1, float *X, cudaStream_t s);
cudaError_t launch_oh_finalize(const Run1Dev &r, cudaStream_t s);
cudaError_t launch_oh_diag(const Run1Dev &r, cudaStream_t s);
cudaError_t launch_fill(float *p, uint64_t n, float v, cudaStream_t s);

}  // namespace qcoh
