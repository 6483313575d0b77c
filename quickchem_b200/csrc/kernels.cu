// kernels.cu — hand-written sm_100a kernels of the OH hot path.
//
//   K3  seal_tiles_kernel     the missing / inf scan of XGDMatrixCreateFromMat (OH_GridCompMod.F90:347) fused with
//                             the conversion of the row-major float matrix into the DMatrix's device form:
//                             feature-major tiles of order-preserving integer keys
//   K2b predict_tiles_kernel  tree-ensemble traversal on two-level 16-byte records (one gather per two tree
//                             levels), replaces libxgboost's CPUPredictor behind XGBoosterPredict (reference call
//                             site :356) with the export transform 10**x * OHscale (:369,:1569) fused as epilogue;
//                             sums / leaf indices, clean matrices / missing entries; tiles arrive by TMA, optionally
//                             in a persistent double-buffered loop (shallow forests: the HBM-bound regime)
//   K2  predict_tiles_nodes8_kernel   the same on the 8-byte depth-ordered nodes (boosters the records do not hold)
//       predict_soa_kernel    K2 / K2b with the tile assembled straight from the Run1 SoA fields
//   K1  oh_state / oh_sums / oh_pack    Run1 feature assembly (:1240-1257, :1441-1488, :303-345)
//   K5  oh_finalize           troposphere mask + unit conversion (:1579-1595)
//   K4  oh_diag               build-defined mass-weighted mean OH / CH4 lifetime partial sums
//
// Numerics: compiled with -fmad=false and written with explicit round-to-nearest intrinsics where
// the Fortran evaluation order matters, so every float32 feature is bit-identical to the CPU
// evaluation.  No tensor cores: tree traversal is not a dense contraction.
#include "kernels.hpp"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "forest.hpp"

namespace qcoh {

static thread_local uint64_t g_launches = 0;
uint64_t launch_count() { return g_launches; }

#define QC_LAUNCHED() (++g_launches, cudaGetLastError())

// launches per kernel family: what the parity tests assert ("this launch was served by the two-level kernel")
namespace {
struct FamilyCount {
  char name[32];
  uint64_t n;
};
thread_local FamilyCount g_family[24];
thread_local int g_nfamily = 0;
thread_local const char *g_last_predict = "";
const char *count_family(const char *base, bool hm, bool pl) {
  char name[32];
  snprintf(name, sizeof name, "%s%s%s", base, hm ? "_missing" : "", pl ? "_leaf" : "");
  for (int i = 0; i < g_nfamily; ++i)
    if (!strcmp(g_family[i].name, name)) return ++g_family[i].n, g_family[i].name;
  if (g_nfamily == 24) return "";
  strcpy(g_family[g_nfamily].name, name);
  g_family[g_nfamily].n = 1;
  return g_family[g_nfamily++].name;
}
}  // namespace
uint64_t kernel_launches(const char *family) {
  for (int i = 0; i < g_nfamily; ++i)
    if (!strcmp(g_family[i].name, family)) return g_family[i].n;
  return 0;
}
const char *last_predict_kernel() { return g_last_predict; }

__global__ void fill_kernel(float *p, uint64_t n, float v) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}
cudaError_t launch_fill(float *p, uint64_t n, float v, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  fill_kernel<<<blocks, 256, 0, s>>>(p, n, v);
  return QC_LAUNCHED();
}

// =====================================================================================
// Keys, tiles, constant tables
// =====================================================================================
// One thread = one row (grid cell).  A CTA's 256 rows sit in shared memory TRANSPOSED, skey[f][tid]:
// whatever feature each lane asks for, lane L always hits bank L — the data-dependent feature fetch is
// bank-conflict-free by construction.  The tile holds order-preserving integer keys of the values
// (below); slot `nfeat` of every row holds key 0: leaves are encoded with feat = nfeat, so the compare
// never says "right" there — no leaf test inside the descent.  Missing entries (NaN or == missing) are
// key 0xFFFFFFFF.
//
// XGBoost semantics restated (xgboost 1.6.0 src/predictor/predict_fn.h GetNextNode,
// src/predictor/cpu_predictor.cc PredictByAllTrees): missing -> default child, else
// left + !(fvalue < split_cond); out = base_score, then += leaf value tree by tree in float32.
constexpr int kBlock = kTileRows;  // rows (= threads) per CTA
constexpr int kStride = kTileRows; // feature stride of the transposed tile, in keys (1 KB: the PRMT address trick)
static_assert(kStride * 4 == 1024, "the address arithmetic of the walks assumes a 1 KB feature stride");

// Order-preserving integer keys: key(v) = bits ^ (sign ? 0xFFFFFFFF : 0x80000000) of the value with -0.0
// folded into +0.0, which is monotone: a < b  <=>  key(a) < key(b) (unsigned) for all non-NaN floats, and
// key(-0) == key(+0) like the float compare.  An internal node stores x = -key(threshold) (mod 2^32;
// key(thr) is never 0), so
//     !(v < thr)  <=>  key(thr) <= key(v)  <=>  x + key(v) >= 2^32
// which is the carry of x + key(v): `add.cc` + `addc` fold the compare into the index update in two
// instructions and no predicate.  Missing is key 0xFFFFFFFF (above +inf: it carries against every
// threshold, and the walks that may see it test for it explicitly and take the default child).
constexpr uint32_t kKeyMissing = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t float_key(float v) {
  const uint32_t b = __float_as_uint(__fadd_rn(v, 0.0f));  // -0.0 + 0.0 = +0.0
  return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
// idx + rel + carry(x + kv)
__device__ __forceinline__ uint32_t step_index(uint32_t idx, uint32_t rel, uint32_t x, uint32_t kv) {
  asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %2, %3;\n\taddc.u32 %0, %0, %1;\n\t}" : "+r"(idx) : "r"(rel), "r"(x), "r"(kv));
  return idx;
}
__device__ __forceinline__ uint32_t add_carry_out(uint32_t a, uint32_t b) {
  uint32_t c;
  asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, 0, 0;\n\t}" : "=r"(c) : "r"(a), "r"(b));
  return c;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}

// Top of every tree in constant memory: constant loads go through the constant cache, not the LSU / TEX
// data pipes; lanes of a warp mostly agree at those levels, so the per-address serialisation of divergent
// constant loads stays short.  The table belongs to one booster (and one range of its trees) at a time:
//   8-byte-node walk: the first 2^CTOP - 1 nodes of a tree in breadth-first order = its levels 0..CTOP-1
//   two-level records: DuoForest::top_xy (complete heap-ordered tops) + c_duo_base, the global index of each
//                      tree's first record (the block pointers are tree-relative)
constexpr int kConstTopNodes = kConstTreesMax << kDuoTop;  // 61 440 B of the 64 KB constant bank
__constant__ uint2 c_top[kConstTopNodes];
__constant__ uint32_t c_duo_base[kConstTreesMax];

cudaError_t upload_const_top(const uint32_t *dev_nodes_xy, const uint32_t *tree_offset, int ntree, int levels, cudaStream_t s) {
  const int stride = 1 << levels;
  if (levels <= 0 || (int64_t)ntree * stride > kConstTopNodes) return cudaErrorInvalidValue;
  std::vector<uint2> host((size_t)ntree * stride);
  for (int t = 0; t < ntree; ++t) {
    const uint32_t n0 = tree_offset[t], n1 = tree_offset[t + 1];
    for (int i = 0; i < stride; ++i) {
      const uint32_t n = n0 + (uint32_t)i;
      host[t * stride + i] = n < n1 ? make_uint2(dev_nodes_xy[2 * n], dev_nodes_xy[2 * n + 1]) : make_uint2(0u, 0u);
    }
  }
  // pageable source: the call returns once the data is staged, so the local buffer may go
  return cudaMemcpyToSymbolAsync(c_top, host.data(), sizeof(uint2) * (size_t)ntree * stride, 0, cudaMemcpyHostToDevice, s);
}

cudaError_t upload_const_duo(const uint32_t *top_xy, const uint32_t *tree_slot, int ntree, cudaStream_t s) {
  if (ntree > kConstTreesMax) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemcpyToSymbolAsync(c_top, top_xy, sizeof(uint2) * ((size_t)ntree << kDuoTop), 0, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbolAsync(c_duo_base, tree_slot, sizeof(uint32_t) * (size_t)ntree, 0, cudaMemcpyHostToDevice, s);
}

// ---- TMA (bulk async copy) + mbarrier helpers: global -> shared without the LSU ------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// The tile loop shared by the predict kernels.  A CTA walks tiles blockIdx.x, + gridDim.x, ...: with a grid of
// one CTA per tile that is the classic launch (the resident CTAs of an SM overlap each other's loads); with a
// persistent grid and NBUF = 2 the next tile's bulk copy (TMA, cp.async.bulk completing on an mbarrier —
// SASS UBLKCP + SYNCS) is in flight while this one is walked.  A tile is ncol KB of keys in its final
// layout: no LSU instruction, no register, no arithmetic between HBM and the walk.  One thread issues the
// copy and polls the mbarrier, the CTA then passes a barrier (256 threads polling one word showed up as
// shared-memory bank-conflict wavefronts on the LSU data pipe, profiles/README.md v4).
template <int NBUF>
struct TilePipe {
  uint32_t smem0, bar0, buf_bytes, tile_bytes, src_skip, dst_skip;
  const uint32_t *Xt;
  uint64_t ntile, tile, tile_stride;
  uint32_t it;
  __device__ __forceinline__ void issue(uint64_t t, int b) const {
    mbar_expect_tx(bar0 + 8u * b, tile_bytes);
    if (tile_bytes) bulk_g2s(smem0 + buf_bytes * b + dst_skip, Xt + t * tile_stride + src_skip, tile_bytes, bar0 + 8u * b);
  }
  // A buffer is [row order | nfeat feature slots | key-0 slot], 1 KB each.  Slots the matrix does not have are
  // missing (xgboost FVec::Fill leaves them flagged); the last slot is the key-0 slot leaves point at.  Both are
  // written once per buffer: the bulk copies only touch the order slot (with_order) and slots 0..ncol-1.
  __device__ __forceinline__ void begin(uint32_t *smem, unsigned long long *bars, const PredictArgs &a, int nfeat, bool with_order, int tid) {
    smem0 = (uint32_t)__cvta_generic_to_shared(smem), bar0 = (uint32_t)__cvta_generic_to_shared(bars);
    buf_bytes = (uint32_t)(nfeat + 2) * kStride * 4u;
    tile_stride = (uint64_t)(a.ncol + 1) * kStride;
    tile_bytes = (uint32_t)(a.ncol + (with_order ? 1 : 0)) * kStride * 4u;
    src_skip = with_order ? 0u : (uint32_t)kStride, dst_skip = with_order ? 0u : (uint32_t)kStride * 4u;
    Xt = a.Xt, ntile = (a.nrow + kTileRows - 1) / kTileRows, tile = blockIdx.x, it = 0;
#pragma unroll
    for (int b = 0; b < NBUF; ++b) {
      uint32_t *k = smem + (size_t)b * (nfeat + 2) * kStride + kStride;
      for (int c = a.ncol; c < nfeat; ++c) k[c * kStride + tid] = kKeyMissing;
      k[nfeat * kStride + tid] = 0u;
    }
    if (tid == 0) {
#pragma unroll
      for (int b = 0; b < NBUF; ++b) mbar_init(bar0 + 8u * b, 1);
      mbar_init_fence();
    }
    __syncthreads();
    if (tid == 0 && tile < ntile) issue(tile, 0);
  }
  __device__ __forceinline__ bool more() const { return tile < ntile; }
  // returns the shared address of this iteration's buffer (its row-order slot; the feature slots follow 1 KB later)
  __device__ __forceinline__ uint32_t acquire(int tid) {
    const int cur = NBUF == 2 ? (int)(it & 1u) : 0;
    if (tid == 0) {
      if (NBUF == 2 && tile + gridDim.x < ntile) issue(tile + gridDim.x, cur ^ 1);
      mbar_wait(bar0 + 8u * cur, (it / NBUF) & 1u);
    }
    __syncthreads();
    return smem0 + buf_bytes * cur;
  }
  __device__ __forceinline__ void release(int tid) {
    // every thread is done with this buffer before it is refilled
    if (NBUF == 2) {
      if (tile + gridDim.x < ntile) __syncthreads();
    } else if (tile + gridDim.x < ntile) {
      __syncthreads();
      if (tid == 0) issue(tile + gridDim.x, 0);
    }
    tile += gridDim.x, ++it;
  }
};

// =====================================================================================
// K3 — seal: row-major float matrix -> key tiles, with the missing / inf scan
// =====================================================================================
__global__ void __launch_bounds__(kTileRows) seal_tiles_kernel(const float *__restrict__ X, uint64_t nrow, int ncol, float missing,
                                                               int check_inf, uint32_t *__restrict__ Xt, int *flags) {
  extern __shared__ __align__(16) float srows[];  // [256][ld] row-major, ld odd: the per-thread row reads are conflict-free
  __shared__ int warp_clean[kTileRows / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int ld = ncol | 1;
  const uint64_t r0 = (uint64_t)blockIdx.x * kTileRows;
  const uint64_t left = nrow - r0;
  const int nr = left < (uint64_t)kTileRows ? (int)left : kTileRows;
  const float *__restrict__ src = X + r0 * (uint64_t)ncol;
  const int n = nr * ncol;
  if (ld == ncol) {  // the tile's rows are one contiguous run of X: flat, coalesced copy
    if ((((uintptr_t)src) & 15u) == 0u) {
      const float4 *s4 = reinterpret_cast<const float4 *>(src);
      float4 *d4 = reinterpret_cast<float4 *>(srows);
      for (int i = tid; i < (n >> 2); i += kTileRows) d4[i] = __ldg(s4 + i);
      for (int i = (n & ~3) + tid; i < n; i += kTileRows) srows[i] = __ldg(src + i);
    } else {
      for (int i = tid; i < n; i += kTileRows) srows[i] = __ldg(src + i);
    }
  } else {
    for (int i = tid; i < n; i += kTileRows) srows[(i / ncol) * ld + (i % ncol)] = __ldg(src + i);
  }
  __syncthreads();
  // pass 1: does my row hold a missing entry?  (rows past the end of the matrix count as clean; never stored)
  int fl = 0;
  if (tid < nr) {
    for (int c = 0; c < ncol; ++c) {
      const float x = srows[tid * ld + c];
      if (x != x || x == missing) fl |= 1;
      if (check_inf && isinf(x)) fl |= 2;
    }
  }
  // row order: clean rows first, rows with a missing entry last, both in their original order (stable)
  const unsigned clean_mask = __ballot_sync(0xffffffffu, !(fl & 1));
  if (lane == 0) warp_clean[wid] = __popc(clean_mask);
  __syncthreads();
  int clean_before = 0, clean_total = 0;
#pragma unroll
  for (int w = 0; w < kTileRows / 32; ++w) {
    const int cw = warp_clean[w];
    clean_total += cw;
    if (w < wid) clean_before += cw;
  }
  const int my_clean_rank = clean_before + __popc(clean_mask & ((1u << lane) - 1u));
  const int pos = (fl & 1) ? clean_total + (tid - my_clean_rank) : my_clean_rank;
  uint32_t *__restrict__ tile = Xt + (size_t)blockIdx.x * (size_t)(ncol + 1) * kTileRows;
  tile[pos] = (uint32_t)tid | ((uint32_t)(fl & 1) << 8);
  // pass 2: my row's keys, column by column, to my position (1 KB per column; coalesced where the order is the identity)
  uint32_t *__restrict__ dst = tile + kTileRows + pos;
  for (int c = 0; c < ncol; ++c) {
    uint32_t k = 0u;
    if (tid < nr) {
      const float x = srows[tid * ld + c];
      k = (x != x || x == missing) ? kKeyMissing : float_key(x);
    }
    dst[(size_t)c * kTileRows] = k;
  }
  fl = __reduce_or_sync(0xffffffffu, fl);
  if (lane == 0 && fl) atomicOr(flags, fl);
}

cudaError_t launch_seal_tiles(const float *X, uint64_t nrow, int ncol, float missing, uint32_t *Xt, int *flags, cudaStream_t s) {
  if (nrow == 0) return cudaSuccess;
  const uint64_t ntile = tile_count(nrow);
  if (ntile > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  const size_t smem = (size_t)kTileRows * (size_t)(ncol | 1) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(seal_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // xgboost src/data/data.cc: `!std::isinf(missing) && std::isinf(value)` invalidates the input
  seal_tiles_kernel<<<(unsigned)ntile, kTileRows, smem, s>>>(X, nrow, ncol, missing, std::isinf(missing) ? 0 : 1, Xt, flags);
  count_family("seal_tiles", false, false);
  return QC_LAUNCHED();
}

// =====================================================================================
// K2 — walk on the 8-byte depth-ordered nodes
// =====================================================================================
// TEXMODE is a bit mask over the ILP trees in flight: tree j fetches its nodes through the texture pipe
// (tex1Dfetch on the same buffer) if bit j is set, through the LSU (LDG) otherwise.  ctree0: the first tree
// the constant table holds.
template <int ILP, bool HAS_MISSING, bool PARK, int TEXMODE = 0, int CTOP = 0>
__device__ __forceinline__ void walk_group(const uint2 *__restrict__ nodes, cudaTextureObject_t tex, const uint32_t *__restrict__ toff,
                                           const int32_t *__restrict__ tdepth, int t, int ctree0, uint32_t my_saddr,
                                           uint32_t (&idx)[ILP], uint32_t (&xbits)[ILP]) {
  int depth = 0, minleaf = 255;  // tdepth[] = deepest leaf | shallowest leaf << 8
  uint2 nd[ILP];
  uint32_t rel[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) {
    idx[j] = __ldg(toff + t + j);
    const int dd = __ldg(tdepth + t + j);
    depth = max(depth, dd & 0xFF);
    minleaf = min(minleaf, dd >> 8);
    nd[j] = make_uint2(0u, 0u);
    rel[j] = 1u;
  }
  // depth + 1 fetches per tree: the last one reads the leaf itself, whose x word is the leaf value.
  // PARK: a lane that has reached its leaf (rel == 0) stops fetching — it would otherwise re-read its
  // own leaf line at every remaining level, and at the deep levels those are up to 32 different
  // lines per warp request; the L1TEX data pipe is the bound of this kernel.  The
  // fetch is predicated, so nd[j] keeps the leaf node and no per-level copy of the value is needed.
  auto visit = [&](int j) {
    // shared address of skey[feat][tid] = my_saddr + feat * 1024: the top byte of the meta word
    // is feat * 4, so moving it to byte 1 (one PRMT) gives feat * 1024
    static_assert(kMetaFeatShift == 26, "address trick assumes feat in the top 6 bits");
    const uint32_t kv = lds_u32(my_saddr + __byte_perm(nd[j].y, 0u, 0x4434));
    rel[j] = nd[j].y & kMetaRelMask;
    if (HAS_MISSING && kv == kKeyMissing)  // default child: left = idx + rel, right = left + 1
      idx[j] += rel[j] + ((nd[j].y & kMetaDefaultLeftBit) ? 0u : (rel[j] != 0u ? 1u : 0u));
    else
      idx[j] = step_index(idx[j], rel[j], nd[j].x, kv);
  };
  if (CTOP > 0) {
    uint32_t cbase[ILP];  // index of this tree's table row minus its first node
#pragma unroll
    for (int j = 0; j < ILP; ++j) cbase[j] = ((uint32_t)(t + j - ctree0) << CTOP) - idx[j];
    if (minleaf >= CTOP) {
      // no leaf above level CTOP in any of these trees (the usual case for deep trees): every lane walks
      // all CTOP levels — no park predicate, no branches
#pragma unroll
      for (int d = 0; d < CTOP; ++d) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
          nd[j] = c_top[cbase[j] + idx[j]];
          visit(j);
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < CTOP; ++d) {
        if (d <= depth) {
#pragma unroll
          for (int j = 0; j < ILP; ++j) {
            if (!PARK || rel[j] != 0u) {
              nd[j] = c_top[cbase[j] + idx[j]];
              visit(j);
            }
            if (PARK) asm volatile("" : "+r"(rel[j]));
          }
        }
      }
    }
  }
  auto fetch = [&](int j) {
    if ((TEXMODE >> j) & 1)
      nd[j] = tex1Dfetch<uint2>(tex, (int)idx[j]);
    else
      nd[j] = __ldg(nodes + idx[j]);
  };
  for (int d = CTOP; d < depth; ++d) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      if (!PARK || rel[j] != 0u) {
        fetch(j);
        visit(j);
      }
      // keep "still walking" in the rel register only (one ISETP per level instead of predicate shuffling)
      if (PARK) asm volatile("" : "+r"(rel[j]));
    }
  }
  // level `depth` holds only leaves: the last fetch reads the leaf a lane stands on — its value is all that is needed
  if (depth >= CTOP) {
#pragma unroll
    for (int j = 0; j < ILP; ++j)
      if (!PARK || rel[j] != 0u) fetch(j);
  }
#pragma unroll
  for (int j = 0; j < ILP; ++j) xbits[j] = nd[j].x;
}

// all trees [t0, t1) on the 8-byte nodes, in tree order; emit(t, value bits, node index)
template <int ILP, bool HAS_MISSING, bool PARK, int TEXMODE, int CTOP, class Emit>
__device__ __forceinline__ void forest_walk(const DeviceForest &f, uint32_t my, int t0, int t1, Emit &&emit) {
  int t = t0;
  for (; t + ILP <= t1; t += ILP) {
    uint32_t idx[ILP], xb[ILP];
    walk_group<ILP, HAS_MISSING, PARK, TEXMODE, CTOP>(f.nodes, f.tex, f.tree_offset, f.tree_depth, t, f.const_tree0, my, idx, xb);
#pragma unroll
    for (int j = 0; j < ILP; ++j) emit(t + j, xb[j], idx[j]);
  }
  for (; t < t1; ++t) {
    uint32_t idx[1], xb[1];
    walk_group<1, HAS_MISSING, PARK, 0, 0>(f.nodes, f.tex, f.tree_offset, f.tree_depth, t, 0, my, idx, xb);
    emit(t, xb[0], idx[0]);
  }
}

// =====================================================================================
// K2b — walk on the two-level records (forest.hpp DuoForest)
// =====================================================================================
// Levels 0..3 come from constant memory as complete heap-ordered tops (entry i, children 2i and 2i + 1: no
// branch, no predicate); from level 4 on one 16-byte record {root, left, right thresholds; features; block
// pointer} decides two levels, and the four possible successors are contiguous.  Against the 8-byte nodes
// this halves the dependent gathers of a walk and takes ~40 % of the distinct lines per warp request off the
// L1TEX data pipes (tools/replay_two_level_records.py).
// Default shape (B200, profiles/README.md "two-level records"): 6 trees in flight, 4 of them gathering through
// the texture pipe, 5 resident CTAs per SM — LSU and TEX data pipes, ALU and issue slots all end up at
// 76-86 % busy.
// HAS_MISSING: a missing entry (key 0xFFFFFFFF) takes the node's default child (bits 17..15 of w3 / bit 0 of a
// top entry's y word); the records carry those bits when DeviceForest::duo_has_dl.
constexpr int kDuoIlp = 6, kDuoTexMask = 0x36, kDuoMinBlocks = 5;
template <int ILP, int TEXMODE, bool HAS_MISSING>
__device__ __forceinline__ void walk_group_duo(const uint4 *__restrict__ recs, cudaTextureObject_t tex4,
                                               const int32_t *__restrict__ tdepth, uint32_t blk_mul, int t, int ctree0, uint32_t my_saddr,
                                               uint32_t (&xbits)[ILP], uint32_t (&ids)[ILP]) {
  constexpr int CTOP = kDuoTop;
  int depth = CTOP - 1;  // at least one record (a shallower tree ends in the terminal records of its padded leaves)
  uint32_t idx[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) {
    depth = max(depth, __ldg(tdepth + t + j) & 0xFF);
    idx[j] = 1u;  // heap position in the complete top: children of i are 2i and 2i + 1
  }
  const int tl = t - ctree0;  // row of the constant tables
#pragma unroll
  for (int d = 0; d < CTOP; ++d) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      const uint2 nd = c_top[((uint32_t)(tl + j) << CTOP) + idx[j]];
      const uint32_t kv = lds_u32(my_saddr + __byte_perm(nd.y, 0u, 0x4434));
      if (HAS_MISSING) {
        uint32_t right = add_carry_out(nd.x, kv);
        if (kv == kKeyMissing) right = (nd.y & kTopDefaultLeftBit) ^ 1u;
        idx[j] = 2u * idx[j] + right;
      } else {
        idx[j] = step_index(idx[j], idx[j], nd.x, kv);  // 2i + right
      }
    }
  }
  uint4 r[ILP];
  uint32_t walking[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) {
    idx[j] += c_duo_base[tl + j] - (1u << CTOP);  // heap position 16..31 -> record index
    r[j] = make_uint4(0u, 0u, 0u, 0u);
    walking[j] = 1u;
  }
  // the terminal record of a leaf at depth D is rooted at depth D (D even) or D + 1 (D odd).  Only the
  // gather is predicated (a lane that has reached its terminal record stops fetching and keeps it); the
  // arithmetic below runs unconditionally, phase by phase across the ILP trees, so that the trees'
  // dependent chains interleave: on a terminal record (w3 == 0) it reads feature 0 and leaves walking at 0.
  auto gather = [&](int j) {
    if (walking[j] != 0u) {
      if ((TEXMODE >> j) & 1)
        r[j] = tex1Dfetch<uint4>(tex4, (int)idx[j]);
      else
        r[j] = __ldg(recs + idx[j]);
    }
  };
  // Records rooted at the deepest level a walk of these trees can reach (depth, or depth + 1 for an odd depth) are
  // all terminal: the last iteration is the gather alone — its arithmetic would only confirm w3 == 0.  For the
  // depth-18 booster that is one eighth of the record arithmetic and of its feature fetches; for a depth-6 forest half.
  int d = CTOP;
  for (; d + 2 <= depth + 1; d += 2) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) gather(j);
    uint32_t kv0[ILP], kv[ILP], right1[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      // skey[feat(root)][tid]: feat is the low 5 bits of w3; one AND + one multiply-add
      uint32_t sa;
      asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(sa) : "r"(r[j].w & 31u), "r"((uint32_t)(kStride * 4)), "r"(my_saddr));
      kv0[j] = lds_u32(sa);
    }
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      right1[j] = add_carry_out(r[j].x, kv0[j]);
      if (HAS_MISSING && kv0[j] == kKeyMissing) right1[j] = (r[j].w & kDuoDlRoot) ? 0u : 1u;
      // feat(left) sits at bits 14..10 — already feat * 1024 —, feat(right) at 9..5
      const uint32_t ms = right1[j] ? (r[j].w << 5) : r[j].w;
      kv[j] = lds_u32(my_saddr + (ms & (31u << 10)));
    }
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      const uint32_t xs = right1[j] ? r[j].z : r[j].y;
      // blk = w3 >> blk_shift as a multiply-high: the FMA pipe has room, the ALU pipe (SEL / LOP3 / IADD3) does not
      uint32_t blk;
      asm("mul.hi.u32 %0, %1, %2;" : "=r"(blk) : "r"(r[j].w), "r"(blk_mul));
      if (HAS_MISSING) {
        uint32_t right2 = add_carry_out(xs, kv[j]);
        // a leaf child compares the key-0 slot: never missing.  Default bit of the child that was taken.
        if (kv[j] == kKeyMissing) right2 = (r[j].w & (right1[j] ? kDuoDlRight : kDuoDlLeft)) ? 0u : 1u;
        const uint32_t half = 2u * blk + right1[j];
        idx[j] = c_duo_base[tl + j] + 2u * half + right2;
      } else {
        // next record = base + blk * 4 + 2 * right1 + right2
        // half = 2 * blk + right1, the carry of the root compare folded into a multiply-add (FMA pipe again)
        uint32_t half;
        asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\tmadc.lo.u32 %0, %3, 2, 0;\n\t}" : "=r"(half) : "r"(r[j].x), "r"(kv0[j]), "r"(blk));
        idx[j] = step_index(c_duo_base[tl + j] + half, half, xs, kv[j]);
      }
      walking[j] = blk;  // 0: this was a terminal record, r[j].x is the leaf value
      asm volatile("" : "+r"(walking[j]));
    }
  }
#pragma unroll
  for (int j = 0; j < ILP; ++j) gather(j);
#pragma unroll
  for (int j = 0; j < ILP; ++j) xbits[j] = r[j].x, ids[j] = r[j].y;
}

// all trees [t0, t1) on the two-level records: ILP-wide groups, then half-width groups (one LSU tree, the rest on
// the texture pipe), then one by one — always in tree order (float32 sum order is part of parity).
// emit(t, leaf value bits, XGBoost node id of the leaf)
template <int ILP, int TEXMODE, bool HAS_MISSING, class Emit>
__device__ __forceinline__ void forest_walk_duo(const DeviceForest &f, uint32_t my, int t0, int t1, Emit &&emit) {
  int t = t0;
  for (; t + ILP <= t1; t += ILP) {
    uint32_t xb[ILP], id[ILP];
    walk_group_duo<ILP, TEXMODE, HAS_MISSING>(f.recs, f.tex4, f.tree_depth, f.duo_blk_mul, t, f.const_tree0, my, xb, id);
#pragma unroll
    for (int j = 0; j < ILP; ++j) emit(t + j, xb[j], id[j]);
  }
  constexpr int H = ILP / 2;
  if (H >= 2) {
    for (; t + H <= t1; t += H) {
      uint32_t xb[H > 0 ? H : 1], id[H > 0 ? H : 1];
      walk_group_duo<(H > 0 ? H : 1), ((1 << H) - 2), HAS_MISSING>(f.recs, f.tex4, f.tree_depth, f.duo_blk_mul, t, f.const_tree0, my, xb, id);
#pragma unroll
      for (int j = 0; j < H; ++j) emit(t + j, xb[j], id[j]);
    }
  }
  for (; t < t1; ++t) {
    uint32_t xb[1], id[1];
    walk_group_duo<1, 0, HAS_MISSING>(f.recs, f.tex4, f.tree_depth, f.duo_blk_mul, t, f.const_tree0, my, xb, id);
    emit(t, xb[0], id[0]);
  }
}

__device__ __forceinline__ float export_transform(float acc, int exp10_on, float scale) {
  if (!exp10_on) return acc;
  // OH_ML = 10.0 ** pred (OH_GridCompMod.F90:369), then OH_ML * OHscale (:1569): two float32
  // roundings.  10**x is evaluated in float64 and rounded once (SURVEY.md hard part 7).
  const float p = (float)exp10((double)acc);
  return __fmul_rn(p, scale);
}

// what a row does with its leaves: float32 sum in tree order (+ export transform on the last launch of a
// chunked forest), or the leaf's XGBoost node id per tree (option_mask = 2)
template <bool PRED_LEAF, class Walk>
__device__ __forceinline__ void row_result(const DeviceForest &f, const PredictArgs &a, uint64_t row, bool live, Walk &&walk) {
  if (PRED_LEAF) {
    float *o = a.out + row * (uint64_t)a.out_stride;
    walk([&](int t, uint32_t, uint32_t id) {
      if (live) o[t] = (float)id;
    });
  } else {
    float acc = f.base_score;
    if (!a.first && live) acc = a.out[row];
    walk([&](int, uint32_t xb, uint32_t) { acc = __fadd_rn(acc, __uint_as_float(xb)); });
    if (live) a.out[row] = a.last ? export_transform(acc, a.exp10, a.scale) : acc;
  }
}

// A launch on a matrix with missing entries reads the tile's row order (kernels.hpp): thread tid walks the row at
// position tid — clean rows first — and a warp whose 32 rows are all clean takes the walk without the
// default-direction test.  Returns the row's index in the matrix; `hm` = this warp needs the test.
template <bool HAS_MISSING>
__device__ __forceinline__ uint64_t tile_row(uint32_t buf, uint64_t tile, int tid, bool absent_columns, bool &hm) {
  uint32_t orig = (uint32_t)tid;
  hm = false;
  if (HAS_MISSING) {
    const uint32_t p = lds_u32(buf + 4u * (uint32_t)tid);
    orig = p & 0xFFu;
    hm = __any_sync(0xffffffffu, (p >> 8) & 1u) || absent_columns;
  }
  return tile * kTileRows + orig;
}

template <int ILP, int MINB, int TEXMODE, bool HAS_MISSING, bool PRED_LEAF, int NBUF>
__global__ void __launch_bounds__(kBlock, MINB) predict_tiles_kernel(DeviceForest f, PredictArgs a) {
  extern __shared__ __align__(128) uint32_t skey[];
  __shared__ __align__(8) unsigned long long bars[NBUF];
  const int tid = threadIdx.x;
  TilePipe<NBUF> pipe;
  pipe.begin(skey, bars, a, f.nfeat, HAS_MISSING, tid);
  while (pipe.more()) {
    const uint32_t buf = pipe.acquire(tid);
    const uint32_t my = buf + 4u * (uint32_t)(kStride + tid);
    bool hm;
    const uint64_t row = tile_row<HAS_MISSING>(buf, pipe.tile, tid, a.ncol < f.nfeat, hm);
    row_result<PRED_LEAF>(f, a, row, row < a.nrow, [&](auto &&emit) {
      if (HAS_MISSING && hm)
        forest_walk_duo<ILP, TEXMODE, true>(f, my, a.tree_begin, a.tree_end, emit);
      else
        forest_walk_duo<ILP, TEXMODE, false>(f, my, a.tree_begin, a.tree_end, emit);
    });
    pipe.release(tid);
  }
}

template <int ILP, bool HAS_MISSING, bool PRED_LEAF, bool PARK, int MINB, int TEXMODE, int CTOP>
__global__ void __launch_bounds__(kBlock, MINB) predict_tiles_nodes8_kernel(DeviceForest f, PredictArgs a) {
  extern __shared__ __align__(128) uint32_t skey[];
  __shared__ __align__(8) unsigned long long bars[1];
  const int tid = threadIdx.x;
  TilePipe<1> pipe;
  pipe.begin(skey, bars, a, f.nfeat, HAS_MISSING, tid);
  while (pipe.more()) {
    const uint32_t buf = pipe.acquire(tid);
    const uint32_t my = buf + 4u * (uint32_t)(kStride + tid);
    bool hm;
    const uint64_t row = tile_row<HAS_MISSING>(buf, pipe.tile, tid, a.ncol < f.nfeat, hm);
    row_result<PRED_LEAF>(f, a, row, row < a.nrow, [&](auto &&emit) {
      auto leaf = [&](int t, uint32_t xb, uint32_t idx) { emit(t, xb, PRED_LEAF ? (uint32_t)__ldg(f.orig_id + idx) : 0u); };
      if (HAS_MISSING && hm)
        forest_walk<ILP, true, PARK, TEXMODE, CTOP>(f, my, a.tree_begin, a.tree_end, leaf);
      else
        forest_walk<ILP, false, PARK, TEXMODE, CTOP>(f, my, a.tree_begin, a.tree_end, leaf);
    });
    pipe.release(tid);
  }
}

// ---- fused Run1 variant: the tile is assembled straight from the SoA feature fields ------------------
// (OH_GridCompMod.F90:303-345 pack + :347 create + :356 predict + :369,:1569 transform in one kernel).
// Row m of the slab is cell e = e0 + m; per feature the CTA's 256 cells are contiguous in the source
// field, so every load is coalesced and lands directly in the transposed tile — the [N x 27] matrix is
// never formed, and no transposition is needed.  Whether a tile holds missing entries is decided per
// tile (block-wide OR) and selects the walk specialisation at run time; +-inf raises the error flag.
// DUO: the constant table holds the heap-ordered tops of the two-level records; a tile with missing entries
// walks them too when they carry the default bits, else the 8-byte nodes without a table (CTOP must be 0).
template <int TEXMODE, int CTOP, bool DUO = false>
__global__ void __launch_bounds__(kBlock, DUO ? kDuoMinBlocks : 6) predict_soa_kernel(DeviceForest f, SoaArgs a) {
  extern __shared__ __align__(128) uint32_t skey[];
  const int tid = threadIdx.x;
  constexpr int B = kBlock;
  const uint64_t m = (uint64_t)blockIdx.x * B + tid;
  const bool live = m < a.nrow;
  int flags = 0;
  if (live) {
    const size_t e = a.e0 + m;
    const int c = (int)(m % (uint64_t)a.ncol);
    const bool chk_inf = !isinf(a.missing);
#pragma unroll
    for (int ft = 0; ft < 27; ++ft) {
      float x;
      if (ft == 1)  // PL_BST = (PLE(k-1) + PLE(k)) * 0.5 (:1488), / 100.0 as a true divide (:314)
        x = __fdiv_rn(__fmul_rn(__fadd_rn(__ldg(a.ple + e), __ldg(a.ple + e + a.ncol)), 0.5f), 100.0f);
      else
        x = a.src3[ft] ? __ldg(a.src3[ft] + e) : __ldg(a.src2[ft] + c);
      uint32_t k = float_key(x);
      if (x != x || x == a.missing) k = kKeyMissing, flags |= 1;
      if (chk_inf && isinf(x)) flags |= 2;
      skey[ft * kStride + tid] = k;
    }
  } else {
#pragma unroll
    for (int ft = 0; ft < 27; ++ft) skey[ft * kStride + tid] = 0u;
  }
  skey[27 * kStride + tid] = 0u;  // the slot leaves point at (nfeat == 27 is checked by the host)
  // __syncthreads_or returns a truth value, not the bitwise OR: one vote per flag
  const bool tile_missing = __syncthreads_or(flags & 1) != 0;
  if (__syncthreads_or(flags & 2) != 0) {
    if (tid == 0) atomicOr(a.flags, 2);
  }
  if (!live) return;
  const uint32_t my = (uint32_t)__cvta_generic_to_shared(skey + tid);
  float acc = f.base_score;
  auto add = [&](int, uint32_t xb, uint32_t) { acc = __fadd_rn(acc, __uint_as_float(xb)); };
  if (DUO && (!tile_missing || f.duo_has_dl)) {
    if (tile_missing)
      forest_walk_duo<kDuoIlp, kDuoTexMask, true>(f, my, 0, a.ntree_used, add);
    else
      forest_walk_duo<kDuoIlp, kDuoTexMask, false>(f, my, 0, a.ntree_used, add);
  } else if (tile_missing) {
    forest_walk<4, true, true, TEXMODE, CTOP>(f, my, 0, a.ntree_used, add);
  } else {
    forest_walk<4, false, true, TEXMODE, CTOP>(f, my, 0, a.ntree_used, add);
  }
  if (a.pred) a.pred[m] = acc;
  a.out[m] = export_transform(acc, a.exp10, a.scale);
}

cudaError_t launch_predict_soa(const DeviceForest &f, const SoaArgs &a, const Tunables &t, cudaStream_t s) {
  if (a.nrow == 0) return cudaSuccess;
  if (f.nfeat != 27) return cudaErrorInvalidValue;
  const size_t smem = (size_t)kStride * 28 * sizeof(float);
  const uint64_t nblk = (a.nrow + kBlock - 1) / kBlock;
  if (nblk > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  const bool tex = f.tex != 0 && t.variant >= 0;
  const bool duo = f.duo_ready && f.const_tree0 == 0 && f.const_ntree >= a.ntree_used;
  auto k = !tex ? predict_soa_kernel<0, 0> : (f.const_top_levels == 4 ? predict_soa_kernel<0xA, 4> : predict_soa_kernel<0xA, 0>);
  if (duo) k = predict_soa_kernel<0xA, 0, true>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<dim3((unsigned)nblk), kBlock, smem, s>>>(f, a);
  count_family(duo ? "soa_duo" : "soa_nodes8", false, false);
  return QC_LAUNCHED();
}

// ---- launches ------------------------------------------------------------------------------------------
static thread_local int g_sm_count = 0;
static int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

template <class K>
static cudaError_t launch_tiles(K k, const DeviceForest &f, const PredictArgs &a, int nbuf, int persistent_ctas_per_sm, cudaStream_t s) {
  const size_t smem = (size_t)nbuf * kStride * (size_t)(f.nfeat + 2) * sizeof(uint32_t);
  const uint64_t ntile = tile_count(a.nrow);
  uint64_t grid = ntile;
  if (persistent_ctas_per_sm > 0) grid = std::min<uint64_t>(ntile, (uint64_t)sm_count() * persistent_ctas_per_sm);
  if (grid > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<dim3((unsigned)grid), kBlock, smem, s>>>(f, a);
  return QC_LAUNCHED();
}

// two-level records, default shape; HM / PL chosen at run time
template <int ILP, int MINB, int TEXMODE, int NBUF>
static cudaError_t launch_duo(const DeviceForest &f, const PredictArgs &a, int persistent, cudaStream_t s) {
  const bool hm = a.has_missing != 0, pl = a.pred_leaf != 0;
  if (hm && pl) return launch_tiles(predict_tiles_kernel<ILP, MINB, TEXMODE, true, true, NBUF>, f, a, NBUF, persistent, s);
  if (hm) return launch_tiles(predict_tiles_kernel<ILP, MINB, TEXMODE, true, false, NBUF>, f, a, NBUF, persistent, s);
  if (pl) return launch_tiles(predict_tiles_kernel<ILP, MINB, TEXMODE, false, true, NBUF>, f, a, NBUF, persistent, s);
  return launch_tiles(predict_tiles_kernel<ILP, MINB, TEXMODE, false, false, NBUF>, f, a, NBUF, persistent, s);
}

template <int ILP, bool PARK, int MINB, int TEXMODE, int CTOP>
static cudaError_t launch_nodes8(const DeviceForest &f, const PredictArgs &a, cudaStream_t s) {
  const bool hm = a.has_missing != 0, pl = a.pred_leaf != 0;
  if (hm && pl) return launch_tiles(predict_tiles_nodes8_kernel<ILP, true, true, PARK, MINB, TEXMODE, CTOP>, f, a, 1, 0, s);
  if (hm) return launch_tiles(predict_tiles_nodes8_kernel<ILP, true, false, PARK, MINB, TEXMODE, CTOP>, f, a, 1, 0, s);
  if (pl) return launch_tiles(predict_tiles_nodes8_kernel<ILP, false, true, PARK, MINB, TEXMODE, CTOP>, f, a, 1, 0, s);
  return launch_tiles(predict_tiles_nodes8_kernel<ILP, false, false, PARK, MINB, TEXMODE, CTOP>, f, a, 1, 0, s);
}

// Persistent double-buffered variant (TMA prefetch of the next tile while this one is walked), for the HBM-bound
// end of the booster sweep.  Two 28 KB buffers per CTA: three CTAs per SM (four would need 64 bytes more than the SM's 228 KB once the 1 KB
// the system reserves per CTA is counted).
constexpr int kPersistCtasPerSm = 3;

cudaError_t launch_predict(const DeviceForest &f, const PredictArgs &a, const Tunables &t, cudaStream_t s) {
  if (a.nrow == 0 || a.tree_end <= a.tree_begin) return cudaSuccess;
  if (a.ncol > f.nfeat || f.nfeat > (int)kMaxFeatures) return cudaErrorInvalidValue;
  const bool hm = a.has_missing != 0, pl = a.pred_leaf != 0;
  const bool in_table = a.tree_begin >= f.const_tree0 && a.tree_end <= f.const_tree0 + f.const_ntree;
  // ---- two-level records (the constant tables hold these trees' tops: capi_xgb.cpp sync_const_top)
  if (f.duo_ready && in_table && (!hm || f.duo_has_dl)) {
    g_last_predict = count_family("duo", hm, pl);
    // measured (profiles/README.md): with two 28 KB buffers only three CTAs fit an SM, and the lost residency costs
    // more than the prefetch wins — the persistent loop is opt-in (qcoh_set_param persist=1), not the default
    const bool persist = t.persist > 0;
#ifdef QC_EXPERIMENTS
    // experiment grid (qcoh_set_param duo=1 + ilp / minb / duo_mask): trees in flight x resident CTAs x which
    // of the trees gather through the texture pipe; measurements in profiles/README.md
#define QC_DUO(I, M, MASK) \
  if (t.ilp == I && t.minb == M && t.duo_mask == MASK) return launch_duo<I, M, MASK, 1>(f, a, 0, s);
    QC_DUO(4, 6, 0xA) QC_DUO(4, 6, 0xE) QC_DUO(4, 6, 0xF) QC_DUO(3, 6, 0x6) QC_DUO(3, 6, 0x2) QC_DUO(6, 5, 0x2A) QC_DUO(8, 4, 0xEE)
    // shallow forests (the HBM-bound end): more resident CTAs, fewer trees in flight, LSU-only gathers
    QC_DUO(4, 6, 0x100) QC_DUO(3, 6, 0x100) QC_DUO(2, 6, 0x100) QC_DUO(2, 6, 0x2) QC_DUO(6, 5, 0x100) QC_DUO(6, 5, 0x14)
    QC_DUO(6, 5, 0x3F) QC_DUO(4, 7, 0x100) QC_DUO(4, 7, 0xA) QC_DUO(3, 7, 0x100) QC_DUO(3, 7, 0x2) QC_DUO(2, 7, 0x100) QC_DUO(5, 6, 0x100) QC_DUO(5, 6, 0xA)
#undef QC_DUO
#endif
    if (persist) return launch_duo<kDuoIlp, kPersistCtasPerSm, kDuoTexMask, 2>(f, a, kPersistCtasPerSm, s);
    if (f.duo_shallow && t.ilp == 0) return launch_duo<4, 6, 0xA, 1>(f, a, 0, s);  // kernels.hpp kShallowRecBytesPerTree
    return launch_duo<kDuoIlp, kDuoMinBlocks, kDuoTexMask, 1>(f, a, 0, s);
  }
  // ---- 8-byte depth-ordered nodes: 4 trees in flight per thread; levels 0..3 of every tree from constant
  // memory when the table holds them; below that, trees 1 and 3 of each group fetch their nodes through the
  // texture pipe and trees 0 and 2 through the LSU (mask 0xA).  Measurements: profiles/README.md.
  g_last_predict = count_family("nodes8", hm, pl);
  constexpr int kTex = 0xA;
  const bool tex = f.tex != 0 && t.variant >= 0;
  const int ctop = in_table ? f.const_top_levels : 0;  // 0 if the table does not hold these trees
#ifdef QC_EXPERIMENTS
  if (!hm && !pl) {
    if (t.park == 0) return launch_nodes8<4, false, 6, 0, 0>(f, a, s);
    if (t.variant > 0 && f.tex) {
#define QC_TEX(V, I, MASK) \
  if (t.variant == V) return ctop == 4 ? launch_nodes8<I, true, 6, MASK, 4>(f, a, s) : launch_nodes8<I, true, 6, MASK, 0>(f, a, s);
      QC_TEX(1, 4, 0xF) QC_TEX(2, 4, 0xA) QC_TEX(3, 4, 0x8) QC_TEX(4, 4, 0xE)
      QC_TEX(5, 3, 0x4) QC_TEX(6, 3, 0x6) QC_TEX(7, 6, 0x2A) QC_TEX(8, 6, 0x24) QC_TEX(9, 8, 0xAA)
      QC_TEX(10, 2, 0x2) QC_TEX(11, 5, 0x0A) QC_TEX(12, 5, 0x15)
#undef QC_TEX
    }
    if (t.ilp > 0 || t.minb > 0) {  // LSU-only builds
      const int ilp = t.ilp > 0 ? t.ilp : 3, minb = t.minb > 0 ? t.minb : 6;
#define QC_CASE(I, M) \
  if (ilp == I && minb == M) return launch_nodes8<I, true, M, 0, 0>(f, a, s);
      QC_CASE(1, 6) QC_CASE(2, 6) QC_CASE(3, 6) QC_CASE(4, 6) QC_CASE(6, 6)
      QC_CASE(2, 5) QC_CASE(3, 5) QC_CASE(4, 5) QC_CASE(6, 5)
      QC_CASE(2, 4) QC_CASE(3, 4) QC_CASE(4, 4) QC_CASE(6, 4) QC_CASE(8, 4)
      QC_CASE(4, 3) QC_CASE(6, 3) QC_CASE(8, 3)
#undef QC_CASE
    }
    if (tex && (ctop == 3 || ctop == 5 || ctop == 6)) {
      if (ctop == 3) return launch_nodes8<4, true, 6, kTex, 3>(f, a, s);
      if (ctop == 5) return launch_nodes8<4, true, 6, kTex, 5>(f, a, s);
      return launch_nodes8<4, true, 6, kTex, 6>(f, a, s);
    }
  }
#endif
  if (!tex) return launch_nodes8<4, true, 6, 0, 0>(f, a, s);
  if (ctop == 4) return launch_nodes8<4, true, 6, kTex, 4>(f, a, s);
  return launch_nodes8<4, true, 6, kTex, 0>(f, a, s);
}

// =====================================================================================
// K1 — Run1 feature assembly
// =====================================================================================
// oh_state: PL_MOD, NDWET (OH_GridCompMod.F90:1247-1257) and the level-slab count (:275-298).
// One thread per column, coalesced across columns at every level.
__global__ void __launch_bounds__(128) oh_state_kernel(Run1Dev r) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  int cnt = 0, bad = 0;
  if (c < r.ncol) {
    const float tropp = r.TROPP[c];
    const float cmp = r.dynamic_k ? tropp : r.tropp_min;
    if (!r.dynamic_k && tropp <= r.tropp_min) bad = 1;  // :287
    float ple_up = r.PLE_MOD[c];
    for (int k = 0; k < r.km; ++k) {
      const size_t e = (size_t)k * r.ncol + c;
      const float ple_dn = r.PLE_MOD[e + r.ncol];
      const float pl = __fmul_rn(__fadd_rn(ple_up, ple_dn), 0.5f);  // :1247
      const float q = r.Q_MOD[e];
      // TV = T * (1 + Q/eps) / (1 + Q), left to right (:1250)
      const float tv = __fdiv_rn(__fmul_rn(r.T_MOD[e], __fadd_rn(1.0f, __fdiv_rn(q, r.eps))), __fadd_rn(1.0f, q));
      r.PL_MOD[e] = pl;
      r.NDWET[e] = __fdiv_rn(__fmul_rn(r.avogad, pl), __fmul_rn(r.runiv, tv));  // :1257
      cnt += pl > cmp;
      ple_up = ple_dn;
    }
  }
  cnt = __reduce_max_sync(0xffffffffu, cnt);
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&r.ctl[0], cnt);
    if (bad) atomicAdd(&r.ctl[1], bad);
  }
}

cudaError_t launch_oh_state(const Run1Dev &r, cudaStream_t s) {
  oh_state_kernel<<<(r.ncol + 127) / 128, 128, 0, s>>>(r);
  return QC_LAUNCHED();
}

// oh_sums: aod (:1451-1466) and the six vertical sums (:1468-1478).  Every SUM(x(:,:,a:b),3)
// restarts at its first level and adds downward, so the DN sums are NOT a suffix scan: level k
// needs its own forward chain x(k) + x(k+1) + ... (SURVEY.md hard part 6).  One thread per
// column keeps KB chains per field in registers and sweeps the column once per block of KB
// levels; the UP sums are a running prefix.  aod and PL_BST are kept (DIAG_AOD, DIAG_PL).
// (Tried: parking the column's TAUCLW / TAUCLI / aod in shared memory for the second pass — 108 KB per 128-column CTA
// leaves 8 warps per SM and the streaming first pass then starves: the fused Run1 step grew from 43.3 to 44.2 ms at
// C360.  The sweeps of the second pass hit L2.)
constexpr int KB = 8;
__global__ void __launch_bounds__(128) oh_sums_kernel(Run1Dev r) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= r.ncol) return;
  const int nc = r.ncol, km = r.km;
  float *__restrict__ aod = r.aod;
  float *wdn = r.sums[0], *idn = r.sums[1], *iup = r.sums[2], *wup = r.sums[3], *aup = r.sums[4], *adn = r.sums[5];
  if (r.lat_deg) r.lat_deg[c] = __fmul_rn(r.LATS[c], r.r2d);         // latarr (:1444)
  if (r.so3) r.so3[c] = __fadd_rn(r.GMITO3[c], -r.GMITTO3[c]);       // stratO3 (:1446)
  // pass 1: aod and the UP prefixes
  float s_iup = 0.f, s_wup = 0.f, s_aup = 0.f;
  float z_up = r.ZLE_BST[c];
  float p_up = r.PLE_BST[c];
  for (int k = 0; k < km; ++k) {
    const size_t e = (size_t)k * nc + c;
    const float z_dn = r.ZLE_BST[e + nc];
    const float p_dn = r.PLE_BST[e + nc];
    r.pl_bst[e] = __fmul_rn(__fadd_rn(p_up, p_dn), 0.5f);  // PL_BST (:1488): bb%PL, the DIAG_PL export (:1666)
    p_up = p_dn;
    const float thick = __fadd_rn(z_up, -z_dn);  // REAL*8 gridBoxThickness holds this float exactly
    float sc = __fadd_rn(r.SCA[0][e], r.SCA[1][e]);
    sc = __fadd_rn(sc, r.SCA[2][e]);
    sc = __fadd_rn(sc, r.SCA[3][e]);
    sc = __fadd_rn(sc, r.SCA[4][e]);
    sc = __fadd_rn(sc, r.SCA[5][e]);
    sc = __fadd_rn(sc, r.SCA[6][e]);
    // double(thick) * double(sc) is exact (24 + 24 bits), so rounding it to REAL equals the
    // correctly rounded float32 product
    const float a = __fmul_rn(thick, sc);
    aod[e] = a;
    s_iup = __fadd_rn(s_iup, r.TAUCLI[e]);
    s_wup = __fadd_rn(s_wup, r.TAUCLW[e]);
    s_aup = __fadd_rn(s_aup, a);
    iup[e] = s_iup, wup[e] = s_wup, aup[e] = s_aup;
    z_up = z_dn;
  }
  // pass 2: DN chains, KB start levels at a time
  for (int k0 = 0; k0 < km; k0 += KB) {
    float cw[KB], ci[KB], ca[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) cw[j] = ci[j] = ca[j] = 0.f;
    for (int kk = k0; kk < km; ++kk) {
      const size_t e = (size_t)kk * nc + c;
      const float w = r.TAUCLW[e], i = r.TAUCLI[e], a = aod[e];
#pragma unroll
      for (int j = 0; j < KB; ++j)
        if (k0 + j <= kk) {
          cw[j] = __fadd_rn(cw[j], w);
          ci[j] = __fadd_rn(ci[j], i);
          ca[j] = __fadd_rn(ca[j], a);
        }
    }
#pragma unroll
    for (int j = 0; j < KB; ++j)
      if (k0 + j < km) {
        const size_t e = (size_t)(k0 + j) * nc + c;
        wdn[e] = cw[j], idn[e] = ci[j], adn[e] = ca[j];
      }
  }
}

cudaError_t launch_oh_sums(const Run1Dev &r, cudaStream_t s) {
  oh_sums_kernel<<<(r.ncol + 127) / 128, 128, 0, s>>>(r);
  return QC_LAUNCHED();
}

// oh_pack: xx_carr(27, N) in the reference's order (OH_GridCompMod.F90:308-345): row m runs
// over k = k1..km, then column (i fastest).  A CTA packs 256 consecutive rows: per feature the
// 256 sources are contiguous (coalesced), the [256][27] tile is transposed through shared
// memory (stride 27 is odd: conflict-free) and written back as one contiguous run.
__global__ void __launch_bounds__(256) oh_pack_kernel(Run1Dev r, int k1, uint64_t npred, float *__restrict__ X) {
  __shared__ float tile[256 * 27];
  const int tid = threadIdx.x;
  const uint64_t m0 = (uint64_t)blockIdx.x * 256;
  const uint64_t m = m0 + tid;
  int flags = 0;
  if (m < npred) {
    const size_t e = (size_t)(k1 - 1) * r.ncol + m;
    const int c = (int)(m % (uint64_t)r.ncol);
    float *row = tile + tid * 27;
    row[0] = __fmul_rn(r.LATS[c], r.r2d);  // :1444
    // PL_BST = (PLE(k-1) + PLE(k)) * 0.5 (:1488), then / 100.0 as a true divide (:314)
    row[1] = __fdiv_rn(__fmul_rn(__fadd_rn(r.PLE_BST[e], r.PLE_BST[e + r.ncol]), 0.5f), 100.0f);
    row[2] = r.T_BST[e];
    row[3] = r.NO2[e];
    row[4] = r.O3[e];
    row[5] = r.CH4[e];
    row[6] = r.CO[e];
    row[7] = r.ISOP[e];
    row[8] = r.ACET[e];
    row[9] = r.C2H6[e];
    row[10] = r.C3H8[e];
    row[11] = r.PRPE[e];
    row[12] = r.ALK4[e];
    row[13] = r.MP[e];
    row[14] = r.H2O2[e];
    row[15] = r.sums[0][e];  // TAUCLWDN
    row[16] = r.sums[1][e];  // TAUCLIDN
    row[17] = r.sums[2][e];  // TAUCLIUP
    row[18] = r.sums[3][e];  // TAUCLWUP
    row[19] = r.FCLD[e];
    row[20] = r.Q_BST[e];
    row[21] = __fadd_rn(r.GMITO3[c], -r.GMITTO3[c]);  // :1446
    row[22] = r.ALBUV[c];
    row[23] = r.sums[4][e];  // AODUP
    row[24] = r.sums[5][e];  // AODDN
    row[25] = r.CH2O[e];
    row[26] = r.SZA[c];
    const bool fin = !isinf(r.missing);
#pragma unroll
    for (int f = 0; f < 27; ++f) {
      const float v = row[f];
      if (v != v || v == r.missing) flags |= 1;
      if (fin && isinf(v)) flags |= 2;
    }
  }
  flags = __reduce_or_sync(0xffffffffu, flags);
  if ((tid & 31) == 0 && flags) atomicOr(&r.ctl[2], flags);
  __syncthreads();
  const uint64_t left = npred - m0;
  const int n = (int)(left < 256 ? left : 256) * 27;
  float *dst = X + m0 * 27;
  for (int i = tid; i < n; i += 256) dst[i] = tile[i];
}

cudaError_t launch_oh_pack(const Run1Dev &r, int k1, float *X, cudaStream_t s) {
  const uint64_t npred = (uint64_t)r.ncol * (uint64_t)(r.km - k1 + 1);
  if (npred == 0) return cudaSuccess;
  oh_pack_kernel<<<(unsigned)((npred + 255) / 256), 256, 0, s>>>(r, k1, npred, X);
  return QC_LAUNCHED();
}

// oh_finalize (OH_GridCompMod.F90:1579-1599): OH_boost export, troposphere mask against the
// climatological OH using the CURRENT model PL / TROPP, mol/mol -> molec/cm3; optionally the
// loss frequencies k(T)[OH] the downstream CH4 / CO chemistry needs.
__global__ void __launch_bounds__(256) oh_finalize_kernel(Run1Dev r, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const int c = (int)(e % (uint64_t)r.ncol);
    const float ml = r.OH_ML[e];
    if (r.OH_boost) r.OH_boost[e] = ml;
    const float oh = (r.PL_MOD[e] > r.TROPP[c]) ? ml : r.OH_CLIM[e];
    const float ohn = __fmul_rn(__fmul_rn(oh, r.NDWET[e]), 1.0e-6f);
    r.OH[e] = ohn;
    // first-order loss frequencies for the CH4 / CO consumers of OH (build-defined; qcoh.h): float64, rounded once
    if (r.LOSS_CH4) r.LOSS_CH4[e] = (float)(2.45e-12 * exp(-1775.0 / (double)r.T_MOD[e]) * (double)ohn);
    if (r.LOSS_CO) r.LOSS_CO[e] = (float)(1.5e-13 * (1.0 + 0.6 * ((double)r.PL_MOD[e] / 101325.0)) * (double)ohn);
  }
}

cudaError_t launch_oh_finalize(const Run1Dev &r, cudaStream_t s) {
  const uint64_t n = (uint64_t)r.ncol * r.km;
  int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  oh_finalize_kernel<<<blocks, 256, 0, s>>>(r, n);
  return QC_LAUNCHED();
}

// K4 — build-defined diagnostic (not in the reference): float64 partial sums
//   [0] sum(OH * w)   [1] sum(w)            w = dP * area / g over tropospheric cells (PL > TROPP)
//   [2] sum(nCH4 * V) [3] sum(k(T) * OH * nCH4 * V),  k = 2.45e-12 exp(-1775/T), V = area * dz
// OH in molec/cm3, nCH4 = CH4 * NDWET.  The host (or NCCL) adds the ranks' partial sums.
__global__ void __launch_bounds__(256) oh_diag_kernel(Run1Dev r, uint64_t n) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const int c = (int)(e % (uint64_t)r.ncol);
    if (r.PL_MOD[e] > r.TROPP[c]) {
      const double area = r.AREA[c];
      const double dp = (double)r.PLE_MOD[e + r.ncol] - (double)r.PLE_MOD[e];
      const double dz = (double)r.ZLE_BST[e] - (double)r.ZLE_BST[e + r.ncol];
      const double w = dp * area / 9.80665;
      const double oh = r.OH[e];
      const double nch4 = (double)r.CH4[e] * (double)r.NDWET[e];
      const double kt = 2.45e-12 * exp(-1775.0 / (double)r.T_MOD[e]);
      s0 += oh * w, s1 += w, s2 += nch4 * area * dz, s3 += kt * oh * nch4 * area * dz;
    }
  }
  __shared__ double sh[4][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_down_sync(0xffffffffu, s0, o);
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
    s3 += __shfl_down_sync(0xffffffffu, s3, o);
  }
  if (lane == 0) sh[0][wid] = s0, sh[1][wid] = s1, sh[2][wid] = s2, sh[3][wid] = s3;
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
    atomicAdd(&r.diag[threadIdx.x], t);
  }
}

cudaError_t launch_oh_diag(const Run1Dev &r, cudaStream_t s) {
  const uint64_t n = (uint64_t)r.ncol * r.km;
  int blocks = (int)((n + 255) / 256 < 148 * 4 ? (n + 255) / 256 : 148 * 4);
  oh_diag_kernel<<<blocks, 256, 0, s>>>(r, n);
  return QC_LAUNCHED();
}

}  // namespace qcoh
