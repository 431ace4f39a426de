// shard.cu — device side of the row-sharded embedding exchange (BASELINE config 5).
//
// Row r of every table lives on rank r % G as local row r / G.  Per step each rank
//   1. buckets its B x F lookups by owner (b2ctr_shard_bucketize + b2ctr_shard_fill): keys grouped by
//      destination rank, and pos[b,f] = where the answer for (b,f) will sit in the returned row buffer;
//   2. exchanges the keys (NCCL all-to-all, host side: deepctr_b200/parallel.py);
//   3. owners serve the received keys from their shard (b2ctr_shard_gather_rows) and the rows travel
//      back (second all-to-all); the requester assembles X / FM / linear with the ordinary fused gather,
//      using the returned buffer as its "table" and pos as ids;
//   4. backward: gradient rows go back along the same route and the owner applies them with
//      b2ctr_shard_scatter_rows (fused SGD, red.global.add.v4.f32).
// The reference has no counterpart (no sharding, no collectives): new capability.
#include "common.cuh"

namespace b2ctr {

constexpr int kMaxWorld = 64;
constexpr int kShardFeat = 64;

struct ShardIdx {
  const void* idx[kShardFeat];
  int64_t stride[kShardFeat];
  int64_t vocab[kShardFeat];     // FULL (unsharded) vocabulary of feature f
  unsigned long long* oob;
  int32_t nfeat;
  int32_t dtype;
};
struct ShardTables {
  float* table[kShardFeat];
  float* lin[kShardFeat];
};

// key = feature << 40 | local_row
__device__ __forceinline__ int64_t make_key(int f, int64_t local_row) { return ((int64_t)f << 40) | local_row; }

// pass 1: slot[item] = owner << 32 | rank inside the owner's bucket; counts[owner] accumulates totals.
// Block-local histogram in shared memory, one global atomic per (block, owner).
__global__ void __launch_bounds__(256)
    shard_bucketize_kernel(const __grid_constant__ ShardIdx si, int64_t batch, int world, int32_t* counts,
                           int64_t* slot) {
  __shared__ int32_t hist[kMaxWorld];
  __shared__ int32_t base[kMaxWorld];
  const int64_t total = batch * si.nfeat;
  const int64_t per_block = 256 * 8;
  for (int64_t blk0 = (int64_t)blockIdx.x * per_block; blk0 < total; blk0 += (int64_t)gridDim.x * per_block) {
    if (threadIdx.x < world) hist[threadIdx.x] = 0;
    __syncthreads();
    int owner[8], rank[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t item = blk0 + u * 256 + threadIdx.x;
      owner[u] = -1;
      if (item < total) {
        const int64_t b = item / si.nfeat;
        const int f = (int)(item - b * si.nfeat);
        int64_t id = load_idx(si.idx[f], b * si.stride[f], si.dtype);
        if (!id_in_range(id, si.vocab[f])) { note_oob(si.oob); id = 0; }   // routed as row 0, raised by the host
        owner[u] = (int)(id % world);
        rank[u] = atomicAdd(&hist[owner[u]], 1);
      }
    }
    __syncthreads();
    if (threadIdx.x < world) base[threadIdx.x] = atomicAdd(&counts[threadIdx.x], hist[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t item = blk0 + u * 256 + threadIdx.x;
      if (owner[u] >= 0) slot[item] = ((int64_t)owner[u] << 32) | (uint32_t)(base[owner[u]] + rank[u]);
    }
    __syncthreads();
  }
}
// pass 2: offsets = exclusive scan of counts; keys[offset[owner] + rank] = key; pos[item] = that index
__global__ void __launch_bounds__(256)
    shard_fill_kernel(const __grid_constant__ ShardIdx si, int64_t batch, int world,
                      const int32_t* __restrict__ counts, const int64_t* __restrict__ slot, int64_t* keys,
                      int32_t* pos) {
  __shared__ int32_t off[kMaxWorld];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int d = 0; d < world; ++d) { off[d] = acc; acc += counts[d]; }
  }
  __syncthreads();
  const int64_t total = batch * si.nfeat;
  for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < total;
       item += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = item / si.nfeat;
    const int f = (int)(item - b * si.nfeat);
    int64_t id = load_idx(si.idx[f], b * si.stride[f], si.dtype);
    if (!id_in_range(id, si.vocab[f])) id = 0;
    const int64_t s = slot[item];
    const int owner = (int)(s >> 32);
    const int p = off[owner] + (int)(s & 0xffffffff);
    keys[p] = make_key(f, id / world);
    pos[item] = p;
  }
}

// owner side: rows[i, :] = table_f[row], lin_out[i] = lin_f[row] for key i
template <int LPR>
__global__ void __launch_bounds__(256)
    shard_gather_rows_kernel(const __grid_constant__ ShardTables st, const int64_t* __restrict__ keys, int64_t n,
                             int dim, float* rows, float* lin_out) {
  constexpr int KPW = 32 / LPR;  // keys per warp pass
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, chunk = lane % LPR;
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * KPW * 4; i0 < n;
       i0 += nw * KPW * 4) {
    float4 v[4];
    int64_t ii[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ii[u] = i0 + u * KPW + sub;
      if (ii[u] < n) {
        const int64_t k = keys[ii[u]];
        const int f = (int)(k >> 40);
        const int64_t row = k & ((1ll << 40) - 1);
        v[u] = ldg_stream_f4(st.table[f] + row * dim + chunk * 4);
        if (lin_out && chunk == 0) lin_out[ii[u]] = st.lin[f] ? st.lin[f][row] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ii[u] < n) stg_stream_f4(rows + ii[u] * dim + chunk * 4, v[u]);
  }
}
// owner side backward: table_f[row] += scale * grows[i]; lin_f[row] += lin_scale * glin[i]
template <int LPR>
__global__ void __launch_bounds__(256)
    shard_scatter_rows_kernel(const __grid_constant__ ShardTables st, const int64_t* __restrict__ keys, int64_t n,
                              int dim, const float* __restrict__ grows, const float* __restrict__ glin,
                              float scale, float lin_scale) {
  constexpr int KPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, chunk = lane % LPR;
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * KPW * 4; i0 < n;
       i0 += nw * KPW * 4) {
    float4 v[4];
    int64_t ii[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ii[u] = i0 + u * KPW + sub;
      if (ii[u] < n) v[u] = ldg_stream_f4(grows + ii[u] * dim + chunk * 4);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (ii[u] < n) {
        const int64_t k = keys[ii[u]];
        const int f = (int)(k >> 40);
        const int64_t row = k & ((1ll << 40) - 1);
        v[u].x *= scale; v[u].y *= scale; v[u].z *= scale; v[u].w *= scale;
        red_add_f4(st.table[f] + row * dim + chunk * 4, v[u]);
        if (glin && chunk == 0 && st.lin[f]) red_add_f1(st.lin[f] + row, lin_scale * glin[ii[u]]);
      }
    }
  }
}

static b2ctr_status_t fill_idx(const b2ctr_feature_t* feats, int32_t nfeat, ShardIdx* si) {
  B2_REQUIRE(feats && nfeat > 0 && nfeat <= kShardFeat, "shard: nfeat must be in [1,%d]", kShardFeat);
  si->nfeat = nfeat;
  si->oob = oob_counter();
  si->dtype = feats[0].idx_dtype;
  for (int f = 0; f < nfeat; ++f) {
    B2_REQUIRE(feats[f].idx && feats[f].idx_dtype == si->dtype, "shard: feature %d bad idx / mixed dtypes", f);
    si->idx[f] = feats[f].idx;
    si->stride[f] = feats[f].idx_stride;
    si->vocab[f] = feats[f].vocab;
  }
  return B2CTR_OK;
}

}  // namespace b2ctr

using namespace b2ctr;
#define ST ((cudaStream_t)stream)

extern "C" {

b2ctr_status_t b2ctr_shard_bucketize(const b2ctr_feature_t* feats, int32_t nfeat, int64_t batch, int32_t world,
                                     int32_t* counts, int64_t* slot, void* stream) {
  ShardIdx si;
  b2ctr_status_t s = fill_idx(feats, nfeat, &si);
  if (s != B2CTR_OK) return s;
  B2_REQUIRE(world >= 1 && world <= kMaxWorld && counts && slot, "shard_bucketize: bad arguments");
  if (batch <= 0) return B2CTR_OK;
  shard_bucketize_kernel<<<grid_for(batch * nfeat, 2048, 4), 256, 0, ST>>>(si, batch, world, counts, slot);
  B2_CHECK_LAUNCH("b2ctr_shard_bucketize");
  return B2CTR_OK;
}

b2ctr_status_t b2ctr_shard_fill(const b2ctr_feature_t* feats, int32_t nfeat, int64_t batch, int32_t world,
                                const int32_t* counts, const int64_t* slot, int64_t* keys, int32_t* pos,
                                void* stream) {
  ShardIdx si;
  b2ctr_status_t s = fill_idx(feats, nfeat, &si);
  if (s != B2CTR_OK) return s;
  B2_REQUIRE(world >= 1 && world <= kMaxWorld && counts && slot && keys && pos, "shard_fill: bad arguments");
  if (batch <= 0) return B2CTR_OK;
  shard_fill_kernel<<<grid_for(batch * nfeat, 256, 8), 256, 0, ST>>>(si, batch, world, counts, slot, keys, pos);
  B2_CHECK_LAUNCH("b2ctr_shard_fill");
  return B2CTR_OK;
}

static b2ctr_status_t fill_tables(float* const* tables, float* const* lin_tables, int32_t nfeat, int32_t dim,
                                  ShardTables* st) {
  B2_REQUIRE(tables && nfeat > 0 && nfeat <= kShardFeat, "shard: bad tables");
  B2_REQUIRE(dim == 4 || dim == 8 || dim == 16 || dim == 32 || dim == 64 || dim == 128,
             "shard: dim must be one of 4,8,16,32,64,128");
  for (int f = 0; f < nfeat; ++f) {
    B2_REQUIRE(tables[f] && ((uintptr_t)tables[f] & 15) == 0, "shard: table %d NULL or misaligned", f);
    st->table[f] = tables[f];
    st->lin[f] = lin_tables ? lin_tables[f] : nullptr;
  }
  return B2CTR_OK;
}

#define B2_DISPATCH_LPR2(KERNEL, dim, ...)                                                    \
  switch ((dim) / 4) {                                                                        \
    case 1: KERNEL<1><<<grid, 256, 0, ST>>>(__VA_ARGS__); break;                              \
    case 2: KERNEL<2><<<grid, 256, 0, ST>>>(__VA_ARGS__); break;                              \
    case 4: KERNEL<4><<<grid, 256, 0, ST>>>(__VA_ARGS__); break;                              \
    case 8: KERNEL<8><<<grid, 256, 0, ST>>>(__VA_ARGS__); break;                              \
    case 16: KERNEL<16><<<grid, 256, 0, ST>>>(__VA_ARGS__); break;                            \
    default: KERNEL<32><<<grid, 256, 0, ST>>>(__VA_ARGS__); break;                            \
  }

b2ctr_status_t b2ctr_shard_gather_rows(float* const* tables, float* const* lin_tables, int32_t nfeat,
                                       int32_t dim, const int64_t* keys, int64_t n, float* rows,
                                       float* lin_out, void* stream) {
  ShardTables st;
  b2ctr_status_t s = fill_tables(tables, lin_tables, nfeat, dim, &st);
  if (s != B2CTR_OK) return s;
  B2_REQUIRE(keys && rows, "shard_gather_rows: NULL pointer");
  if (n <= 0) return B2CTR_OK;
  const int grid = grid_for(n, 8 * 4 * (128 / dim), 4);
  B2_DISPATCH_LPR2(shard_gather_rows_kernel, dim, st, keys, n, dim, rows, lin_out);
  B2_CHECK_LAUNCH("b2ctr_shard_gather_rows");
  return B2CTR_OK;
}

b2ctr_status_t b2ctr_shard_scatter_rows(float* const* tables, float* const* lin_tables, int32_t nfeat,
                                        int32_t dim, const int64_t* keys, int64_t n, const float* grows,
                                        const float* glin, float scale, float lin_scale, void* stream) {
  ShardTables st;
  b2ctr_status_t s = fill_tables(tables, lin_tables, nfeat, dim, &st);
  if (s != B2CTR_OK) return s;
  B2_REQUIRE(keys && grows, "shard_scatter_rows: NULL pointer");
  if (n <= 0) return B2CTR_OK;
  const int grid = grid_for(n, 8 * 4 * (128 / dim), 4);
  B2_DISPATCH_LPR2(shard_scatter_rows_kernel, dim, st, keys, n, dim, grows, glin, scale, lin_scale);
  B2_CHECK_LAUNCH("b2ctr_shard_scatter_rows");
  return B2CTR_OK;
}

}  // extern "C"
