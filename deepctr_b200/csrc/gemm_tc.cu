// gemm_tc.cu — fp32-in / fp32-out GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Precision mode B2CTR_GEMM_BF16X3: every fp32 operand x is split on the fly into two bf16 values
//   x = hi + lo,  hi = bf16(x),  lo = bf16(x - hi)           (|x - hi - lo| <= 2^-17 |x|)
// and the product is accumulated in fp32 TMEM as  hi*hi + hi*lo + lo*hi  (the dropped lo*lo term is
// <= 2^-16 relative), i.e. three kind::f16 UMMAs per K-step.  That keeps the reference's fp32 logits
// within the 1e-4 bar (north_star) at ~1/3 of the bf16 tensor throughput instead of the FFMA pipe.
//
// Kernel anatomy (one 128 x BN output tile per CTA, K split over gridDim.z):
//   warps 0-7  : producers. Read the fp32 operand tiles straight from global memory with coalesced
//                loads in whatever layout they are stored (row- or column-major: each thread gathers
//                8 consecutive K values of one row), split them, and write the bf16 hi / lo planes into
//                shared memory in the canonical K-major SWIZZLE_128B UMMA layout (16 B chunk index
//                XOR row%8).  fence.proxy.async + mbarrier hand the stage to the MMA warp.
//                After the main loop the same warps run the epilogue: tcgen05.ld the accumulator,
//                alpha / bias / activation / accumulate, store fp32 C.
//   warp 8     : one elected lane issues tcgen05.mma (SS form, cta_group::1, M=128, N=BN, K=16) x 4 K-steps
//                x 3 split terms per stage, then tcgen05.commit to release the stage / publish the tile.
// Stages: kStages x (A hi+lo 32 KB + B hi+lo BN*256 B).  TMEM: BN fp32 columns x 128 lanes.
#include <cuda.h>          // CUtensorMap + enums only: the encoder is resolved through the runtime (no libcuda link)
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"

namespace b2ctr {

constexpr int kTM = 128;       // UMMA M (rows of the output tile, TMEM lanes)
constexpr int kTK = 64;        // K elements per stage = one 128-byte swizzle atom of bf16
constexpr int kProducerWarps = 8;
constexpr int kTcThreads = (kProducerWarps + 1) * 32;

struct TcArgs {
  const float* a; const float* b; float* c; const float* bias; float* ws;
  int64_t m, n, k;
  int64_t sam, sak, sbn, sbk;   // A(m,k) = a[m*sam + k*sak];  B(n,k) = b[n*sbn + k*sbk]
  int64_t ldc;
  int64_t k_per_split;
  float alpha;
  int act, accumulate, splits;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   start address >> 4 | LBO (ignored for swizzled K-major, set to 1) << 16 | SBO = 1024 B >> 4 << 32 |
//   version 1 << 46 | layout SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major, SWIZZLE_128B: the tile is a row of 64-element (128 B) atoms along M/N, each atom holding 64
// k-rows of 128 B: LBO = atom stride (8192 B), SBO = stride between groups of 8 k-rows (1024 B).
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (512ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t umma_desc_any(uint32_t base, int ks, int mn) {
  // one UMMA K-step = 16 bf16 along K: 32 B inside the swizzled row (K-major) or 16 k-rows = 2048 B (MN-major)
  return mn ? umma_desc_mn(base + ks * 2048) : umma_desc(base + ks * 32);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 lo, __nv_bfloat16 hi) {
  return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

// Gather 8 consecutive K values of row r (global k in [k0, k0+8)), split, and store both planes.
// tile layout: row r at byte r*128, 16-byte chunk c stored at slot (c ^ (r & 7)).
template <bool K_CONTIG>
__device__ __forceinline__ void produce_chunk(const float* __restrict__ base, int64_t sr, int64_t sk,
                                              int64_t row, int64_t nrows, int64_t k0, int64_t kend,
                                              bool vec_ok, unsigned char* hi_tile, unsigned char* lo_tile,
                                              int r, int c) {
  float v[8];
  if (row < nrows) {
    const float* p = base + row * sr + k0 * sk;
    if (K_CONTIG && vec_ok && k0 + 8 <= kend) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (k0 + j < kend) ? __ldg(p + j * sk) : 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * j] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(h1));
    h[j] = pack_bf16(h0, h1);
    l[j] = pack_bf16(l0, l1);
  }
  const int off = r * 128 + ((c ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_tile + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int BN, int STAGES, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(kTcThreads, STAGES == 1 ? 3 : 1) gemm_bf16x3_kernel(const TcArgs g) {
  constexpr int A_PLANE = kTM * 128;       // bytes of one bf16 plane of the A tile (128 rows x 64 k)
  constexpr int B_PLANE = BN * 128;
  constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                                          ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.y * kTM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_per_split;
  const int64_t kend = kbeg + g.k_per_split < g.k ? kbeg + g.k_per_split : g.k;
  const int nkb = kend > kbeg ? (int)((kend - kbeg + kTK - 1) / kTK) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], kProducerWarps);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kProducerWarps) {  // the MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_slot)),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  if (warp < kProducerWarps) {
    // ------------------------------- producers -------------------------------------------------
    const bool a_vec = A_KC && (g.sam % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.a) & 15) == 0);
    const bool b_vec = B_KC && (g.sbn % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.b) & 15) == 0);
    const int tid = threadIdx.x;  // 0..255
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
      unsigned char* st = tiles + (size_t)s * STAGE;
      const int64_t k0 = kbeg + (int64_t)kb * kTK;
      // A tile: 128 rows x 8 chunks.  K-contiguous storage: chunk fastest (coalesced along k);
      // otherwise row fastest (coalesced along m).
#pragma unroll 4
      for (int t = tid; t < kTM * 8; t += kProducerWarps * 32) {
        const int r = A_KC ? t >> 3 : t % kTM;
        const int c = A_KC ? t & 7 : t / kTM;
        const int64_t kk = k0 + c * 8;
        produce_chunk<A_KC>(g.a, g.sam, g.sak, m0 + r, g.m, kk, kend, a_vec && ((kk & 3) == 0), st,
                            st + A_PLANE, r, c);
      }
#pragma unroll 4
      for (int t = tid; t < BN * 8; t += kProducerWarps * 32) {
        const int r = B_KC ? t >> 3 : t % BN;
        const int c = B_KC ? t & 7 : t / BN;
        const int64_t kk = k0 + c * 8;
        produce_chunk<B_KC>(g.b, g.sbn, g.sbk, n0 + r, g.n, kk, kend, b_vec && ((kk & 3) == 0),
                            st + 2 * A_PLANE, st + 2 * A_PLANE + B_PLANE, r, c);
      }
      // make the generic-proxy writes visible to the tensor core (async proxy), then signal
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
    // ------------------------------- epilogue --------------------------------------------------
    if (nkb > 0) {
      mbar_wait(&accum_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int sub = warp & 3;                 // TMEM sub-partition this warp may read
    const int half = warp >> 2;               // which half of the BN columns
    constexpr int HALVES = BN >= 64 ? 2 : 1;
    constexpr int COLS = BN / HALVES;         // columns per warp
    const int64_t gm = m0 + sub * 32 + lane;
    const bool vec_c = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.c) & 15) == 0) &&
                       (!g.bias || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0) && g.splits == 1;
    if (half < HALVES) {
#pragma unroll 1
      for (int c0 = 0; c0 < COLS; c0 += 32) {
        uint32_t r[32];
        if (nkb > 0) {
          const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(half * COLS + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
              "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
                "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (gm < g.m) {
          const int64_t gn0 = n0 + half * COLS + c0;
          if (vec_c && gn0 + 32 <= g.n) {
            float* crow = g.c + gm * g.ldc + gn0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v = make_float4(g.alpha * __uint_as_float(r[j]), g.alpha * __uint_as_float(r[j + 1]),
                                     g.alpha * __uint_as_float(r[j + 2]), g.alpha * __uint_as_float(r[j + 3]));
              if (g.accumulate) {
                const float4 o = *reinterpret_cast<const float4*>(crow + j);
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
              }
              if (g.bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + gn0 + j));
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
              }
              v.x = act_apply(v.x, g.act); v.y = act_apply(v.y, g.act);
              v.z = act_apply(v.z, g.act); v.w = act_apply(v.w, g.act);
              *reinterpret_cast<float4*>(crow + j) = v;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int64_t gn = gn0 + j;
              if (gn < g.n) {
                float v = g.alpha * __uint_as_float(r[j]);
                if (g.splits > 1) {
                  g.ws[((int64_t)blockIdx.z * g.m + gm) * g.n + gn] = v;
                } else {
                  if (g.accumulate) v += g.c[gm * g.ldc + gn];
                  if (g.bias) v += g.bias[gn];
                  g.c[gm * g.ldc + gn] = act_apply(v, g.act);
                }
              }
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ------------------------------- MMA issuer ------------------------------------------------
    const uint32_t idesc = umma_idesc(BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&full_bar[s], (kb / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE);
        const uint32_t a_hi = sa, a_lo = sa + A_PLANE, b_hi = sa + 2 * A_PLANE,
                       b_lo = sa + 2 * A_PLANE + B_PLANE;
#pragma unroll
        for (int ks = 0; ks < kTK / 16; ++ks) {
          const uint32_t ko = ks * 32;  // 16 bf16 = 32 bytes along K inside the swizzle atom
          umma_f16(tmem_base, umma_desc(a_hi + ko), umma_desc(b_hi + ko), idesc, (kb | ks) ? 1u : 0u);
          umma_f16(tmem_base, umma_desc(a_hi + ko), umma_desc(b_lo + ko), idesc, 1u);
          umma_f16(tmem_base, umma_desc(a_lo + ko), umma_desc(b_hi + ko), idesc, 1u);
        }
        umma_commit(&empty_bar[s]);               // frees the stage once these MMAs retire
        if (kb == nkb - 1) umma_commit(&accum_bar);  // accumulator complete
      }
      __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == kProducerWarps) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
  }
}

__global__ void tc_splitk_reduce_kernel(const TcArgs g) {
  const int64_t total = g.m * g.n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t gm = i / g.n, gn = i - gm * g.n;
    float v = 0.f;
    for (int z = 0; z < g.splits; ++z) v += g.ws[(int64_t)z * total + i];
    if (g.accumulate) v += g.c[gm * g.ldc + gn];
    if (g.bias) v += g.bias[gn];
    g.c[gm * g.ldc + gn] = act_apply(v, g.act);
  }
}

template <int BN, int STAGES>
static cudaError_t launch_tc(const TcArgs& ta, bool akc, bool bkc, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (2 * kTM * 128 + 2 * BN * 128) + 1024;
  dim3 grid((unsigned)ceil_div(ta.n, BN), (unsigned)ceil_div(ta.m, kTM), (unsigned)ta.splits);
#define B2_TC_LAUNCH(AK, BK)                                                                         \
  do {                                                                                               \
    auto kern = gemm_bf16x3_kernel<BN, STAGES, AK, BK>;                                              \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                  \
    kern<<<grid, kTcThreads, smem, st>>>(ta);                                                        \
  } while (0)
  if (akc && bkc) B2_TC_LAUNCH(true, true);
  else if (akc && !bkc) B2_TC_LAUNCH(true, false);
  else if (!akc && bkc) B2_TC_LAUNCH(false, true);
  else B2_TC_LAUNCH(false, false);
#undef B2_TC_LAUNCH
  return cudaGetLastError();
}

// ================================================================================================
// Variant 2 ("planes"): the fp32 -> (hi, lo) bf16 split is done ONCE per operand by a streaming kernel
// into K-major planes [rows_pad, k_pad] in workspace memory; the GEMM kernel then only moves bytes:
// 8 warps cp.async 16-byte chunks of the planes into the swizzled UMMA tiles (3-stage pipeline, no
// register staging, no conversion instructions), 1 warp issues the same 3 UMMAs per K-step.
// ncu on variant 1 (profiles/r1_gemm_tc_3stage.txt: tensor pipe ~7 %, warps active 14 %) showed the
// producers, not the tensor core, bound the kernel: every CTA re-converted both operand tiles.
// ================================================================================================
struct PlaneArgs {
  // K-major planes: [rows_pad, k_pad] (k contiguous).  MN-major planes: [k_pad, rows_pad] (row index
  // contiguous) - the natural layout of a row-major operand whose reduction dim is its row index
  // (wgrad: X^T, dZ^T; forward: W).  *_pitch = elements between consecutive plane rows.
  const __nv_bfloat16* a_hi; const __nv_bfloat16* a_lo;
  const __nv_bfloat16* b_hi; const __nv_bfloat16* b_lo;
  int64_t a_pitch, b_pitch;
  int a_mn, b_mn;
  float* c; const float* bias; float* ws;
  int64_t m, n, k_pad;
  int64_t ldc;
  int64_t k_per_split;
  float alpha;
  int act, accumulate, splits;
  // CIN mode (cin_on): the A operand is never stored anywhere - the producer warps GENERATE its bf16 hi/lo tile
  // in shared memory from the two factors of the outer product (deepctr/layers/interaction.py:287-297):
  //   A[r, i*hp + j] = t0[r*ld0 + i] * xk[r*ldk + j]   (j < h, i < m; zero otherwise),  r = (sample, embedding dim)
  // a_mn = 0: A is [rows, m*hp] (forward, M = r);  a_mn = 1: A^T, i.e. M = i*hp + j and K = r (filter gradient).
  const float* cin_t0; const float* cin_xk;
  int64_t cin_ld0, cin_ldk, cin_rows;
  int cin_m, cin_h, cin_hp, cin_on;
  // FOLD epilogue (CIN backward): dT0 [rows, cin_ld0] and dXk [rows, fold_ldx], both accumulated with red.add
  float* fold_dt0; float* fold_dxk; int64_t fold_ldx;
  int gen_groups;     // generating producers: 2 = two groups of 128 threads alternate stages, 1 = all 256 share every stage
  int debug;          // B2CTR_TC_DEBUG knock-outs (WRONG RESULTS; tools/gemm_sweep.py attributes the per-tile cost with them):
                      // 1 = no global stores in the epilogue, 2 = no tcgen05.ld, 4 = no MMAs issued, 8 = no operand loads
};

// dst planes [rows_pad, k_pad] <- src(r, k) = p[r*sr + k*sk]; zero outside [rows, k).
// K-contiguous source: thread = (row, 8 consecutive k).
__global__ void __launch_bounds__(256)
    split_planes_kernel(const float* __restrict__ p, int64_t sr, int64_t rows, int64_t k, int64_t rows_pad,
                        int64_t k_pad, __nv_bfloat16* hi, __nv_bfloat16* lo, int vec_ok) {
  const int64_t chunks = k_pad / 8;
  const int64_t total = rows_pad * chunks;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / chunks;
    const int64_t k0 = (t - r * chunks) * 8;
    float v[8];
    if (r < rows && vec_ok && k0 + 8 <= k) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p + r * sr + k0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p + r * sr + k0) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (r < rows && k0 + j < k) ? __ldg(p + r * sr + k0 + j) : 0.f;
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
      h[j] = pack_bf16(h0, h1);
      l[j] = pack_bf16(__float2bfloat16_rn(v[2 * j] - __bfloat162float(h0)),
                       __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(h1)));
    }
    *reinterpret_cast<uint4*>(hi + r * k_pad + k0) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + r * k_pad + k0) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}
// Row-contiguous source (src(r,k) = p[k*sk + r]): 64 x 64 tile transposed through shared memory.
__global__ void __launch_bounds__(256)
    split_planes_t_kernel(const float* __restrict__ p, int64_t sk, int64_t rows, int64_t k, int64_t rows_pad,
                          int64_t k_pad, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  __shared__ float tile[64][65];
  const int64_t r0 = (int64_t)blockIdx.x * 64, k0 = (int64_t)blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
#pragma unroll 4
  for (int kk = ty; kk < 64; kk += 4) {
    const int64_t gr = r0 + tx, gk = k0 + kk;
    tile[kk][tx] = (gr < rows && gk < k) ? __ldg(p + gk * sk + gr) : 0.f;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 64 * 8; t += 256) {
    const int r = t >> 3, c = t & 7;
    const int64_t gr = r0 + r;
    if (gr >= rows_pad) continue;
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v0 = tile[c * 8 + 2 * j][r], v1 = tile[c * 8 + 2 * j + 1][r];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
      h[j] = pack_bf16(h0, h1);
      l[j] = pack_bf16(__float2bfloat16_rn(v0 - __bfloat162float(h0)), __float2bfloat16_rn(v1 - __bfloat162float(h1)));
    }
    *reinterpret_cast<uint4*>(hi + gr * k_pad + k0 + c * 8) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + gr * k_pad + k0 + c * 8) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// CTAs per SM follow from the stage footprint: a single-stage 128x128 tile (65 KB) fits three times, so
// short-K GEMMs (dgrad of the first layer: K = 256 = 4 k-blocks) overlap one CTA's epilogue with its
// neighbours' main loops instead of idling the SM.
template <int BN, int STAGES>
__global__ void __launch_bounds__(kTcThreads, (STAGES * (2 * kTM * 128 + 2 * BN * 128) + 1024 <= 75 * 1024) ? 3 : 1)
    gemm_planes_kernel(const PlaneArgs g) {
  constexpr int A_PLANE = kTM * 128;
  constexpr int B_PLANE = BN * 128;
  constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                                          ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.y * kTM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_per_split;
  const int64_t kend = kbeg + g.k_per_split < g.k_pad ? kbeg + g.k_per_split : g.k_pad;
  const int nkb = kend > kbeg ? (int)((kend - kbeg) / kTK) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], kProducerWarps);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kProducerWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_slot)),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  if (warp < kProducerWarps) {
    const int tid = threadIdx.x;
    // software pipeline over k-blocks: issue the copies of block `kb`, then publish block kb-(STAGES-1)
    for (int kb = 0; kb < nkb + STAGES - 1; ++kb) {
      if (kb < nkb) {
        const int s = kb % STAGES;
        mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
        const uint32_t st = smem_u32(tiles + (size_t)s * STAGE);
        const int64_t k0 = kbeg + (int64_t)kb * kTK;
#pragma unroll
        for (int t = tid; t < kTM * 8; t += kProducerWarps * 32) {
          uint32_t off;
          int64_t src;
          if (g.a_mn) {   // chunk c of k-row kk: 8 consecutive m inside atom c/8
            const int kk = t / (kTM / 8), c = t % (kTM / 8);
            off = (c >> 3) * 8192 + kk * 128 + (((c & 7) ^ (kk & 7)) << 4);
            src = (k0 + kk) * g.a_pitch + m0 + c * 8;
          } else {
            const int r = t >> 3, c = t & 7;
            off = r * 128 + ((c ^ (r & 7)) << 4);
            src = (m0 + r) * g.a_pitch + k0 + c * 8;
          }
          cp_async16(st + off, g.a_hi + src);
          cp_async16(st + A_PLANE + off, g.a_lo + src);
        }
#pragma unroll
        for (int t = tid; t < BN * 8; t += kProducerWarps * 32) {
          uint32_t off;
          int64_t src;
          if (g.b_mn) {
            const int kk = t / (BN / 8), c = t % (BN / 8);
            off = (c >> 3) * 8192 + kk * 128 + (((c & 7) ^ (kk & 7)) << 4);
            src = (k0 + kk) * g.b_pitch + n0 + c * 8;
          } else {
            const int r = t >> 3, c = t & 7;
            off = r * 128 + ((c ^ (r & 7)) << 4);
            src = (n0 + r) * g.b_pitch + k0 + c * 8;
          }
          cp_async16(st + 2 * A_PLANE + off, g.b_hi + src);
          cp_async16(st + 2 * A_PLANE + B_PLANE + off, g.b_lo + src);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (kb >= STAGES - 1) {
        const int j = kb - (STAGES - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[j % STAGES]);
      }
    }
    // ---- epilogue (same as variant 1) ----
    if (nkb > 0) {
      mbar_wait(&accum_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int sub = warp & 3;
    const int half = warp >> 2;
    constexpr int HALVES = BN >= 64 ? 2 : 1;
    constexpr int COLS = BN / HALVES;
    const int64_t gm = m0 + sub * 32 + lane;
    const bool vec_c = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.c) & 15) == 0) &&
                       (!g.bias || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0) && g.splits == 1;
    if (half < HALVES) {
#pragma unroll 1
      for (int c0 = 0; c0 < COLS; c0 += 32) {
        uint32_t r[32];
        if (nkb > 0) {
          const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(half * COLS + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
              "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
                "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (gm < g.m) {
          const int64_t gn0 = n0 + half * COLS + c0;
          if (vec_c && gn0 + 32 <= g.n) {
            float* crow = g.c + gm * g.ldc + gn0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v = make_float4(g.alpha * __uint_as_float(r[j]), g.alpha * __uint_as_float(r[j + 1]),
                                     g.alpha * __uint_as_float(r[j + 2]), g.alpha * __uint_as_float(r[j + 3]));
              if (g.accumulate) {
                const float4 o = *reinterpret_cast<const float4*>(crow + j);
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
              }
              if (g.bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + gn0 + j));
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
              }
              v.x = act_apply(v.x, g.act); v.y = act_apply(v.y, g.act);
              v.z = act_apply(v.z, g.act); v.w = act_apply(v.w, g.act);
              *reinterpret_cast<float4*>(crow + j) = v;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int64_t gn = gn0 + j;
              if (gn < g.n) {
                float v = g.alpha * __uint_as_float(r[j]);
                if (g.splits > 1) {
                  g.ws[((int64_t)blockIdx.z * g.m + gm) * g.n + gn] = v;
                } else {
                  if (g.accumulate) v += g.c[gm * g.ldc + gn];
                  if (g.bias) v += g.bias[gn];
                  g.c[gm * g.ldc + gn] = act_apply(v, g.act);
                }
              }
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    const uint32_t idesc = umma_idesc(BN) | (g.a_mn ? (1u << 15) : 0u) | (g.b_mn ? (1u << 16) : 0u);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&full_bar[s], (kb / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE);
        const uint32_t a_hi = sa, a_lo = sa + A_PLANE, b_hi = sa + 2 * A_PLANE, b_lo = sa + 2 * A_PLANE + B_PLANE;
#pragma unroll
        for (int ks = 0; ks < kTK / 16; ++ks) {
          const uint64_t dah = umma_desc_any(a_hi, ks, g.a_mn), dal = umma_desc_any(a_lo, ks, g.a_mn);
          const uint64_t dbh = umma_desc_any(b_hi, ks, g.b_mn), dbl = umma_desc_any(b_lo, ks, g.b_mn);
          umma_f16(tmem_base, dah, dbh, idesc, (kb | ks) ? 1u : 0u);
          umma_f16(tmem_base, dah, dbl, idesc, 1u);
          umma_f16(tmem_base, dal, dbh, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);
        if (kb == nkb - 1) umma_commit(&accum_bar);
      }
      __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == kProducerWarps) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
  }
}

template <int BN, int STAGES>
static cudaError_t launch_planes(const PlaneArgs& pa, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (2 * kTM * 128 + 2 * BN * 128) + 1024;
  dim3 grid((unsigned)ceil_div(pa.n, BN), (unsigned)ceil_div(pa.m, kTM), (unsigned)pa.splits);
  auto kern = gemm_planes_kernel<BN, STAGES>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, kTcThreads, smem, st>>>(pa);
  return cudaGetLastError();
}


// ================================================================================================
// Variant 4: persistent, warp-specialised, optionally CTA-PAIRED (tcgen05 cta_group::2) planes GEMM.
//
// The split-bf16 scheme moves 2x the operand bytes of a plain bf16 GEMM for 3x its MMAs, so operand
// delivery (L2 -> shared memory), not the tensor pipe, is what a 128x256 single-CTA tile runs out of
// (96 KB per 64-deep k-block per SM, ~11 TB/s over 148 SMs).  A CTA pair computes a 256 x BN tile with
// each CTA staging its own 128 rows of A and only HALF of the B tile (the pair's tensor cores read both
// halves), i.e. 64 KB per k-block per SM at BN = 256 - which also leaves room for a 3-deep pipeline.
//   warps 0-7  : epilogue (TMEM -> registers -> 32x32 transposition through a swizzled staging block -> global,
//                128 contiguous bytes per row); warp & 3 = TMEM sub-partition, warp >> 2 = column half
//   producers  : TMA (default): one elected thread arms the stage barrier and issues cp.async.bulk.tensor.2d loads
//                of the four plane slices; in a CTA pair both CTAs' loads complete on the LEADER's barrier
//                (.cta_group::2).  GEN = 1 / 2: eight warps GENERATE the A tile in shared memory (CIN outer product /
//                DIN attention input) while B arrives by TMA.  B2CTR_TC_TMA=0: four warps of 16-byte cp.async.
//   last warp  : TMEM allocation; lane 0 of the LEADER CTA issues every tcgen05.mma of the pair
// TMEM holds two BN-column accumulators: the epilogue of tile i overlaps the main loop of tile i+1.
// Measured anatomy of a k-block (tools/gemm_sweep.py, B2CTR_TC_DEBUG knock-outs): the stage round trip alone
// (commit -> empty -> producer -> full -> issuer, nothing loaded or multiplied) costs 0.25 us per k-block over 4 stages.
// Persistent: grid = min(#tiles, #SMs / NCTA) clusters; tile = cluster id + j * #clusters, N-tile fastest
// (neighbouring clusters share the A rows through L2).
// ================================================================================================
constexpr int kWsEpilogueWarps = 8;
constexpr int kWsProducers = 4;   // warps

struct WsArgs {
  PlaneArgs p;
  int tiles_m, tiles_n;   // cluster tiles
  int64_t ntiles;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster.  Default (.release.cta)
// semantics on purpose: a cluster-scope release drains every outstanding cp.async of the warp first, which
// serialises the stage ring (measured: the peer CTA's producers spent >50% of their time in that arrive); the
// data being published has already landed (cp.async.wait_group) and been proxy-fenced by the caller.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
template <int NCTA>
__device__ __forceinline__ void umma_f16_ws(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                            uint32_t accumulate) {
  if constexpr (NCTA == 1) {
    umma_f16(tmem_d, da, db, idesc, accumulate);
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
template <int NCTA>
__device__ __forceinline__ void umma_commit_ws(uint64_t* bar) {
  if constexpr (NCTA == 1) {
    umma_commit(bar);
  } else {   // arrives on `bar` in BOTH CTAs of the pair once the MMAs issued so far have completed
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
  }
}

// ---- TMA (cp.async.bulk.tensor) producer primitives --------------------------------------------------
// One elected thread arms the stage's mbarrier with the bytes that will land (expect_tx) and issues the
// tiled bulk copies; the hardware writes the 128-byte-swizzled rows itself (CU_TENSOR_MAP_SWIZZLE_128B is
// exactly the `chunk ^ (row & 7)` layout the UMMA descriptors above expect) and completes the barrier.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// CTA pair: the load issued by either CTA of the pair signals the LEADER's mbarrier (shared::cluster address `bar`),
// so the MMA issuer waits on one barrier per stage and no warp has to relay "the peer's half has landed".
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// fp32 pair -> bf16 hi pair + bf16 lo pair (v = hi + lo up to 2^-17): two cvt.rn.bf16x2.f32 + four ALU ops
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& h, uint32_t& l) {
  const __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);        // .x = v0 in the low half
  h = *reinterpret_cast<const uint32_t*>(&hh);
  const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xffff0000u);
  const __nv_bfloat162 ll = __floats2bfloat162_rn(v0 - h0, v1 - h1);
  l = *reinterpret_cast<const uint32_t*>(&ll);
}
__device__ __forceinline__ void store_chunk(const float (&v)[8], unsigned char* hi_row, unsigned char* lo_row, int off) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_pair(v[2 * e], v[2 * e + 1], h[e], l[e]);
  *reinterpret_cast<uint4*>(hi_row + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_row + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// Generated A operand, CIN: one producer thread = half a row r of the outer product per stage: 32 consecutive
// q = i*hp + j starting at q0 (a multiple of 32; hp is 32 or a multiple of 64, so the 32 columns share one i), split
// into bf16 hi/lo and stored as four 16-byte chunks [c0, c0+4) of the 128-byte-swizzled row `rr`.
// Two steps, so that the producer loop can issue the global loads of k-block kb + 1 before it multiplies / splits /
// stores k-block kb.
struct GenRegs {       // one register image for both generators (only one of them runs in a launch)
  float4 x[8];         // CIN: 32 values of X_k;  attention: 32 key values
  float a;             // CIN: T0[r, i]
  const float* q;      // attention: the row's query values
  bool ok;             // attention: row inside the matrix
};
__device__ __forceinline__ void cin_load_half(const PlaneArgs& g, int64_t r, int q0, GenRegs& o) {
  const bool row_ok = r < g.cin_rows;
  const float* xk = g.cin_xk + r * g.cin_ldk;
  const int hp = g.cin_hp, h = g.cin_h;
  const int i = q0 / hp, j = q0 - i * hp;          // hp is 32 or a multiple of 64: the 32 columns share one i
  o.a = (row_ok && i < g.cin_m) ? __ldg(g.cin_t0 + r * g.cin_ld0 + i) : 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int jj = j + 8 * c;
    o.x[2 * c] = (row_ok && jj < h) ? __ldg(reinterpret_cast<const float4*>(xk + jj)) : make_float4(0.f, 0.f, 0.f, 0.f);
    o.x[2 * c + 1] = (row_ok && jj + 4 < h) ? __ldg(reinterpret_cast<const float4*>(xk + jj) + 1)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// multiply / split / store the half row held in `io` (k-block kb) and refill each register pair with the operands
// of k-block kb + 1 as soon as it has been consumed: a rolling prefetch that costs no extra registers.
__device__ __forceinline__ void cin_emit_half(const PlaneArgs& g, GenRegs& io, int q0, unsigned char* hi_row,
                                              unsigned char* lo_row, int rr, int c0, bool has_next, int64_t r_next,
                                              int q0_next, bool reload_x) {
  const int hp = g.cin_hp, h = g.cin_h;
  const int j = q0 - (q0 / hp) * hp;
  const float a = io.a;
  const bool row_ok = has_next && r_next < g.cin_rows;
  const float* xk = g.cin_xk + r_next * g.cin_ldk;
  const int in_ = q0_next / hp, jn = q0_next - in_ * hp;
  if (has_next) io.a = (row_ok && in_ < g.cin_m) ? __ldg(g.cin_t0 + r_next * g.cin_ld0 + in_) : 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float v[8] = {a * io.x[2 * c].x, a * io.x[2 * c].y, a * io.x[2 * c].z, a * io.x[2 * c].w,
                  a * io.x[2 * c + 1].x, a * io.x[2 * c + 1].y, a * io.x[2 * c + 1].z, a * io.x[2 * c + 1].w};
    if (reload_x) {       // (the next k-block may need the same 32 values of X_k: they then stay where they are)
      const int jj = jn + 8 * c;
      io.x[2 * c] = (row_ok && jj < h) ? __ldg(reinterpret_cast<const float4*>(xk + jj)) : make_float4(0.f, 0.f, 0.f, 0.f);
      io.x[2 * c + 1] = (row_ok && jj + 4 < h) ? __ldg(reinterpret_cast<const float4*>(xk + jj) + 1)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (h & 7) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (j + 8 * c + e >= h) v[e] = 0.f;
    }
    store_chunk(v, hi_row, lo_row, ((c0 + c) ^ (rr & 7)) << 4);
  }
}

// DIN local-activation-unit input (deepctr/layers/core.py:96-101), generated the same way: row r = (b, t),
//   A[r, :] = [ q_b , k_bt , q_b - k_bt , q_b * k_bt ]   (4 segments of E columns; E % 8 == 0)
// cin_t0 = queries [B, ld0], cin_xk = keys (sample stride cin_ldk, row stride E), cin_m = T, cin_h = E.
// (generic E: not inlined - one copy instead of eight in the producer loop)
__device__ __noinline__ void att_generate_half(const PlaneArgs& g, int64_t r, int col0, unsigned char* hi_row,
                                                  unsigned char* lo_row, int rr, int c0) {
  const int T = g.cin_m, E = g.cin_h;
  const bool row_ok = r < g.cin_rows;
  const uint32_t bu = row_ok ? (uint32_t)r / (uint32_t)T : 0u;      // rows < 2^31 (checked on the host)
  const int64_t b = bu;
  const int t = row_ok ? (int)((uint32_t)r - bu * (uint32_t)T) : 0;
  const float* q = g.cin_t0 + b * g.cin_ld0;
  const float* k = g.cin_xk + b * g.cin_ldk + (int64_t)t * E;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int col = col0 + 8 * c;
    const int seg = col / E, e = col - seg * E;
    float v[8];
    if (row_ok && seg < 4) {
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
      if (seg != 1) { a0 = __ldg(reinterpret_cast<const float4*>(q + e)); a1 = __ldg(reinterpret_cast<const float4*>(q + e) + 1); }
      if (seg != 0) { b0 = __ldg(reinterpret_cast<const float4*>(k + e)); b1 = __ldg(reinterpret_cast<const float4*>(k + e) + 1); }
      const float qa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float ka[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int jx = 0; jx < 8; ++jx)
        v[jx] = seg == 0 ? qa[jx] : seg == 1 ? ka[jx] : seg == 2 ? __fsub_rn(qa[jx], ka[jx]) : __fmul_rn(qa[jx], ka[jx]);
    } else {
#pragma unroll
      for (int jx = 0; jx < 8; ++jx) v[jx] = 0.f;
    }
    store_chunk(v, hi_row, lo_row, ((c0 + c) ^ (rr & 7)) << 4);
  }
}
// Attention input with the thread's 32 key values RESIDENT in registers (64 % E == 0, E >= 32: a thread's e-range
// is the same in every k-block).  A forward tile (4E / 64 k-blocks over the same rows) then loads its keys once - and
// the keys of the NEXT tile (or, kernel gradient, of the next k-block's rows) are loaded into the same registers as
// soon as the last use of each pair has issued: the loads stay in flight across the stage hand-over.  The query
// values come from L1 (a CTA's 128 rows belong to 3-4 samples).
__device__ __forceinline__ void att_locate(const PlaneArgs& g, int64_t r, int e0, const float*& q, const float*& k,
                                           bool& ok) {
  const int T = g.cin_m, E = g.cin_h;
  ok = r < g.cin_rows;
  const uint32_t bu = ok ? (uint32_t)r / (uint32_t)T : 0u;
  const int t = ok ? (int)((uint32_t)r - bu * (uint32_t)T) : 0;
  q = g.cin_t0 + (int64_t)bu * g.cin_ld0 + e0;
  k = g.cin_xk + (int64_t)bu * g.cin_ldk + (int64_t)t * E + e0;
}
__device__ __forceinline__ void att_load_keys(const PlaneArgs& g, int64_t r, int col0, GenRegs& o) {
  const int E = g.cin_h;
  const float* k;
  att_locate(g, r, col0 - (col0 / E) * E, o.q, k, o.ok);
#pragma unroll
  for (int c = 0; c < 8; ++c)
    o.x[c] = o.ok ? __ldg(reinterpret_cast<const float4*>(k) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void att_emit_half(const PlaneArgs& g, GenRegs& io, int col0, unsigned char* hi_row,
                                              unsigned char* lo_row, int rr, int c0, bool reload, int64_t r_next) {
  const int E = g.cin_h;
  const int seg = col0 / E;
  const float* q = io.q;
  const bool ok = io.ok;
  const float* kn = nullptr;
  bool okn = false;
  if (reload) att_locate(g, r_next, col0 - seg * E, io.q, kn, okn);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 k0 = io.x[2 * c], k1 = io.x[2 * c + 1];
    float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
    if (ok && seg != 1 && seg < 4) {
      q0 = __ldg(reinterpret_cast<const float4*>(q) + 2 * c);
      q1 = __ldg(reinterpret_cast<const float4*>(q) + 2 * c + 1);
    }
    if (reload) {
      io.x[2 * c] = okn ? __ldg(reinterpret_cast<const float4*>(kn) + 2 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      io.x[2 * c + 1] = okn ? __ldg(reinterpret_cast<const float4*>(kn) + 2 * c + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float qa[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    const float ka[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    float v[8];
#pragma unroll
    for (int jx = 0; jx < 8; ++jx)
      v[jx] = seg == 0 ? qa[jx] : seg == 1 ? ka[jx] : seg == 2 ? __fsub_rn(qa[jx], ka[jx])
            : seg == 3 ? __fmul_rn(qa[jx], ka[jx]) : 0.f;       // seg >= 4: columns of the M padding (kernel gradient)
    store_chunk(v, hi_row, lo_row, ((c0 + c) ^ (rr & 7)) << 4);
  }
  if (reload) io.ok = okn;
}


// generated-operand kernels run 8 producer warps (two threads per generated row), the others 4
template <bool GENERATED>
struct WsLayout {
  static constexpr int kProducers = GENERATED ? 8 : kWsProducers;
  static constexpr int kMmaWarp = kWsEpilogueWarps + kProducers;
  static constexpr int kThreads = (kMmaWarp + 1) * 32;
};

// GEN: 0 = both operands from memory; 1 = A generated as the CIN outer product; 2 = A generated as the DIN
// attention input (one instantiation per generator: the code of the other one would only fill the instruction cache)
template <int BN, int STAGES, int NCTA, bool TMA, int GEN = 0, bool FOLD = false>
__global__ void __launch_bounds__(WsLayout<GEN != 0>::kThreads, 1)
    gemm_planes_ws_kernel(const __grid_constant__ WsArgs w, const __grid_constant__ CUtensorMap tm_ah,
                          const __grid_constant__ CUtensorMap tm_al, const __grid_constant__ CUtensorMap tm_bh,
                          const __grid_constant__ CUtensorMap tm_bl) {
  const PlaneArgs& g = w.p;
  constexpr int BNH = BN / NCTA;             // rows of the B tile staged by one CTA
  constexpr int A_PLANE = kTM * 128;
  constexpr int B_PLANE = BNH * 128;
  constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                                          ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], peer_full[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  unsigned char* epi_stage = tiles + (size_t)STAGES * STAGE;      // 8 epilogue warps x 4 KB (not in FOLD kernels)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = NCTA == 1 ? 0u : cluster_ctarank();
  const int64_t cluster_id = blockIdx.x / NCTA, nclusters = gridDim.x / NCTA;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      // cp.async producers: one deferred arrival per producer thread of THIS CTA; TMA: one arrive.expect_tx
      // CIN: the B planes arrive by TMA (1 arrive.expect_tx) + one arrival per generating thread
      // generated A: one group of 128 threads per stage + the TMA expect_tx of the B planes
      mbar_init(&full_bar[s], GEN != 0 ? 1 + (g.gen_groups == 2 ? 128 : 256) : (TMA ? 1 : kWsProducers * 32));
      mbar_init(&peer_full[s], 1);                    // leader only: the peer CTA's half of the stage landed
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], kWsEpilogueWarps * NCTA);   // used on the leader only
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WsLayout<GEN != 0>::kMmaWarp) {
    if constexpr (NCTA == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (NCTA > 1) cluster_sync_all();      // peer barriers initialised before any remote arrive
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  // tile -> coordinates (N tile fastest, then M, then the split-K slice)
  auto decode = [&](int64_t tile, int64_t& mt, int64_t& nt, int64_t& kbeg, int& nkb) {
    if constexpr (FOLD) {
      // a cluster owns whole ROW BLOCKS and walks all their N tiles in turn (its epilogue accumulates across
      // them): work item j of cluster c is (row block c + (j / tiles_n) * nclusters, N tile j % tiles_n)
      const int64_t j = tile / nclusters;
      nt = j % w.tiles_n;
      mt = tile % nclusters + (j / w.tiles_n) * nclusters;
      kbeg = 0;
      nkb = mt < w.tiles_m ? (int)(g.k_pad / kTK) : 0;
      return;
    }
    nt = tile % w.tiles_n;
    const int64_t rest = tile / w.tiles_n;
    mt = rest % w.tiles_m;
    const int64_t z = rest / w.tiles_m;
    kbeg = z * g.k_per_split;
    const int64_t kend = kbeg + g.k_per_split < g.k_pad ? kbeg + g.k_per_split : g.k_pad;
    nkb = kend > kbeg ? (int)((kend - kbeg) / kTK) : 0;
  };

  constexpr int kMmaWarp = WsLayout<GEN != 0>::kMmaWarp;
  if (GEN != 0 && warp >= kWsEpilogueWarps && warp < kMmaWarp) {
    // ------------------------------------------------------------------------------ generating producers
    // 8 warps in TWO GROUPS of 128 threads; group g owns the stages with (stage counter & 1) == g and generates a
    // whole row (64 columns) per thread for them.  Two stages are therefore in production at any time: while one
    // group waits for its global loads (the producer is load-latency-bound: ncu source page, FMUL on the loaded
    // operands holds the stall samples), the other one multiplies / splits / stores.  B (filter / dY planes)
    // arrives by TMA, issued by thread 0 of the group that owns the stage.
    const int tid = threadIdx.x - kWsEpilogueWarps * 32;
    const int grp = tid >> 7, t128 = tid & 127;
    if (t128 == 0) { tma_prefetch_desc(&tm_bh); tma_prefetch_desc(&tm_bl); }
    uint32_t it = 0;
    const bool att_res = GEN == 2 && g.cin_h >= 32 && 64 % g.cin_h == 0;
    if (GEN == 1 || (att_res && g.gen_groups == 1)) {
      // 256 threads per stage (two per generated row), software-pipelined over the FLATTENED sequence of k-blocks
      // of all the tiles of this CTA: the operands of the next k-block (the next tile's first one included) are
      // requested while the current one is multiplied / split / stored, and stay in flight across the stage
      // hand-over.  The generator was load-latency-bound (ncu source page: the stall samples sit on the first use
      // of the loaded operands); with K = 4E = 256 the attention GEMM has only 4 k-blocks per tile.
      const int half = tid & 1;
      const int atom = tid >> 7, rr = g.a_mn ? (tid & 127) >> 1 : tid >> 1;
      const int soff = (g.a_mn ? atom * 8192 : 0) + rr * 128;
      // (32-bit coordinates: rows, columns and tile counts of the generated-operand GEMMs fit 31 bits, checked on
      // the host; the state of two k-blocks stays in registers next to the prefetched operands)
      const int ntiles = (int)w.ntiles, stride = (int)nclusters;
      auto dec = [&](int tile, int& m0, int& n0, int& kbeg, int& nkb) {
        int64_t mt, nt, kb64;
        decode(tile, mt, nt, kb64, nkb);
        m0 = (int)((mt * NCTA + cta_rank) * kTM);
        n0 = (int)(nt * BN + (int64_t)cta_rank * BNH);
        kbeg = (int)kb64;
      };
      int tile = (int)cluster_id, m0 = 0, n0 = 0, kbeg = 0, nkb = 0, kb = 0;
      for (; tile < ntiles; tile += stride) {            // first tile with work
        dec(tile, m0, n0, kbeg, nkb);
        if (nkb > 0) break;
      }
      // this thread's row and first column in the k-block at (m0, k0)
      auto row_of = [&](int m0_, int k0_) { return g.a_mn ? k0_ + rr : m0_ + rr; };
      auto col_of = [&](int m0_, int k0_) { return g.a_mn ? m0_ + atom * 64 + half * 32 : k0_ + half * 32; };
      GenRegs gr;
      if (tile < ntiles) {
        if constexpr (GEN == 1) cin_load_half(g, row_of(m0, kbeg), col_of(m0, kbeg), gr);     // (k_of(kbeg, 0, .) == kbeg)
        else att_load_keys(g, row_of(m0, kbeg), col_of(m0, kbeg), gr);
      }
      // CIN forward, hp = nj * 64: the k-blocks of a tile are walked j-block-major (all i for the first 64 columns of
      // X_k, then all i for the next 64 ...).  A thread's 32 values of X_k are then the same for m consecutive
      // k-blocks and stay in registers; only T0[r, i] is fetched per k-block.  (The B planes are fetched at the same
      // k0, and the order of the K accumulation is free.)
      const int nj = (GEN == 1 && !g.a_mn && g.splits == 1 && g.cin_hp >= 128) ? g.cin_hp / kTK : 1;
      auto k_of = [&](int kbeg_, int kb_, int nkb_) {
        if (nj == 1) return kbeg_ + kb_ * kTK;
        const int ni = nkb_ / nj;                      // = m (k_pad = m * hp)
        const int jb = kb_ / ni, i = kb_ - jb * ni;
        return (i * nj + jb) * kTK;
      };
      while (tile < ntiles) {
        const int k0 = k_of(kbeg, kb, nkb);
        // the k-block after this one
        int tile_n = tile, m0_n = m0, n0_n = n0, kbeg_n = kbeg, nkb_n = nkb, kb_n = kb + 1;
        if (kb_n >= nkb) {
          kb_n = 0;
          for (tile_n = tile + stride; tile_n < ntiles; tile_n += stride) {
            dec(tile_n, m0_n, n0_n, kbeg_n, nkb_n);
            if (nkb_n > 0) break;
          }
        }
        const bool has_next = tile_n < ntiles;
        const int k0_n = k_of(kbeg_n, kb_n, nkb_n);
        const int s = it % STAGES;
        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        unsigned char* stp = tiles + (size_t)s * STAGE;
        if (tid == 0) {
          uint64_t* bar = &full_bar[s];
          mbar_expect_tx(bar, (uint32_t)(2 * B_PLANE));
          const uint32_t st = smem_u32(stp);
          if (g.b_mn) {
#pragma unroll
            for (int j = 0; j < (BNH >= 64 ? BNH / 64 : 1); ++j) {
              tma_load_2d(st + 2 * A_PLANE + j * 8192, &tm_bh, n0 + 64 * j, k0, bar);
              tma_load_2d(st + 2 * A_PLANE + B_PLANE + j * 8192, &tm_bl, n0 + 64 * j, k0, bar);
            }
          } else {
            tma_load_2d(st + 2 * A_PLANE, &tm_bh, k0, n0, bar);
            tma_load_2d(st + 2 * A_PLANE + B_PLANE, &tm_bl, k0, n0, bar);
          }
        }
        if constexpr (GEN == 1) {
          const int r = row_of(m0, k0), r_n = row_of(m0_n, k0_n), q = col_of(m0, k0), q_n = col_of(m0_n, k0_n);
          cin_emit_half(g, gr, q, stp + soff, stp + A_PLANE + soff, rr, half * 4, has_next, r_n, q_n,
                        has_next && (r_n != r || q_n % g.cin_hp != q % g.cin_hp));
        } else {
          const int r = row_of(m0, k0), r_n = row_of(m0_n, k0_n);
          att_emit_half(g, gr, col_of(m0, k0), stp + soff, stp + A_PLANE + soff, rr, half * 4, has_next && r_n != r, r_n);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&full_bar[s]);
        ++it;
        tile = tile_n; m0 = m0_n; n0 = n0_n; kbeg = kbeg_n; nkb = nkb_n; kb = kb_n;
      }
    } else
    for (int64_t tile = cluster_id; tile < w.ntiles; tile += nclusters) {
      int64_t mt, nt, kbeg;
      int nkb;
      decode(tile, mt, nt, kbeg, nkb);
      const int64_t m0 = (mt * NCTA + cta_rank) * kTM;
      const int32_t n0 = (int32_t)(nt * BN + (int64_t)cta_rank * BNH);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const bool two = g.gen_groups == 2;
        if (two && (int)(it & 1) != grp) continue;
        const int s = it % STAGES;
        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        unsigned char* stp = tiles + (size_t)s * STAGE;
        const int64_t k0 = kbeg + (int64_t)kb * kTK;
        if (two ? t128 == 0 : tid == 0) {
          uint64_t* bar = &full_bar[s];
          mbar_expect_tx(bar, (uint32_t)(2 * B_PLANE));
          const uint32_t st = smem_u32(stp);
          if (g.b_mn) {
#pragma unroll
            for (int j = 0; j < (BNH >= 64 ? BNH / 64 : 1); ++j) {
              tma_load_2d(st + 2 * A_PLANE + j * 8192, &tm_bh, n0 + 64 * j, (int32_t)k0, bar);
              tma_load_2d(st + 2 * A_PLANE + B_PLANE + j * 8192, &tm_bl, n0 + 64 * j, (int32_t)k0, bar);
            }
          } else {
            tma_load_2d(st + 2 * A_PLANE, &tm_bh, (int32_t)k0, n0, bar);
            tma_load_2d(st + 2 * A_PLANE + B_PLANE, &tm_bl, (int32_t)k0, n0, bar);
          }
        }
        if (two) {
          if (g.a_mn) {      // A^T: M = q (two 64-wide atoms), K = r: thread -> (atom, k-row)
            const int atom = t128 >> 6, rr = t128 & 63;
            unsigned char* row = stp + atom * 8192 + rr * 128;
            att_generate_half(g, k0 + rr, (int)(m0 + atom * 64), row, row + A_PLANE, rr, 0);
            att_generate_half(g, k0 + rr, (int)(m0 + atom * 64 + 32), row, row + A_PLANE, rr, 4);
          } else {           // A: M = r, K = q: thread -> row
            unsigned char* row = stp + t128 * 128;
            att_generate_half(g, m0 + t128, (int)k0, row, row + A_PLANE, t128, 0);
            att_generate_half(g, m0 + t128, (int)(k0 + 32), row, row + A_PLANE, t128, 4);
          }
        } else {
          const int half = tid & 1;           // two threads per generated row: 32 of its 64 columns each
          if (g.a_mn) {
            const int atom = tid >> 7, rr = (tid & 127) >> 1;
            att_generate_half(g, k0 + rr, (int)(m0 + atom * 64 + half * 32), stp + atom * 8192 + rr * 128,
                          stp + A_PLANE + atom * 8192 + rr * 128, rr, half * 4);
          } else {
            const int rr = tid >> 1;
            att_generate_half(g, m0 + rr, (int)(k0 + half * 32), stp + rr * 128, stp + A_PLANE + rr * 128, rr, half * 4);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&full_bar[s]);
      }
    }
  } else if (TMA && warp >= kWsEpilogueWarps && warp < kMmaWarp) {
    // ------------------------------------------------------------------------------ TMA producer
    // warp 8, one elected lane: wait for a free stage, arm its barrier with the stage bytes, issue the bulk
    // tensor copies of the four operand planes.  No registers, no per-element instructions, no proxy fence:
    // data moves global -> swizzled shared memory inside the async proxy, where tcgen05.mma reads it.
    if (warp == kWsEpilogueWarps && lane == 0) {
      tma_prefetch_desc(&tm_ah); tma_prefetch_desc(&tm_al); tma_prefetch_desc(&tm_bh); tma_prefetch_desc(&tm_bl);
      uint32_t it = 0;
      for (int64_t tile = cluster_id; tile < w.ntiles; tile += nclusters) {
        int64_t mt, nt, kbeg;
        int nkb;
        decode(tile, mt, nt, kbeg, nkb);
        const int32_t m0 = (int32_t)((mt * NCTA + cta_rank) * kTM);
        const int32_t n0 = (int32_t)(nt * BN + (int64_t)cta_rank * BNH);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          uint64_t* bar = &full_bar[s];
          if (g.debug & 8) { mbar_arrive(bar); continue; }
          // CTA pair: both CTAs' loads complete on the LEADER's barrier, armed there with the bytes of both halves
          if (NCTA == 1 || cta_rank == 0) mbar_expect_tx(bar, (uint32_t)(NCTA * STAGE));
          const uint32_t lbar = NCTA == 1 ? smem_u32(bar) : mapa_u32(smem_u32(bar), 0);
          auto load = [&](uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1) {
            if constexpr (NCTA == 1) tma_load_2d(dst, map, c0, c1, bar);
            else tma_load_2d_pair(dst, map, c0, c1, lbar);
          };
          const uint32_t st = smem_u32(tiles + (size_t)s * STAGE);
          const int32_t k0 = (int32_t)(kbeg + (int64_t)kb * kTK);
          if (g.a_mn) {        // [k, m] planes: one 64(m) x 64(k) box per 64-wide atom
#pragma unroll
            for (int j = 0; j < kTM / 64; ++j) {
              load(st + j * 8192, &tm_ah, m0 + 64 * j, k0);
              load(st + A_PLANE + j * 8192, &tm_al, m0 + 64 * j, k0);
            }
          } else {             // [m, k] planes: one 64(k) x 128(m) box
            load(st, &tm_ah, k0, m0);
            load(st + A_PLANE, &tm_al, k0, m0);
          }
          if (g.b_mn) {
#pragma unroll
            for (int j = 0; j < (BNH >= 64 ? BNH / 64 : 1); ++j) {
              load(st + 2 * A_PLANE + j * 8192, &tm_bh, n0 + 64 * j, k0);
              load(st + 2 * A_PLANE + B_PLANE + j * 8192, &tm_bl, n0 + 64 * j, k0);
            }
          } else {
            load(st + 2 * A_PLANE, &tm_bh, k0, n0);
            load(st + 2 * A_PLANE + B_PLANE, &tm_bl, k0, n0);
          }
        }
      }
    }
  } else if (warp >= kWsEpilogueWarps && warp < kMmaWarp) {
    // ------------------------------------------------------------------------------ producers
    const int tid = threadIdx.x - kWsEpilogueWarps * 32;
    constexpr int NT = kWsProducers * 32;
    int64_t tile = cluster_id, mt = 0, nt = 0, kbeg = 0;
    int nkb = 0, kb = 0;
    while (tile < w.ntiles) {            // first tile with work
      decode(tile, mt, nt, kbeg, nkb);
      if (nkb > 0) break;
      tile += nclusters;
    }
    // Every producer thread hands its copies to the stage's mbarrier (cp.async.mbarrier.arrive.noinc: the
    // arrival fires when the thread's cp.asyncs have landed), so this warp never waits for data - only for a
    // free stage - and all STAGES blocks are genuinely in flight.
    for (uint32_t it = 0; tile < w.ntiles; ++it) {
      const int s = it % STAGES;
      mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
      const uint32_t st = smem_u32(tiles + (size_t)s * STAGE);
      const int64_t k0 = kbeg + (int64_t)kb * kTK;
      const int64_t m0 = (mt * NCTA + cta_rank) * kTM;
      const int64_t n0 = nt * BN + (int64_t)cta_rank * BNH;
#pragma unroll 4
      for (int t = tid; t < kTM * 8; t += NT) {
        uint32_t off;
        int64_t src;
        if (g.a_mn) {
          const int kk = t / (kTM / 8), c = t % (kTM / 8);
          off = (c >> 3) * 8192 + kk * 128 + (((c & 7) ^ (kk & 7)) << 4);
          src = (k0 + kk) * g.a_pitch + m0 + c * 8;
        } else {
          const int r = t >> 3, c = t & 7;
          off = r * 128 + ((c ^ (r & 7)) << 4);
          src = (m0 + r) * g.a_pitch + k0 + c * 8;
        }
        cp_async16(st + off, g.a_hi + src);
        cp_async16(st + A_PLANE + off, g.a_lo + src);
      }
#pragma unroll 4
      for (int t = tid; t < BNH * 8; t += NT) {
        uint32_t off;
        int64_t src;
        if (g.b_mn) {
          const int kk = t / (BNH / 8), c = t % (BNH / 8);
          off = (c >> 3) * 8192 + kk * 128 + (((c & 7) ^ (kk & 7)) << 4);
          src = (k0 + kk) * g.b_pitch + n0 + c * 8;
        } else {
          const int r = t >> 3, c = t & 7;
          off = r * 128 + ((c ^ (r & 7)) << 4);
          src = (n0 + r) * g.b_pitch + k0 + c * 8;
        }
        cp_async16(st + 2 * A_PLANE + off, g.b_hi + src);
        cp_async16(st + 2 * A_PLANE + B_PLANE + off, g.b_lo + src);
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
      if (++kb == nkb) {               // next tile with work
        kb = 0;
        tile += nclusters;
        while (tile < w.ntiles) {
          decode(tile, mt, nt, kbeg, nkb);
          if (nkb > 0) break;
          tile += nclusters;
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------------------ MMA issuer
    if (cta_rank == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)((kTM * NCTA) >> 4) << 24) | (g.a_mn ? (1u << 15) : 0u) |
                             (g.b_mn ? (1u << 16) : 0u);
      uint32_t it = 0, acc_it = 0;
      for (int64_t tile = cluster_id; tile < w.ntiles; tile += nclusters) {
        int64_t mt, nt, kbeg;
        int nkb;
        decode(tile, mt, nt, kbeg, nkb);
        if (nkb == 0) continue;
        const uint32_t ab = acc_it & 1;
        mbar_wait_cluster(&acc_empty[ab], ((acc_it >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + ab * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          if constexpr (NCTA > 1 && !(TMA && GEN == 0)) mbar_wait_cluster(&peer_full[s], (it / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (lane == 0) {
            const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE);
            const uint32_t a_hi = sa, a_lo = sa + A_PLANE, b_hi = sa + 2 * A_PLANE, b_lo = sa + 2 * A_PLANE + B_PLANE;
#pragma unroll
            for (int ks = 0; ks < kTK / 16; ++ks) {
              if (g.debug & 4) break;
              const uint64_t dah = umma_desc_any(a_hi, ks, g.a_mn), dal = umma_desc_any(a_lo, ks, g.a_mn);
              const uint64_t dbh = umma_desc_any(b_hi, ks, g.b_mn), dbl = umma_desc_any(b_lo, ks, g.b_mn);
              umma_f16_ws<NCTA>(tmem_d, dah, dbh, idesc, (kb | ks) ? 1u : 0u);
              umma_f16_ws<NCTA>(tmem_d, dah, dbl, idesc, 1u);
              umma_f16_ws<NCTA>(tmem_d, dal, dbh, idesc, 1u);
            }
            umma_commit_ws<NCTA>(&empty_bar[s]);
            if (kb == nkb - 1) umma_commit_ws<NCTA>(&acc_full[ab]);
          }
          __syncwarp();
        }
        ++acc_it;
      }
    } else if constexpr (!(TMA && GEN == 0)) {
      // peer CTA (cp.async / generated operands): relay "my half of stage s has landed" to the leader, which issues
      // the MMAs of the pair.  With TMA for both operands the peer's loads signal the leader's barrier themselves.
      uint32_t it = 0;
      for (int64_t tile = cluster_id; tile < w.ntiles; tile += nclusters) {
        int64_t mt, nt, kbeg;
        int nkb;
        decode(tile, mt, nt, kbeg, nkb);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          if (lane == 0) mbar_arrive_cluster(&peer_full[s], 0);
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------ epilogue
    // 8 warps: warp & 3 = TMEM sub-partition (32 rows), warp >> 2 = column half of the accumulator
    const int sub = warp & 3, half = warp >> 2;
    constexpr int HALVES = BN >= 64 ? 2 : 1;
    constexpr int COLS = BN / HALVES;
    const bool vec_c = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.c) & 15) == 0) &&
                       (!g.bias || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0);
    const bool vec_ws = (g.n % 4 == 0);
    uint32_t acc_it = 0;
    if constexpr (FOLD) {
      // CIN backward: the accumulator tile is dZ[r, q] (q = i*hp + j) = dY W'^T and is never stored - it is folded
      // onto the two factors of the outer product right here:
      //   dT0[r, i] += sum_j dZ[r, i*hp + j] * xk[r, j]          (one red.add per 32-column group)
      //   dXk[r, j] += sum_i dZ[r, i*hp + j] * t0[r, i]          (registers across the N tiles of the row block,
      //                                                           red.add.v4 once per row block)
      // BN % hp == 0, so a thread's column groups keep their j-range from tile to tile.
      constexpr int GROUPS = COLS / 32;
      float dxk[GROUPS][32];
      for (int64_t tile = cluster_id; tile < w.ntiles; tile += nclusters) {
        int64_t mt, nt, kbeg;
        int nkb;
        decode(tile, mt, nt, kbeg, nkb);
        if (nkb == 0) continue;
        const uint32_t ab = acc_it & 1;
        mbar_wait(&acc_full[ab], (acc_it >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t gm = (mt * NCTA + cta_rank) * kTM + sub * 32 + lane;
        const bool row_ok = gm < g.m;
        const float* t0 = g.cin_t0 + (row_ok ? gm : 0) * g.cin_ld0;
        const float* xk = g.cin_xk + (row_ok ? gm : 0) * g.cin_ldk;
        if (nt == 0) {
#pragma unroll
          for (int gi = 0; gi < GROUPS; ++gi)
#pragma unroll
            for (int jx = 0; jx < 32; ++jx) dxk[gi][jx] = 0.f;
        }
#pragma unroll
        for (int gi = 0; gi < GROUPS; ++gi) {
          const int c0 = half * COLS + gi * 32;
          const int64_t q = nt * BN + c0;
          if (q < g.n) {                   // warp-uniform
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + ab * BN + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                  "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
                  "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                  "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int i = (int)(q / g.cin_hp), j0 = (int)(q - (int64_t)i * g.cin_hp);
            if (row_ok && i < g.cin_m) {
              const float a = __ldg(t0 + i);
              float p = 0.f;
#pragma unroll
              for (int jx = 0; jx < 32; jx += 4) {
                float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j0 + jx < g.cin_h) xv = __ldg(reinterpret_cast<const float4*>(xk + j0 + jx));
                const float d0 = __uint_as_float(r[jx]), d1 = __uint_as_float(r[jx + 1]);
                const float d2 = __uint_as_float(r[jx + 2]), d3 = __uint_as_float(r[jx + 3]);
                if (j0 + jx + 1 >= g.cin_h) xv.y = 0.f;       // ragged h (rows padded to hp: loads stay in bounds)
                if (j0 + jx + 2 >= g.cin_h) xv.z = 0.f;
                if (j0 + jx + 3 >= g.cin_h) xv.w = 0.f;
                p += d0 * xv.x + d1 * xv.y + d2 * xv.z + d3 * xv.w;
                dxk[gi][jx] += d0 * a; dxk[gi][jx + 1] += d1 * a; dxk[gi][jx + 2] += d2 * a; dxk[gi][jx + 3] += d3 * a;
              }
              red_add_f1(g.fold_dt0 + gm * g.cin_ld0 + i, p);
            }
          }
        }
        if (nt == w.tiles_n - 1 && row_ok) {
#pragma unroll
          for (int gi = 0; gi < GROUPS; ++gi) {
            const int j0 = (half * COLS + gi * 32) % g.cin_hp;
            float* dst = g.fold_dxk + gm * g.fold_ldx + j0;
#pragma unroll
            for (int jx = 0; jx < 32; jx += 4)
              if (j0 + jx < g.cin_h)
                red_add_f4(dst + jx, make_float4(dxk[gi][jx], dxk[gi][jx + 1], dxk[gi][jx + 2], dxk[gi][jx + 3]));
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 1 || cta_rank == 0) mbar_arrive(&acc_empty[ab]);
          else mbar_arrive_cluster(&acc_empty[ab], 0);
        }
        ++acc_it;
      }
    } else
    for (int64_t tile = cluster_id; tile < w.ntiles; tile += nclusters) {
      int64_t mt, nt, kbeg;
      int nkb;
      decode(tile, mt, nt, kbeg, nkb);
      const int64_t z = tile / ((int64_t)w.tiles_n * w.tiles_m);
      const uint32_t ab = acc_it & 1;
      if (nkb > 0) {
        mbar_wait(&acc_full[ab], (acc_it >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      // The accumulator arrives lane = row (tcgen05.ld 32x32b): a direct store would put the 32 lanes of every
      // instruction in 32 different rows of C (16 bytes each, half a sector per lane - measured: ~4 us per
      // 256 x 128 tile, the whole cost of a short-K tile).  Each warp therefore turns its 32 x 32 block through a
      // 4 KB XOR-swizzled staging buffer and stores 8 lanes = 128 contiguous bytes per row, 4 rows per instruction.
      const int64_t row_base = (mt * NCTA + cta_rank) * kTM + sub * 32;
      unsigned char* stg = epi_stage + warp * 4096;
      if (half < HALVES) {
#pragma unroll 1
        for (int c0 = half * COLS; c0 < (half + 1) * COLS; c0 += 32) {
          const int64_t gn0 = nt * BN + c0;
          if (gn0 >= g.n) break;           // warp-uniform
          uint32_t r[32];
          if (nkb > 0 && !(g.debug & 2)) {
            const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + ab * BN + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                  "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
                  "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                  "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0u;
          }
          if (row_base >= g.m || (g.debug & 1)) continue;   // warp-uniform: the whole 32-row block is padding
#pragma unroll
          for (int c = 0; c < 8; ++c)      // row `lane`, 16-byte chunk c -> physical chunk c ^ (lane & 7)
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
          __syncwarp();
          const int chunk = lane & 7;
          const int64_t gn = gn0 + chunk * 4;
          const bool full4 = gn + 4 <= g.n;
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g.bias && g.splits == 1 && gn < g.n) {
            if (full4 && vec_c) bb = __ldg(reinterpret_cast<const float4*>(g.bias + gn));
            else {
              bb.x = g.bias[gn];
              if (gn + 1 < g.n) bb.y = g.bias[gn + 1];
              if (gn + 2 < g.n) bb.z = g.bias[gn + 2];
              if (gn + 3 < g.n) bb.w = g.bias[gn + 3];
            }
          }
          const bool relu = g.act == B2CTR_ACT_RELU, other = g.act != B2CTR_ACT_NONE && !relu;
#pragma unroll 2
          for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + (lane >> 3);
            const uint4 u = *reinterpret_cast<const uint4*>(stg + row * 128 + ((chunk ^ (row & 7)) << 4));
            const int64_t gm = row_base + row;
            if (gm >= g.m || gn >= g.n) continue;
            float4 v = make_float4(g.alpha * __uint_as_float(u.x), g.alpha * __uint_as_float(u.y),
                                   g.alpha * __uint_as_float(u.z), g.alpha * __uint_as_float(u.w));
            if (g.splits > 1) {
              float* wrow = g.ws + (z * g.m + gm) * g.n + gn;
              if (vec_ws && full4) *reinterpret_cast<float4*>(wrow) = v;
              else {
                wrow[0] = v.x;
                if (gn + 1 < g.n) wrow[1] = v.y;
                if (gn + 2 < g.n) wrow[2] = v.z;
                if (gn + 3 < g.n) wrow[3] = v.w;
              }
              continue;
            }
            float* crow = g.c + gm * g.ldc + gn;
            const bool vec = vec_c && full4;
            if (g.accumulate) {
              if (vec) {
                const float4 o = *reinterpret_cast<const float4*>(crow);
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
              } else {
                v.x += crow[0];
                if (gn + 1 < g.n) v.y += crow[1];
                if (gn + 2 < g.n) v.z += crow[2];
                if (gn + 3 < g.n) v.w += crow[3];
              }
            }
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            if (relu) {
              v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            } else if (other) {
              v.x = act_apply(v.x, g.act); v.y = act_apply(v.y, g.act);
              v.z = act_apply(v.z, g.act); v.w = act_apply(v.w, g.act);
            }
            if (vec) *reinterpret_cast<float4*>(crow) = v;
            else {
              crow[0] = v.x;
              if (gn + 1 < g.n) crow[1] = v.y;
              if (gn + 2 < g.n) crow[2] = v.z;
              if (gn + 3 < g.n) crow[3] = v.w;
            }
          }
          __syncwarp();                    // the staging block is rewritten by the next column group
        }
      }
      if (nkb > 0) {      // hand the accumulator back to the MMA issuer of the pair
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 1 || cta_rank == 0) mbar_arrive(&acc_empty[ab]);
          else mbar_arrive_cluster(&acc_empty[ab], 0);
        }
        ++acc_it;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (NCTA > 1) cluster_sync_all();      // no CTA leaves while its peer may still signal it
  if (warp == WsLayout<GEN != 0>::kMmaWarp) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (NCTA == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

static int tc_debug() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B2CTR_TC_DEBUG"); v = e ? atoi(e) : 0; }
  return v;
}

struct TmaMaps {
  CUtensorMap ah, al, bh, bl;
  bool ok;
};

template <int BN, int STAGES, int NCTA, bool TMA, int GEN = 0, bool FOLD = false>
static cudaError_t launch_ws_impl(const WsArgs& wa, const TmaMaps& tm, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (2 * kTM * 128 + 2 * (BN / NCTA) * 128) + 1024 +
                          (FOLD ? 0 : kWsEpilogueWarps * 4096);     // + the epilogue's transpose staging
  static_assert(smem + 256 <= 227 * 1024, "stage ring exceeds shared memory");
  auto kern = gemm_planes_ws_kernel<BN, STAGES, NCTA, TMA, GEN, FOLD>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t max_clusters = kNumSMs / NCTA;
  int64_t nclusters = wa.ntiles < max_clusters ? wa.ntiles : max_clusters;
  WsArgs wcopy = wa;
  if (FOLD) {      // clusters own row blocks: ntiles = work-item slots of the round-robin over row blocks
    nclusters = wa.tiles_m < max_clusters ? wa.tiles_m : max_clusters;
    wcopy.ntiles = ceil_div(wa.tiles_m, nclusters) * wa.tiles_n * nclusters;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nclusters * NCTA));
  cfg.blockDim = dim3(WsLayout<GEN != 0>::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, wcopy, tm.ah, tm.al, tm.bh, tm.bl);
}
template <int BN, int STAGES, int NCTA>
static cudaError_t launch_ws(const WsArgs& wa, const TmaMaps& tm, cudaStream_t st) {
  return tm.ok ? launch_ws_impl<BN, STAGES, NCTA, true>(wa, tm, st) : launch_ws_impl<BN, STAGES, NCTA, false>(wa, tm, st);
}

// ---- tensor maps: cuTensorMapEncodeTiled resolved through the runtime's driver entry-point query ---------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tma_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* ev = getenv("B2CTR_TC_TMA");
    if (ev && atoi(ev) == 0) return nullptr;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}
// bf16 matrix [outer, inner] with `pitch` elements between rows; box = box_inner x box_outer, 128-byte swizzle,
// out-of-range box parts read as zero.
static bool tma_map_2d(CUtensorMap* m, const void* base, int64_t inner, int64_t outer, int64_t pitch,
                       int box_inner, int box_outer) {
  EncodeTiledFn enc = tma_encoder();
  if (!enc || ((uintptr_t)base & 15) || (pitch * 2) % 16 || box_inner * 2 > 128 || box_outer > 256) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline int64_t round_up(int64_t v, int64_t q) { return (v + q - 1) / q * q; }
// Padded extent of caller planes: rows to 256 (largest N tile of a K-major B / M tile pair), columns to 128
// (M tile of an MN-major A, N tiles of an MN-major B); narrow matrices (<= 64 columns) pad to one 64-wide
// swizzle atom only - a 128-wide MN-major tile then runs into the next plane row, which only feeds output
// rows/columns >= m/n that are never stored.
static inline int planes_bn(int64_t n, int64_t k) {
  if (n <= 32) return 32;
  if (n <= 64) return 64;
  if (n <= 128 || k <= 4 * kTK) return 128;   // short K: prefer 128-wide tiles, 3 CTAs per SM
  return 256;
}

static int tc_variant(const b2ctr_gemm_t* g) {
  if (g->variant >= 1 && g->variant <= 4) return g->variant;
  static int mode = -1;
  if (mode < 0) {
    const char* ev = getenv("B2CTR_TC_VARIANT");
    mode = ev ? atoi(ev) : 4;
  }
  return (mode >= 1 && mode <= 4) ? mode : 4;
}

size_t gemm_bf16x3_workspace_bytes(const b2ctr_gemm_t* g) {
  size_t splitk = g->split_k > 1 ? (size_t)g->split_k * g->m * g->n * sizeof(float) : 0;
  if (tc_variant(g) == 1) return splitk;
  const int bn = planes_bn(g->n, g->k / (g->split_k > 1 ? g->split_k : 1));
  const int64_t kp = round_up(g->k > 0 ? g->k : 1, kTK), mp = round_up(g->m, 2 * kTM), np = round_up(g->n, 256);
  return splitk + (size_t)(mp + np) * kp * 2 * sizeof(__nv_bfloat16) + 512;
}

static b2ctr_status_t gemm_planes(const b2ctr_gemm_t* g, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const size_t need = gemm_bf16x3_workspace_bytes(g);
  if (!workspace || workspace_bytes < need) {
    set_error("gemm(bf16x3): needs %zu workspace bytes, got %zu", need, workspace_bytes);
    return B2CTR_ERR_WORKSPACE;
  }
  const int splits = g->split_k > 1 ? g->split_k : 1;
  const bool ws_kernel = tc_variant(g) == 4;
  int bn = ws_kernel ? (g->n <= 32 ? 32 : g->n <= 64 ? 64 : g->n <= 128 ? 128 : 256)
                     : planes_bn(g->n, g->k / (g->split_k > 1 ? g->split_k : 1));
  if (ws_kernel && bn == 256) {
    // ragged N (dgrad of the first DNN layer: N = 845): 256-wide tiles pad it to 1024 (17 % of the MMAs and of the
    // epilogue work are dead), 128-wide tiles to 896 (6 %).  Take the narrower tile when it saves > 8 % of the tile area.
    static int tail = -1;
    if (tail < 0) { const char* ev = getenv("B2CTR_TC_BN_TAIL"); tail = ev ? atoi(ev) : 1; }
    const int64_t a256 = round_up(g->n, 256), a128 = round_up(g->n, 128);
    if (tail && (a256 - a128) * 100 > 8 * a256) bn = 128;
  }
  const int64_t kp = round_up(g->k > 0 ? g->k : 1, kTK), mp = round_up(g->m, 2 * kTM);
  unsigned char* w = (unsigned char*)workspace;
  float* ws = (float*)w;
  w += splits > 1 ? (size_t)splits * g->m * g->n * sizeof(float) : 0;
  w = (unsigned char*)(((uintptr_t)w + 255) & ~(uintptr_t)255);
  // variant 2: always K-major planes (row-contiguous sources go through a transposing split);
  // variant 3: row-contiguous sources keep their layout (MN-major planes) and the UMMA descriptors do the
  // transposition - the same planes then serve every GEMM that reads the tensor (forward / dgrad / wgrad),
  // which is what caller-provided planes (g->a_planes / g->b_planes) exploit.
  const bool mn_ok = tc_variant(g) >= 3;
  const bool a_kc = !g->trans_a, b_kc = g->trans_b != 0;
  const int64_t a_sr = g->trans_a ? g->k : g->m, a_sc = g->trans_a ? g->m : g->k;   // stored rows / cols
  const int64_t b_sr = g->trans_b ? g->n : g->k, b_sc = g->trans_b ? g->k : g->n;
  const bool a_given = g->a_planes && (a_kc || mn_ok);
  bool b_given = g->b_planes && (b_kc || (mn_ok && bn >= 64));
  if (b_given && !b_kc && planes_cols_pad(b_sc) % bn != 0) bn = planes_cols_pad(b_sc) == 64 ? 64 : 128;   // N tiles inside the pad
  const int64_t np = round_up(g->n, bn);
  const int a_mn = a_given ? !a_kc : ((!a_kc && mn_ok) ? 1 : 0);
  const int b_mn = b_given ? !b_kc : ((!b_kc && mn_ok && bn >= 64) ? 1 : 0);   // an MN atom is 64 elements wide
  __nv_bfloat16 *a_hi, *a_lo, *b_hi, *b_lo;
  int64_t a_pitch, b_pitch;
  int64_t a_prows, b_prows;      // rows of the plane matrices as stored (TMA extents)
  auto split_k = [&](const float* p, int64_t ld, int64_t rows, int64_t cols, int64_t rows_pad, int64_t cols_pad,
                     __nv_bfloat16* hi, __nv_bfloat16* lo) {      // planes[r, c] = p[r*ld + c]
    const int vec = (ld % 4 == 0) && (((uintptr_t)p & 15) == 0);
    split_planes_kernel<<<grid_for(rows_pad * (cols_pad / 8), 256, 8), 256, 0, st>>>(p, ld, rows, cols, rows_pad,
                                                                                   cols_pad, hi, lo, vec);
  };
  auto split_t = [&](const float* p, int64_t ld, int64_t rows, int64_t rows_pad, __nv_bfloat16* hi,
                     __nv_bfloat16* lo) {                        // planes[r, k] = p[k*ld + r]
    dim3 grid((unsigned)ceil_div(rows_pad, 64), (unsigned)(kp / 64));
    split_planes_t_kernel<<<grid, 256, 0, st>>>(p, ld, rows, g->k, rows_pad, kp, hi, lo);
  };
  if (a_given) {
    const int64_t rp = round_up(a_sr, 256), cp = planes_cols_pad(a_sc);
    a_hi = (__nv_bfloat16*)g->a_planes; a_lo = a_hi + rp * cp; a_pitch = cp; a_prows = rp;
  } else {
    a_hi = (__nv_bfloat16*)w; a_lo = a_hi + mp * kp; w = (unsigned char*)(a_lo + mp * kp);
    a_pitch = a_mn ? mp : kp; a_prows = a_mn ? kp : mp;
    if (a_kc) split_k(g->a, g->lda, g->m, g->k, mp, kp, a_hi, a_lo);
    else if (a_mn) split_k(g->a, g->lda, g->k, g->m, kp, mp, a_hi, a_lo);
    else split_t(g->a, g->lda, g->m, mp, a_hi, a_lo);
    B2_CHECK_LAUNCH("b2ctr_gemm(bf16x3 split A)");
  }
  if (b_given) {
    const int64_t rp = round_up(b_sr, 256), cp = planes_cols_pad(b_sc);
    b_hi = (__nv_bfloat16*)g->b_planes; b_lo = b_hi + rp * cp; b_pitch = cp; b_prows = rp;
  } else {
    b_hi = (__nv_bfloat16*)w; b_lo = b_hi + np * kp;
    b_pitch = b_mn ? np : kp; b_prows = b_mn ? kp : np;
    if (b_kc) split_k(g->b, g->ldb, g->n, g->k, np, kp, b_hi, b_lo);
    else if (b_mn) split_k(g->b, g->ldb, g->k, g->n, kp, np, b_hi, b_lo);
    else split_t(g->b, g->ldb, g->n, np, b_hi, b_lo);
    B2_CHECK_LAUNCH("b2ctr_gemm(bf16x3 split B)");
  }
  PlaneArgs pa;
  pa.a_hi = a_hi; pa.a_lo = a_lo; pa.b_hi = b_hi; pa.b_lo = b_lo;
  pa.a_mn = a_mn; pa.b_mn = b_mn;
  pa.a_pitch = a_pitch; pa.b_pitch = b_pitch;
  pa.c = g->c; pa.bias = g->bias; pa.ws = ws;
  pa.m = g->m; pa.n = g->n; pa.k_pad = kp; pa.ldc = g->ldc;
  pa.k_per_split = ceil_div(ceil_div(kp, splits), kTK) * kTK;
  pa.alpha = g->alpha; pa.act = g->act; pa.accumulate = g->accumulate; pa.splits = splits;
  pa.cin_on = 0; pa.cin_t0 = pa.cin_xk = nullptr; pa.cin_ld0 = pa.cin_ldk = pa.cin_rows = 0; pa.cin_m = pa.cin_h = pa.cin_hp = 0;
  pa.fold_dt0 = pa.fold_dxk = nullptr; pa.fold_ldx = 0; pa.gen_groups = 0; pa.debug = tc_debug();
  cudaError_t e;
  const bool short_k = pa.k_per_split <= 4 * kTK;
  if (ws_kernel) {
    WsArgs wa;
    wa.p = pa;
    int ncta = g->m > kTM ? 2 : 1;
    if (b_mn && bn / ncta < 64) ncta = 1;        // an MN-major B half-tile must hold whole 64-wide atoms
    wa.tiles_m = (int)ceil_div(g->m, (int64_t)kTM * ncta);
    wa.tiles_n = (int)ceil_div(g->n, bn);
    wa.ntiles = (int64_t)wa.tiles_m * wa.tiles_n * splits;
    // TMA producers: four tiled tensor maps over the operand planes (box = one stage's slice of a plane)
    TmaMaps tm;
    const int bnh = bn / ncta;
    tm.ok = tma_map_2d(&tm.ah, a_hi, a_pitch, a_prows, a_pitch, 64, a_mn ? 64 : kTM) &&
            tma_map_2d(&tm.al, a_lo, a_pitch, a_prows, a_pitch, 64, a_mn ? 64 : kTM) &&
            tma_map_2d(&tm.bh, b_hi, b_pitch, b_prows, b_pitch, 64, b_mn ? 64 : bnh) &&
            tma_map_2d(&tm.bl, b_lo, b_pitch, b_prows, b_pitch, 64, b_mn ? 64 : bnh);
    if (ncta == 2) {
      if (bn == 32) e = launch_ws<32, 5, 2>(wa, tm, st);
      else if (bn == 64) e = launch_ws<64, 4, 2>(wa, tm, st);
      else if (bn == 128) e = launch_ws<128, 4, 2>(wa, tm, st);
      else e = launch_ws<256, 3, 2>(wa, tm, st);
    } else {
      if (bn == 32) e = launch_ws<32, 4, 1>(wa, tm, st);
      else if (bn == 64) e = launch_ws<64, 4, 1>(wa, tm, st);
      else if (bn == 128) e = launch_ws<128, 3, 1>(wa, tm, st);
      else e = launch_ws<256, 2, 1>(wa, tm, st);
    }
  } else if (bn == 32) e = short_k ? launch_planes<32, 1>(pa, st) : launch_planes<32, 4>(pa, st);
  else if (bn == 64) e = short_k ? launch_planes<64, 1>(pa, st) : launch_planes<64, 4>(pa, st);
  else if (bn == 128) e = short_k ? launch_planes<128, 1>(pa, st) : launch_planes<128, 3>(pa, st);
  else e = launch_planes<256, 2>(pa, st);
  if (e != cudaSuccess) {
    set_error("b2ctr_gemm(bf16x3 planes): CUDA launch failed: %s", cudaGetErrorString(e));
    return B2CTR_ERR_CUDA;
  }
  count_launch();
  if (splits > 1) {
    TcArgs ta;
    ta.c = g->c; ta.bias = g->bias; ta.ws = ws; ta.m = g->m; ta.n = g->n; ta.ldc = g->ldc;
    ta.act = g->act; ta.accumulate = g->accumulate; ta.splits = splits;
    tc_splitk_reduce_kernel<<<grid_for(g->m * g->n, 256, 4), 256, 0, st>>>(ta);
    B2_CHECK_LAUNCH("b2ctr_gemm(bf16x3 splitk_reduce)");
  }
  return B2CTR_OK;
}

// ================================================================================================
// CIN on the tensor cores without ever materialising the outer product.
// ================================================================================================
// planes of the PADDED filter W'[i*hp + j, n] = W[i*h + j, n] (j < h), zero rows for h <= j < hp
__global__ void __launch_bounds__(256)
    cin_filter_planes_kernel(const float* __restrict__ w, int m, int h, int hp, int64_t n, int64_t rows_pad,
                             int64_t cols_pad, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  const int64_t chunks = cols_pad / 8;
  const int64_t total = rows_pad * chunks;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / chunks;
    const int64_t c0 = (t - r * chunks) * 8;
    const int i = (int)(r / hp), j = (int)(r - (int64_t)i * hp);
    const bool ok = i < m && j < h;
    uint32_t hh[4], ll[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v0 = (ok && c0 + 2 * e < n) ? __ldg(w + ((int64_t)i * h + j) * n + c0 + 2 * e) : 0.f;
      const float v1 = (ok && c0 + 2 * e + 1 < n) ? __ldg(w + ((int64_t)i * h + j) * n + c0 + 2 * e + 1) : 0.f;
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
      hh[e] = pack_bf16(h0, h1);
      ll[e] = pack_bf16(__float2bfloat16_rn(v0 - __bfloat162float(h0)), __float2bfloat16_rn(v1 - __bfloat162float(h1)));
    }
    *reinterpret_cast<uint4*>(hi + r * cols_pad + c0) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    *reinterpret_cast<uint4*>(lo + r * cols_pad + c0) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
  }
}

b2ctr_status_t cin_filter_planes(const float* w, int m, int h, int hp, int64_t n, void* planes, cudaStream_t st) {
  const int64_t rp = round_up((int64_t)m * hp, 256), cp = planes_cols_pad(n);
  __nv_bfloat16* hi = (__nv_bfloat16*)planes;
  cin_filter_planes_kernel<<<grid_for(rp * (cp / 8), 256, 8), 256, 0, st>>>(w, m, h, hp, n, rp, cp, hi, hi + rp * cp);
  B2_CHECK_LAUNCH("b2ctr_cin_filter_planes");
  return B2CTR_OK;
}

// A GEMM whose A operand is generated by the producer warps (kind 1: CIN outer product, kind 2: DIN attention input)
struct GenSpec {
  int kind;
  const float* p0; int64_t ld0;
  const float* p1; int64_t ld1;
  int64_t rows;
  int m, h, hp;
  int64_t kq;            // columns of the generated matrix
};

static size_t gen_gemm_workspace_bytes(const GenSpec& sp, int mode, int64_t n, int split_k) {
  const int64_t M = mode == 0 ? sp.rows : sp.kq;
  return split_k > 1 ? (size_t)split_k * M * n * sizeof(float) + 256 : 0;
}

template <int GEN>
static cudaError_t launch_gen(const WsArgs& wa, const TmaMaps& tm, int bn, int ncta, cudaStream_t st) {
  if (ncta == 2) {
    if (bn == 128) return launch_ws_impl<128, 4, 2, true, GEN>(wa, tm, st);
    return launch_ws_impl<256, 3, 2, true, GEN>(wa, tm, st);
  }
  if (bn == 64) return launch_ws_impl<64, 4, 1, true, GEN>(wa, tm, st);
  if (bn == 128) return launch_ws_impl<128, 3, 1, true, GEN>(wa, tm, st);
  return launch_ws_impl<256, 2, 1, true, GEN>(wa, tm, st);
}

// mode 0: c[rows, n] = act(A B + bias), B = planes of a [kq, n] row-major matrix; mode 1: c[kq, n] = A^T dY,
// dY given as the planes of a [rows, n] row-major matrix.
static b2ctr_status_t gen_gemm(const GenSpec& sp, int mode, int64_t n, const void* planes, float* c, int64_t ldc,
                               const float* bias, int act, int split_k, void* workspace, size_t workspace_bytes,
                               cudaStream_t st, const char* what) {
  B2_REQUIRE(sp.rows < (1ll << 31) && sp.kq < (1ll << 31), "%s: rows and generated width must fit 31 bits", what);
  const size_t need = gen_gemm_workspace_bytes(sp, mode, n, split_k);
  if (need && (!workspace || workspace_bytes < need)) {
    set_error("%s: needs %zu workspace bytes, got %zu", what, need, workspace_bytes);
    return B2CTR_ERR_WORKSPACE;
  }
  const int splits = split_k > 1 ? split_k : 1;
  PlaneArgs pa;
  pa.a_hi = pa.a_lo = nullptr; pa.a_pitch = 0;
  pa.cin_on = sp.kind; pa.cin_t0 = sp.p0; pa.cin_xk = sp.p1; pa.cin_ld0 = sp.ld0; pa.cin_ldk = sp.ld1;
  pa.cin_rows = sp.rows; pa.cin_m = sp.m; pa.cin_h = sp.h; pa.cin_hp = sp.hp;
  pa.fold_dt0 = pa.fold_dxk = nullptr; pa.fold_ldx = 0; pa.debug = 0;
  {
    static int groups = -1;
    if (groups < 0) { const char* ev = getenv("B2CTR_GEN_GROUPS"); groups = ev ? atoi(ev) : 0; }
    // measured (profiles/README.md): the CIN generator is faster with all 256 threads on every stage (C3 8.02 vs
    // 8.32 ms), the (non-resident) attention generator with two groups alternating stages (C4 3.55 vs 3.59 ms)
    // measured (C4, E = 64): the register-resident attention generator (B2CTR_GEN_GROUPS=1) is slower than the
    // two-group one, 2.88 vs 2.61 ms per step
    pa.gen_groups = sp.kind == 1 ? 1 : (groups == 1 ? 1 : 2);
  }
  pa.b_mn = 1;      // both B operands are row-major matrices whose reduction dim is their row index
  pa.c = c; pa.bias = bias; pa.ws = (float*)workspace; pa.ldc = ldc;
  pa.alpha = 1.f; pa.act = act; pa.accumulate = 0; pa.splits = splits;
  pa.n = n;
  const int64_t cp = planes_cols_pad(n);
  int64_t b_prows;
  if (mode == 0) {
    pa.a_mn = 0; pa.m = sp.rows; pa.k_pad = round_up(sp.kq, kTK);
    b_prows = round_up(sp.kq, 256);
  } else {
    pa.a_mn = 1; pa.m = sp.kq; pa.k_pad = round_up(sp.rows, kTK);
    b_prows = round_up(sp.rows, 256);
  }
  const __nv_bfloat16* bp = (const __nv_bfloat16*)planes;
  pa.b_hi = bp; pa.b_lo = bp + b_prows * cp; pa.b_pitch = cp;
  pa.k_per_split = ceil_div(ceil_div(pa.k_pad, splits), kTK) * kTK;
  const int bn = n <= 64 ? 64 : (n <= 128 ? 128 : 256);
  int ncta = pa.m > kTM ? 2 : 1;
  if (bn / ncta < 64) ncta = 1;
  WsArgs wa;
  wa.p = pa;
  wa.tiles_m = (int)ceil_div(pa.m, (int64_t)kTM * ncta);
  wa.tiles_n = (int)ceil_div(n, bn);
  wa.ntiles = (int64_t)wa.tiles_m * wa.tiles_n * splits;
  TmaMaps tm;
  tm.ok = tma_map_2d(&tm.bh, pa.b_hi, cp, b_prows, cp, 64, 64) && tma_map_2d(&tm.bl, pa.b_lo, cp, b_prows, cp, 64, 64);
  if (!tm.ok) {
    set_error("%s: cuTensorMapEncodeTiled unavailable (the kernel loads its B operand by TMA)", what);
    return B2CTR_ERR_UNSUPPORTED;
  }
  tm.ah = tm.bh; tm.al = tm.bl;
  const cudaError_t e = sp.kind == 1 ? launch_gen<1>(wa, tm, bn, ncta, st) : launch_gen<2>(wa, tm, bn, ncta, st);
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return B2CTR_ERR_CUDA;
  }
  count_launch();
  if (splits > 1) {
    TcArgs ta;
    ta.c = c; ta.bias = bias; ta.ws = pa.ws; ta.m = pa.m; ta.n = n; ta.ldc = ldc;
    ta.act = act; ta.accumulate = 0; ta.splits = splits;
    tc_splitk_reduce_kernel<<<grid_for(pa.m * n, 256, 4), 256, 0, st>>>(ta);
    B2_CHECK_LAUNCH("generated-operand gemm (splitk_reduce)");
  }
  return B2CTR_OK;
}

static GenSpec cin_spec(const b2ctr_cin_gemm_t* g) {
  GenSpec sp;
  sp.kind = 1; sp.p0 = g->t0; sp.ld0 = g->ld0; sp.p1 = g->xk; sp.ld1 = g->ldk; sp.rows = g->rows;
  sp.m = g->m; sp.h = g->h; sp.hp = g->hp; sp.kq = (int64_t)g->m * g->hp;
  return sp;
}
size_t cin_gemm_workspace_bytes(const b2ctr_cin_gemm_t* g) {
  return gen_gemm_workspace_bytes(cin_spec(g), g->mode, g->n, g->split_k);
}
b2ctr_status_t cin_gemm(const b2ctr_cin_gemm_t* g, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  B2_REQUIRE(g && g->t0 && g->xk && g->c && g->rows > 0 && g->m > 0 && g->h > 0 && g->n > 0, "cin_gemm: bad arguments");
  B2_REQUIRE(g->hp >= g->h && (g->hp == 32 || g->hp % 64 == 0), "cin_gemm: hp must be 32 or a multiple of 64 and >= h");
  B2_REQUIRE(g->ldk % 4 == 0 && ((uintptr_t)g->xk & 15) == 0 && g->ld0 >= g->m, "cin_gemm: xk rows must be 16-byte aligned");
  B2_REQUIRE(g->h % 4 == 0 || g->ldk >= g->hp, "cin_gemm: h must be a multiple of 4 unless the xk rows are padded to hp");
  B2_REQUIRE(g->mode == 0 ? g->w_planes != nullptr : g->dy_planes != nullptr, "cin_gemm: operand planes missing");
  return gen_gemm(cin_spec(g), g->mode, g->n, g->mode == 0 ? g->w_planes : g->dy_planes, g->c, g->ldc, g->bias, g->act,
                  g->split_k, workspace, workspace_bytes, st, "b2ctr_cin_gemm");
}

// CIN backward, data gradient: dZ = dY W'^T folded onto T0 and X_k inside the epilogue (FOLD kernel).
// dY given as K-major planes of [rows, n]; W' planes (b2ctr_cin_filter_planes) read K-major (rows = q).
b2ctr_status_t cin_fold(const b2ctr_cin_gemm_t* g, float* dt0, float* dxk, int64_t ldx, cudaStream_t st) {
  B2_REQUIRE(g && g->t0 && g->xk && g->w_planes && g->dy_planes && dt0 && dxk, "cin_fold: bad arguments");
  B2_REQUIRE(g->hp == 32 || g->hp == 64 || g->hp == 128, "cin_fold: hp must be 32, 64 or 128");
  B2_REQUIRE(g->ldk % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)dxk & 15) == 0 && ((uintptr_t)g->xk & 15) == 0 &&
                 (g->h % 4 == 0 || (g->ldk >= g->hp && ldx >= g->hp)) && ldx >= g->h,
             "cin_fold: xk / dxk rows must be 16-byte aligned (and padded to hp when h is not a multiple of 4)");
  const int64_t kq = (int64_t)g->m * g->hp;
  PlaneArgs pa;
  const int64_t cpn = planes_cols_pad(g->n);
  const __nv_bfloat16* ap = (const __nv_bfloat16*)g->dy_planes;
  const int64_t a_prows = round_up(g->rows, 256);
  pa.a_hi = ap; pa.a_lo = ap + a_prows * cpn; pa.a_pitch = cpn; pa.a_mn = 0;
  const __nv_bfloat16* bp = (const __nv_bfloat16*)g->w_planes;
  const int64_t b_prows = round_up(kq, 256);
  pa.b_hi = bp; pa.b_lo = bp + b_prows * cpn; pa.b_pitch = cpn; pa.b_mn = 0;
  pa.c = nullptr; pa.bias = nullptr; pa.ws = nullptr; pa.ldc = 0;
  pa.m = g->rows; pa.n = kq; pa.k_pad = round_up(g->n, kTK);
  pa.k_per_split = pa.k_pad; pa.alpha = 1.f; pa.act = 0; pa.accumulate = 0; pa.splits = 1;
  pa.cin_on = 0; pa.cin_t0 = g->t0; pa.cin_xk = g->xk; pa.cin_ld0 = g->ld0; pa.cin_ldk = g->ldk; pa.cin_rows = g->rows;
  pa.cin_m = g->m; pa.cin_h = g->h; pa.cin_hp = g->hp;
  pa.fold_dt0 = dt0; pa.fold_dxk = dxk; pa.fold_ldx = ldx; pa.gen_groups = 0; pa.debug = 0;
  constexpr int bn = 128;
  const int ncta = g->rows > kTM ? 2 : 1;
  WsArgs wa;
  wa.p = pa;
  wa.tiles_m = (int)ceil_div(g->rows, (int64_t)kTM * ncta);
  wa.tiles_n = (int)ceil_div(kq, bn);
  wa.ntiles = (int64_t)wa.tiles_m * wa.tiles_n;
  TmaMaps tm;
  const int bnh = bn / ncta;
  tm.ok = tma_map_2d(&tm.ah, pa.a_hi, cpn, a_prows, cpn, 64, kTM) && tma_map_2d(&tm.al, pa.a_lo, cpn, a_prows, cpn, 64, kTM) &&
          tma_map_2d(&tm.bh, pa.b_hi, cpn, b_prows, cpn, 64, bnh) && tma_map_2d(&tm.bl, pa.b_lo, cpn, b_prows, cpn, 64, bnh);
  if (!tm.ok) {
    set_error("cin_fold: cuTensorMapEncodeTiled unavailable");
    return B2CTR_ERR_UNSUPPORTED;
  }
  cudaError_t e = ncta == 2 ? launch_ws_impl<128, 4, 2, true, 0, true>(wa, tm, st)
                            : launch_ws_impl<128, 3, 1, true, 0, true>(wa, tm, st);
  if (e != cudaSuccess) {
    set_error("b2ctr_cin_fold: CUDA launch failed: %s", cudaGetErrorString(e));
    return B2CTR_ERR_CUDA;
  }
  count_launch();
  return B2CTR_OK;
}

static GenSpec att_spec(const b2ctr_att_gemm_t* g) {
  GenSpec sp;
  sp.kind = 2; sp.p0 = g->query; sp.ld0 = g->ldq; sp.p1 = g->keys; sp.ld1 = g->key_batch_stride;
  sp.rows = g->batch * g->maxlen; sp.m = g->maxlen; sp.h = g->dim; sp.hp = g->dim; sp.kq = 4 * (int64_t)g->dim;
  return sp;
}
size_t att_gemm_workspace_bytes(const b2ctr_att_gemm_t* g) {
  return gen_gemm_workspace_bytes(att_spec(g), g->mode, g->n, g->split_k);
}
b2ctr_status_t att_gemm(const b2ctr_att_gemm_t* g, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  B2_REQUIRE(g && g->query && g->keys && g->c && g->planes && g->batch > 0 && g->maxlen > 0 && g->n > 0,
             "att_gemm: bad arguments");
  B2_REQUIRE(g->dim > 0 && g->dim % 8 == 0, "att_gemm: the embedding size must be a multiple of 8");
  B2_REQUIRE(g->ldq % 4 == 0 && g->key_batch_stride % 4 == 0 && ((uintptr_t)g->query & 15) == 0 &&
                 ((uintptr_t)g->keys & 15) == 0, "att_gemm: query / keys rows must be 16-byte aligned");
  return gen_gemm(att_spec(g), g->mode, g->n, g->planes, g->c, g->ldc, g->bias, g->act, g->split_k, workspace,
                  workspace_bytes, st, "b2ctr_att_gemm");
}

size_t planes_bytes(int64_t rows, int64_t cols) {
  return (size_t)round_up(rows, 256) * planes_cols_pad(cols) * 2 * sizeof(__nv_bfloat16) + 256;
}
b2ctr_status_t split_planes(const float* src, int64_t ld, int64_t rows, int64_t cols, void* planes,
                            cudaStream_t st) {
  const int64_t rp = round_up(rows, 256), cp = planes_cols_pad(cols);
  __nv_bfloat16* hi = (__nv_bfloat16*)planes;
  const int vec = (ld % 4 == 0) && (((uintptr_t)src & 15) == 0);
  split_planes_kernel<<<grid_for(rp * (cp / 8), 256, 8), 256, 0, st>>>(src, ld, rows, cols, rp, cp, hi, hi + rp * cp,
                                                                      vec);
  B2_CHECK_LAUNCH("b2ctr_split_planes");
  return B2CTR_OK;
}

b2ctr_status_t gemm_bf16x3(const b2ctr_gemm_t* g, void* workspace, size_t workspace_bytes,
                           cudaStream_t st) {
  if (tc_variant(g) >= 2) return gemm_planes(g, workspace, workspace_bytes, st);
  TcArgs ta;
  ta.a = g->a; ta.b = g->b; ta.c = g->c; ta.bias = g->bias; ta.ws = (float*)workspace;
  ta.m = g->m; ta.n = g->n; ta.k = g->k;
  ta.sam = g->trans_a ? 1 : g->lda;  ta.sak = g->trans_a ? g->lda : 1;
  ta.sbn = g->trans_b ? g->ldb : 1;  ta.sbk = g->trans_b ? 1 : g->ldb;
  ta.ldc = g->ldc; ta.alpha = g->alpha; ta.act = g->act; ta.accumulate = g->accumulate;
  ta.splits = g->split_k > 1 ? g->split_k : 1;
  if (ta.splits > 1) {
    const size_t need = gemm_bf16x3_workspace_bytes(g);
    if (!workspace || workspace_bytes < need) {
      set_error("gemm(bf16x3): split_k=%d needs %zu workspace bytes, got %zu", ta.splits, need, workspace_bytes);
      return B2CTR_ERR_WORKSPACE;
    }
  }
  ta.k_per_split = ceil_div(ceil_div(g->k, ta.splits), kTK) * kTK;
  const bool akc = !g->trans_a;   // A stored [M,K]: K contiguous
  const bool bkc = g->trans_b != 0;  // B stored [N,K]: K contiguous
  cudaError_t e;
  static int stages_mode = -1;
  if (stages_mode < 0) {
    const char* ev = getenv("B2CTR_TC_STAGES");
    stages_mode = ev ? atoi(ev) : 1;
  }
  if (stages_mode == 1) {        // single stage, 3 CTAs per SM: neighbours hide each other's load latency
    if (g->n <= 32) e = launch_tc<32, 1>(ta, akc, bkc, st);
    else if (g->n <= 64) e = launch_tc<64, 1>(ta, akc, bkc, st);
    else e = launch_tc<128, 1>(ta, akc, bkc, st);
  } else {                       // deep intra-CTA pipeline, 1 CTA per SM
    if (g->n <= 32) e = launch_tc<32, 4>(ta, akc, bkc, st);
    else if (g->n <= 64) e = launch_tc<64, 4>(ta, akc, bkc, st);
    else e = launch_tc<128, 3>(ta, akc, bkc, st);
  }
  if (e != cudaSuccess) {
    set_error("b2ctr_gemm(bf16x3): CUDA launch failed: %s", cudaGetErrorString(e));
    return B2CTR_ERR_CUDA;
  }
  count_launch();
  if (ta.splits > 1) {
    tc_splitk_reduce_kernel<<<grid_for(g->m * g->n, 256, 4), 256, 0, st>>>(ta);
    B2_CHECK_LAUNCH("b2ctr_gemm(bf16x3 splitk_reduce)");
  }
  return B2CTR_OK;
}

}  // namespace b2ctr

extern "C" {
size_t b2ctr_cin_filter_planes_bytes(int32_t m, int32_t hp, int64_t n) { return b2ctr::planes_bytes((int64_t)m * hp, n); }
b2ctr_status_t b2ctr_cin_filter_planes(const float* w, int32_t m, int32_t h, int32_t hp, int64_t n, void* planes,
                                       void* stream) {
  B2_REQUIRE(w && planes && m > 0 && h > 0 && hp >= h && n > 0, "cin_filter_planes: bad arguments");
  return b2ctr::cin_filter_planes(w, m, h, hp, n, planes, (cudaStream_t)stream);
}
size_t b2ctr_cin_gemm_workspace_bytes(const b2ctr_cin_gemm_t* g) { return g ? b2ctr::cin_gemm_workspace_bytes(g) : 0; }
b2ctr_status_t b2ctr_cin_gemm(const b2ctr_cin_gemm_t* g, void* workspace, size_t workspace_bytes, void* stream) {
  return b2ctr::cin_gemm(g, workspace, workspace_bytes, (cudaStream_t)stream);
}
b2ctr_status_t b2ctr_cin_fold(const b2ctr_cin_gemm_t* g, float* dt0, float* dxk, int64_t ldx, void* stream) {
  return b2ctr::cin_fold(g, dt0, dxk, ldx, (cudaStream_t)stream);
}
size_t b2ctr_att_gemm_workspace_bytes(const b2ctr_att_gemm_t* g) { return g ? b2ctr::att_gemm_workspace_bytes(g) : 0; }
b2ctr_status_t b2ctr_att_gemm(const b2ctr_att_gemm_t* g, void* workspace, size_t workspace_bytes, void* stream) {
  return b2ctr::att_gemm(g, workspace, workspace_bytes, (cudaStream_t)stream);
}
}
