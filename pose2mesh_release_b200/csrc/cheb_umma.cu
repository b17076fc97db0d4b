// tcgen05 (5th-gen tensor core) Chebyshev graph convolution for sm_100a — ONE kernel per layer:
//
//   Y[tile] = epilogue( [T0 | T1 | T2](X)[tile] * W^T ),   T1 = L~ X, T2 = 2 L~ T1 - X
//
// A CTA owns a tile of 128 consecutive vertices of one mesh (a compact patch: the reference's
// binary-tree vertex order makes rows [128p,128p+128) the descendants of one coarse node).
//   * 8 producer warps build the A operand on chip, 32 features at a time: they stage the tile's
//     2-hop halo of X in shared memory (cp.async, 128 B per row = one cache line), run the two
//     sparse products out of shared memory with a tile-local CSR, split every fp32 value into an
//     fp16 (hi, lo) pair and write it straight into the 128B-swizzled K-major UMMA layout.
//     T0/T1/T2 are never materialised in HBM.
//   * 1 thread streams the pre-packed fp16 (hi|lo) weight blocks with cp.async.bulk (TMA engine,
//     mbarrier complete_tx).
//   * 1 thread issues tcgen05.mma (kind::f16, M=128, N=Fout, K=16) into a TMEM accumulator:
//     per 16 features three MMAs — hi*Whi + lo*Whi + hi*Wlo — i.e. an error-compensated product
//     with ~2^-21 relative error, which is what keeps the 1e-4 fp32 parity bar (plain TF32/FP16
//     does not, SURVEY.md §7 "hard parts" 1).
//   * the 8 producer warps then drain TMEM (tcgen05.ld) through the fused epilogue: bias /
//     folded BatchNorm, ReLU, channel-resampled residual, and store.
// The unpool between levels is virtual: with in_unpool the halo rows are read from row r>>1 of the
// coarser tensor.  Accumulation is fp32 in TMEM; weights are pre-scaled by 2^6 so that their lo
// parts stay normal fp16 numbers (undone exactly in the epilogue).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "p2m_internal.h"

namespace p2m {

namespace {

constexpr int TILE_M = 128;
constexpr int FC = 32;                         // features per chunk (= 128 B of fp32 per row)
constexpr int A_BLOCK_BYTES = TILE_M * 128;    // one K-block of A: 128 rows x (32 hi | 32 lo) fp16
constexpr int NUM_WORKERS = 256;               // 8 producer / epilogue warps
constexpr int NUM_THREADS = NUM_WORKERS + 64;  // + warp 8 (weight loader) + warp 9 (MMA issuer, TMEM owner)
constexpr float W_SCALE = 64.f;
constexpr float W_INV_SCALE = 1.f / 64.f;

// ------------------------------------------------------------------ per-tile metadata blob
struct TileHeader {  // 64 bytes
  int n_rows;     // valid vertices in this tile (<= 128)
  int h1;         // rows of T1 kept on chip: 128 tile slots + 1-hop halo
  int h2;         // rows of X staged: h1 + 2-hop halo
  int nnz_a, nnz_b;
  int off_halo;   // int32 [h2]   vertex id of staged row i (-1: empty slot)
  int off_rpa;    // uint16 [h1+1] CSR over T1 rows, columns index staged X rows
  int off_idxa;   // uint16 [nnz_a]
  int off_vala;   // float  [nnz_a]
  int off_rpb;    // uint16 [129]  CSR over the 128 tile rows, columns index T1 rows
  int off_idxb;   // uint16 [nnz_b]
  int off_valb;   // float  [nnz_b]
  int bytes;
  int pad[3];
};
static_assert(sizeof(TileHeader) == 64, "header size");

inline int up16(int x) { return (x + 15) & ~15; }

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU box.  On timeout the CTA-wide abort flag is
// raised, the global status word is set and every later wait falls through immediately.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag, int* status,
                                          int code) {
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    if (mbar_try_wait(bar, parity)) return;
    if (*abort_flag) return;
  }
  *abort_flag = 1;
  atomicExch(status, code);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; kind::f16, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand block (rows of 128 bytes, 8-row atoms of 1024 bytes).
// Field layout: cute::UMMA::SmemDescriptor (start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout_type [61,64) with SWIZZLE_128B = 2).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // LBO (unused for swizzled K-major) = 16 B
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO = 1024 B between 8-row groups
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) at [4,6), a/b_format F16 (0), K-major A and B,
// n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a 128B-swizzled block
__host__ __device__ __forceinline__ uint32_t sw128_off(int row, int chunk) {
  return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
}

struct Half4 {
  __half2 a, b;
};
__device__ __forceinline__ void split4(const float4& v, uint2& hi, uint2& lo) {
  __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
  hi.x = *reinterpret_cast<uint32_t*>(&h0);
  hi.y = *reinterpret_cast<uint32_t*>(&h1);
  lo.x = *reinterpret_cast<uint32_t*>(&l0);
  lo.y = *reinterpret_cast<uint32_t*>(&l1);
}

struct KParams {
  const float* x;
  int in_unpool;
  int V, P, fin;
  const unsigned char* meta;
  const int* meta_bytes;
  int meta_stride, max_h1, max_h2;
  const unsigned char* wpack;
  EpiDev ep;
  float* y;
  int* status;
};

// =====================================================================================
template <int N, int NS>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_cheb_conv_umma(const KParams p) {
  constexpr int B_BLOCK_BYTES = N * 128;
  constexpr int SLOT_BYTES = A_BLOCK_BYTES + B_BLOCK_BYTES;
  constexpr uint32_t IDESC = make_idesc_f16(TILE_M, N);
  constexpr int TMEM_COLS = N < 32 ? 32 : N;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* ring = smem;
  float* Xs = reinterpret_cast<float*>(ring + NS * SLOT_BYTES);
  float* T1s = Xs + (size_t)p.max_h2 * FC;
  unsigned char* meta_s = reinterpret_cast<unsigned char*>(T1s + (size_t)p.max_h1 * FC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta_s + p.meta_stride);
  // bars: full[NS], empty[NS], accum_full, meta_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 2);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int tile = blockIdx.x;
  const int b = tile / p.P;
  const int pat = tile - b * p.P;
  const long long mesh_row0 = (long long)b * p.V;
  const int n_chunk = p.fin / FC;
  const int n_use = 3 * n_chunk;

  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = smem_u32(bars + NS);
  const uint32_t bar_accum = smem_u32(bars + 2 * NS);
  const uint32_t bar_meta = smem_u32(bars + 2 * NS + 1);

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bar_full + 8 * s, NUM_WORKERS + 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_accum, 1);
    mbar_init(bar_meta, 1);
    *abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------ weight-block loader (one thread)
    if ((tid & 31) == 0) {
      const int mbytes = p.meta_bytes[pat];
      mbar_arrive_expect_tx(bar_meta, mbytes);
      bulk_g2s(smem_u32(meta_s), p.meta + (size_t)pat * p.meta_stride, mbytes, bar_meta);
      for (int u = 0; u < n_use; ++u) {
        const int s = u % NS, round = u / NS;
        mbar_wait(bar_empty + 8 * s, (round & 1) ^ 1, abort_flag, p.status, 1);
        mbar_arrive_expect_tx(bar_full + 8 * s, B_BLOCK_BYTES);
        bulk_g2s(smem_u32(ring + s * SLOT_BYTES + A_BLOCK_BYTES), p.wpack + (size_t)u * B_BLOCK_BYTES, B_BLOCK_BYTES,
                 bar_full + 8 * s);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    if ((tid & 31) == 0) {
      for (int u = 0; u < n_use; ++u) {
        const int s = u % NS, round = u / NS;
        mbar_wait(bar_full + 8 * s, round & 1, abort_flag, p.status, 2);
        tc_fence_after();
        const uint32_t a0 = smem_u32(ring + s * SLOT_BYTES);
        const uint32_t b0 = a0 + A_BLOCK_BYTES;
        const uint64_t da = make_desc_sw128(a0), db = make_desc_sw128(b0);
        // A block columns: [hi 0..31 | lo 32..63], B block columns: [Whi 0..31 | Wlo 32..63] (fp16);
        // a 16-element K step is 32 bytes = +2 in the descriptor's start-address field.
        //   hi*Whi            lo*Whi            hi*Wlo
        umma_f16(tmem_base, da + 0, db + 0, IDESC, u > 0);
        umma_f16(tmem_base, da + 2, db + 2, IDESC, 1);
        umma_f16(tmem_base, da + 4, db + 0, IDESC, 1);
        umma_f16(tmem_base, da + 6, db + 2, IDESC, 1);
        umma_f16(tmem_base, da + 0, db + 4, IDESC, 1);
        umma_f16(tmem_base, da + 2, db + 6, IDESC, 1);
        umma_commit(bar_empty + 8 * s);  // frees the slot when these MMAs have read it
      }
      umma_commit(bar_accum);
    }
  } else {
    // ------------------------------------------------------------ producers (8 warps)
    mbar_wait(bar_meta, 0, abort_flag, p.status, 3);
    const TileHeader* hdr = reinterpret_cast<const TileHeader*>(meta_s);
    const int h1 = hdr->h1, h2 = hdr->h2;
    const int* halo = reinterpret_cast<const int*>(meta_s + hdr->off_halo);
    const unsigned short* rpa = reinterpret_cast<const unsigned short*>(meta_s + hdr->off_rpa);
    const unsigned short* idxa = reinterpret_cast<const unsigned short*>(meta_s + hdr->off_idxa);
    const float* vala = reinterpret_cast<const float*>(meta_s + hdr->off_vala);
    const unsigned short* rpb = reinterpret_cast<const unsigned short*>(meta_s + hdr->off_rpb);
    const unsigned short* idxb = reinterpret_cast<const unsigned short*>(meta_s + hdr->off_idxb);
    const float* valb = reinterpret_cast<const float*>(meta_s + hdr->off_valb);
    const int q = tid & 7;     // float4 lane inside the 32-feature chunk
    const int rg = tid >> 3;   // row group 0..31
    const float4* Xs4 = reinterpret_cast<const float4*>(Xs);
    float4* T1s4 = reinterpret_cast<float4*>(T1s);

    for (int c = 0; c < n_chunk; ++c) {
      // (1) stage the 2-hop halo of X for this feature chunk
      for (int i = rg; i < h2; i += 32) {
        const int v = halo[i];
        const uint32_t dst = smem_u32(Xs + (size_t)i * FC + q * 4);
        if (v >= 0) {
          long long r = mesh_row0 + v;
          if (p.in_unpool) r >>= 1;
          cp_async16(dst, p.x + r * p.fin + c * FC + q * 4);
        } else {
          *reinterpret_cast<float4*>(Xs + (size_t)i * FC + q * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      cp_async_wait_all();
      worker_barrier();
      // (2) T1 = L~ X on the tile rows and their 1-hop halo
      for (int i = rg; i < h1; i += 32) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int e1 = rpa[i + 1];
        for (int e = rpa[i]; e < e1; ++e) {
          const float w = vala[e];
          const float4 xv = Xs4[(size_t)idxa[e] * 8 + q];
          acc.x = fmaf(w, xv.x, acc.x);
          acc.y = fmaf(w, xv.y, acc.y);
          acc.z = fmaf(w, xv.z, acc.z);
          acc.w = fmaf(w, xv.w, acc.w);
        }
        T1s4[(size_t)i * 8 + q] = acc;
      }
      worker_barrier();
      // (3) T2 = 2 L~ T1 - X on the tile rows; split to fp16 hi/lo; write the three K-blocks
      float4 t0[4], t1[4], t2[4];
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const int i = ps * 32 + rg;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int e1 = rpb[i + 1];
        for (int e = rpb[i]; e < e1; ++e) {
          const float w = valb[e];
          const float4 tv = T1s4[(size_t)idxb[e] * 8 + q];
          acc.x = fmaf(w, tv.x, acc.x);
          acc.y = fmaf(w, tv.y, acc.y);
          acc.z = fmaf(w, tv.z, acc.z);
          acc.w = fmaf(w, tv.w, acc.w);
        }
        t0[ps] = Xs4[(size_t)i * 8 + q];
        t1[ps] = T1s4[(size_t)i * 8 + q];
        t2[ps] = make_float4(2.f * acc.x - t0[ps].x, 2.f * acc.y - t0[ps].y, 2.f * acc.z - t0[ps].z,
                             2.f * acc.w - t0[ps].w);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int u = 3 * c + k;
        const int s = u % NS, round = u / NS;
        mbar_wait(bar_empty + 8 * s, (round & 1) ^ 1, abort_flag, p.status, 4);
        unsigned char* ablk = ring + s * SLOT_BYTES;
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          const int i = ps * 32 + rg;
          const float4 v = (k == 0) ? t0[ps] : (k == 1 ? t1[ps] : t2[ps]);
          uint2 hi, lo;
          split4(v, hi, lo);
          *reinterpret_cast<uint2*>(ablk + sw128_off(i, q >> 1) + (q & 1) * 8) = hi;
          *reinterpret_cast<uint2*>(ablk + sw128_off(i, 4 + (q >> 1)) + (q & 1) * 8) = lo;
        }
        fence_async_proxy();  // make the generic-proxy stores visible to the tensor core (async proxy)
        mbar_arrive(bar_full + 8 * s);
      }
      worker_barrier();  // Xs / T1s are overwritten by the next chunk
    }

    // ------------------------------------------------------------ epilogue: TMEM -> registers -> HBM
    mbar_wait(bar_accum, 0, abort_flag, p.status, 5);
    tc_fence_after();
    const int lane_base = (warp & 3) * 32;
    const int row_in_tile = lane_base + (tid & 31);
    constexpr int COLS_PER_WARP = N / 2;
    const int col0 = (warp >> 2) * COLS_PER_WARP;
    const long long r = mesh_row0 + (long long)pat * TILE_M + row_in_tile;
    const bool valid = row_in_tile < hdr->n_rows;
    float* yrow = p.y + r * N;
#pragma unroll 1
    for (int cb = 0; cb < COLS_PER_WARP; cb += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(col0 + cb), v);
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o;
          const int n = col0 + cb + j;
          o.x = apply_epilogue(__uint_as_float(v[j + 0]) * W_INV_SCALE, r, n + 0, p.ep);
          o.y = apply_epilogue(__uint_as_float(v[j + 1]) * W_INV_SCALE, r, n + 1, p.ep);
          o.z = apply_epilogue(__uint_as_float(v[j + 2]) * W_INV_SCALE, r, n + 2, p.ep);
          o.w = apply_epilogue(__uint_as_float(v[j + 3]) * W_INV_SCALE, r, n + 3, p.ep);
          *reinterpret_cast<float4*>(yrow + n) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, TMEM_COLS);
}

// fp32 reference-layout weights [Fout, Fin*3] (column = f*3+k) -> K-blocks of fp16 [Whi | Wlo]
// in the exact shared-memory image (128B-swizzled), block u = chunk*3 + k, so the kernel can
// fetch a block with a single cp.async.bulk.
__global__ void __launch_bounds__(256) k_pack_weights(const float* __restrict__ W, int fin, int fout,
                                                      unsigned char* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk each
  const int n_blocks = (fin / FC) * 3;
  const int total = n_blocks * fout * 8;
  if (idx >= total) return;
  const int j = idx & 7;
  const int n = (idx >> 3) % fout;
  const int u = (idx >> 3) / fout;
  const int c = u / 3, k = u % 3;
  const int f0 = c * FC + (j & 3) * 8;
  __align__(16) __half h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float w = W[(size_t)n * fin * 3 + (size_t)(f0 + e) * 3 + k] * W_SCALE;
    const __half hi = __float2half_rn(w);
    h[e] = (j < 4) ? hi : __float2half_rn(w - __half2float(hi));
  }
  *reinterpret_cast<uint4*>(out + (size_t)u * fout * 128 + sw128_off(n, j)) = *reinterpret_cast<const uint4*>(h);
}

template <int N>
struct RingCfg {
  static constexpr int NS = (N == 256) ? 2 : 3;
};

size_t smem_bytes_for(int N, int NS, const DevLevel& g) {
  return 1024 + (size_t)NS * (A_BLOCK_BYTES + N * 128) + (size_t)g.max_h2 * FC * 4 + (size_t)g.max_h1 * FC * 4 +
         (size_t)g.meta_stride + 8 * (2 * NS + 2) + 16;
}

template <int N>
int launch_n(const UmmaConvArgs& a, int* status, cudaStream_t s) {
  constexpr int NS = RingCfg<N>::NS;
  const DevLevel& g = *a.g;
  const size_t smem = smem_bytes_for(N, NS, g);
  auto kern = k_cheb_conv_umma<N, NS>;
  P2M_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KParams p;
  p.x = a.x;
  p.in_unpool = a.in_unpool;
  p.V = g.V;
  p.P = g.n_pattern;
  p.fin = a.fin;
  p.meta = g.tile_meta;
  p.meta_bytes = g.tile_meta_bytes;
  p.meta_stride = g.meta_stride;
  p.max_h1 = g.max_h1;
  p.max_h2 = g.max_h2;
  p.wpack = static_cast<const unsigned char*>(a.wpack);
  p.ep = to_dev(a.ep);
  p.y = a.y;
  p.status = status;
  kern<<<a.batch * g.n_pattern, NUM_THREADS, smem, s>>>(p);
  P2M_LAUNCH_OK();
  return P2M_OK;
}

}  // namespace

// =====================================================================================
// host side
// =====================================================================================
int build_umma_level_meta(const int* rowptr, const int* colidx, const float* val, int V, DevLevel* out,
                          std::vector<void*>* owned) {
  const int P = (V + TILE_M - 1) / TILE_M;
  std::vector<std::vector<unsigned char>> blobs(P);
  int max_h1 = 0, max_h2 = 0, stride = 0;
  std::vector<int> slot_of(V, -1);
  for (int pt = 0; pt < P; ++pt) {
    const int v0 = pt * TILE_M;
    const int n_rows = std::min(TILE_M, V - v0);
    // staged-row list: slots 0..127 = tile rows, then 1-hop halo, then 2-hop halo
    std::vector<int> halo(TILE_M, -1);
    for (int i = 0; i < n_rows; ++i) {
      halo[i] = v0 + i;
      slot_of[v0 + i] = i;
    }
    auto add = [&](int v) {
      if (slot_of[v] < 0) {
        slot_of[v] = (int)halo.size();
        halo.push_back(v);
      }
    };
    for (int i = 0; i < n_rows; ++i)
      for (int e = rowptr[v0 + i]; e < rowptr[v0 + i + 1]; ++e) add(colidx[e]);
    const int h1 = (int)halo.size();
    for (int i = 0; i < h1; ++i) {
      if (halo[i] < 0) continue;
      for (int e = rowptr[halo[i]]; e < rowptr[halo[i] + 1]; ++e) add(colidx[e]);
    }
    const int h2 = (int)halo.size();
    if (h2 > 65535) {
      set_error("umma meta: halo too large");
      return P2M_ERR_INVALID;
    }
    std::vector<unsigned short> rpa(h1 + 1, 0), idxa, rpb(TILE_M + 1, 0), idxb;
    std::vector<float> vala, valb;
    for (int i = 0; i < h1; ++i) {
      if (halo[i] >= 0)
        for (int e = rowptr[halo[i]]; e < rowptr[halo[i] + 1]; ++e) {
          idxa.push_back((unsigned short)slot_of[colidx[e]]);
          vala.push_back(val[e]);
        }
      rpa[i + 1] = (unsigned short)idxa.size();
    }
    for (int i = 0; i < TILE_M; ++i) {
      if (i < n_rows)
        for (int e = rowptr[v0 + i]; e < rowptr[v0 + i + 1]; ++e) {
          idxb.push_back((unsigned short)slot_of[colidx[e]]);  // < h1 by construction
          valb.push_back(val[e]);
        }
      rpb[i + 1] = (unsigned short)idxb.size();
    }
    if (idxa.size() > 65535) {
      set_error("umma meta: too many entries in a tile");
      return P2M_ERR_INVALID;
    }
    TileHeader h{};
    h.n_rows = n_rows;
    h.h1 = h1;
    h.h2 = h2;
    h.nnz_a = (int)idxa.size();
    h.nnz_b = (int)idxb.size();
    int off = 64;
    h.off_halo = off; off += up16(h2 * 4);
    h.off_rpa = off;  off += up16((h1 + 1) * 2);
    h.off_idxa = off; off += up16(h.nnz_a * 2);
    h.off_vala = off; off += up16(h.nnz_a * 4);
    h.off_rpb = off;  off += up16((TILE_M + 1) * 2);
    h.off_idxb = off; off += up16(h.nnz_b * 2);
    h.off_valb = off; off += up16(h.nnz_b * 4);
    h.bytes = off;
    std::vector<unsigned char>& blob = blobs[pt];
    blob.assign(off, 0);
    std::memcpy(blob.data(), &h, sizeof(h));
    std::memcpy(blob.data() + h.off_halo, halo.data(), h2 * 4);
    std::memcpy(blob.data() + h.off_rpa, rpa.data(), (h1 + 1) * 2);
    if (h.nnz_a) std::memcpy(blob.data() + h.off_idxa, idxa.data(), h.nnz_a * 2);
    if (h.nnz_a) std::memcpy(blob.data() + h.off_vala, vala.data(), h.nnz_a * 4);
    std::memcpy(blob.data() + h.off_rpb, rpb.data(), (TILE_M + 1) * 2);
    if (h.nnz_b) std::memcpy(blob.data() + h.off_idxb, idxb.data(), h.nnz_b * 2);
    if (h.nnz_b) std::memcpy(blob.data() + h.off_valb, valb.data(), h.nnz_b * 4);
    max_h1 = std::max(max_h1, h1);
    max_h2 = std::max(max_h2, h2);
    stride = std::max(stride, off);
    for (int v : halo)
      if (v >= 0) slot_of[v] = -1;
  }
  stride = (stride + 127) & ~127;
  std::vector<unsigned char> all((size_t)P * stride, 0);
  std::vector<int> bytes(P);
  for (int pt = 0; pt < P; ++pt) {
    std::memcpy(all.data() + (size_t)pt * stride, blobs[pt].data(), blobs[pt].size());
    bytes[pt] = (int)blobs[pt].size();
  }
  unsigned char* d_meta = nullptr;
  int* d_bytes = nullptr;
  P2M_CUDA_OK(cudaMalloc(&d_meta, all.size()));
  owned->push_back(d_meta);
  P2M_CUDA_OK(cudaMalloc(&d_bytes, sizeof(int) * P));
  owned->push_back(d_bytes);
  P2M_CUDA_OK(cudaMemcpy(d_meta, all.data(), all.size(), cudaMemcpyHostToDevice));
  P2M_CUDA_OK(cudaMemcpy(d_bytes, bytes.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
  out->n_pattern = P;
  out->tile_meta = d_meta;
  out->tile_meta_bytes = d_bytes;
  out->meta_stride = stride;
  out->max_h1 = max_h1;
  out->max_h2 = max_h2;
  return P2M_OK;
}

bool umma_conv_supported(const DevLevel& g, int fin, int fout) {
  if (g.tile_meta == nullptr || g.n_pattern <= 0) return false;
  if (fin % FC != 0 || fin < FC || fin > 256) return false;
  if (fout != 64 && fout != 128 && fout != 256) return false;
  const int ns = (fout == 256) ? 2 : 3;
  return smem_bytes_for(fout, ns, g) <= 227 * 1024;
}

size_t umma_wpack_bytes(int fin, int fout) { return (size_t)(fin / FC) * 3 * fout * 128; }

int launch_umma_pack_weights(const float* W, int fin, int fout, void* wpack, cudaStream_t s) {
  const int total = (fin / FC) * 3 * fout * 8;
  k_pack_weights<<<(total + 255) / 256, 256, 0, s>>>(W, fin, fout, static_cast<unsigned char*>(wpack));
  P2M_LAUNCH_OK();
  return P2M_OK;
}

int launch_umma_conv(const UmmaConvArgs& a, int* status, cudaStream_t s) {
  if (!umma_conv_supported(*a.g, a.fin, a.fout)) {
    set_error("umma_conv: unsupported shape");
    return P2M_ERR_INVALID;
  }
  switch (a.fout) {
    case 64: return launch_n<64>(a, status, s);
    case 128: return launch_n<128>(a, status, s);
    case 256: return launch_n<256>(a, status, s);
  }
  return P2M_ERR_INVALID;
}

}  // namespace p2m
