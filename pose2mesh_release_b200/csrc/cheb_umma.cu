// tcgen05 (5th-gen tensor core) kernels of the MeshNet hot path for sm_100a.
//
// A Chebyshev graph-conv layer  Y = epilogue( [T0 | T1 | T2](X) * W^T ),  T1 = L~ X, T2 = 2 L~ T1 - X  is two launches:
//
// k_cheb_t1        — T1 = L~ X for every row, written once to HBM (fp32): a 4-deep cp.async ring per 128-row tile,
//                    gathers out of shared memory.
// k_cheb_conv_umma — one persistent CTA per SM; a tile is 128 consecutive vertices of one mesh (a compact patch: the
//                    reference's binary-tree vertex order makes rows [128p,128p+128) the descendants of one coarse node).
//   * 16 producer warps build the A operand on chip, 32 features at a time: they run the second sparse product from
//     shared memory with a tile-local CSR (the T1 rows of the tile and its 1-hop halo, staged one chunk ahead), read
//     their own X rows straight from global memory, split every fp32 value into an fp16 (hi, lo) pair and write it
//     into the 128B-swizzled K-major UMMA layout.  T2 is never materialised in HBM.  (Without a T1 buffer —
//     p.t1 == nullptr, the split_t1 = 0 ablation — the same kernel stages the 2-hop halo of X and runs both sparse
//     products on chip: the fully fused variant.)
//   * 2 loader warps stage what the producers gather from: the tile's own T1 rows as one 2-D TMA box where the tile
//     is a run of consecutive rows, every other row (halo, index-list tiles) by 16-byte cp.async with completion
//     through cp.async.mbarrier.arrive.noinc; thread 0 also prefetches the next tile's metadata blob.  1 thread
//     streams the pre-packed fp16 (hi|lo) weight blocks (cp.async.bulk, mbarrier complete_tx).
//   * 1 thread issues tcgen05.mma (kind::f16, M=128, N=Fout, K=16) into a double-buffered TMEM accumulator:
//     per 16 features three MMAs — hi*Whi + lo*Whi + hi*Wlo — an error-compensated product with ~2^-21
//     relative error, which is what keeps the 1e-4 fp32 parity bar (plain TF32/FP16 does not, SURVEY.md §7
//     "hard parts" 1).
//   * 4 epilogue warps drain TMEM (tcgen05.ld), transpose the accumulator rows through a swizzled per-warp staging
//     buffer so that global accesses are coalesced, and apply the fused epilogue — bias / folded BatchNorm, ReLU,
//     channel-resampled residual (or, for the network's last block, the 64 -> 3 head's projection) — while the
//     next tile's main loop runs.
// The unpool between levels is virtual: with in_unpool the rows are read from row r>>1 of the coarser
// tensor.  Weights are pre-scaled by 2^6 so that their lo parts stay normal fp16 numbers (undone exactly in the
// epilogue).  The same kernel in `plain` mode is the backward dT GEMM; k_cheb_dw_umma (below) is the dW
// reduction with MN-major operands.  Every mbarrier wait is time-bounded (a protocol bug sets a status word
// instead of hanging the GPU) and tools/umma_trace.py dumps a per-role event timeline of CTA 0.
#include <cuda.h>  // CUtensorMap types only: the encoder is fetched through cudaGetDriverEntryPoint
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "p2m_internal.h"

namespace p2m {

namespace {

constexpr int TILE_M = 128;
constexpr int FC = 32;                         // features per chunk (= 128 B of fp32 per row)
constexpr int A_BLOCK_BYTES = TILE_M * 128;    // one K-block of A: 128 rows x (32 hi | 32 lo) fp16
constexpr float W_SCALE = 64.f;
constexpr float W_INV_SCALE = 1.f / 64.f;

// ------------------------------------------------------------------ per-tile metadata blob
struct TileHeader {  // 64 bytes
  int n_rows;     // valid vertices in this tile (<= 128)
  int h1;         // rows of T1 kept on chip: 128 tile slots + 1-hop halo
  int h2;         // rows of X staged: h1 + 2-hop halo
  int nnz;
  int off_halo;   // int32 [h2]    vertex id of staged row i (-1: empty slot)
  int off_rp;     // uint16 [h1+1] local CSR: row i < h1 lists the neighbours of vertex halo[i]
  int off_ent;    // uint2 [nnz]   {byte offset of the neighbour's staged row (slot*128), value bits}
  int off_ord1;   // uint16 [h1]   T1 rows sorted by decreasing length (warps see equal trip counts)
  int off_ord2;   // uint16 [128]  tile rows sorted by decreasing length
  int bytes;
  int pad[6];
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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(20000u)  // suspend-time hint (ns): sleep in hardware instead of spinning,
      : "memory");                          // so waiting warps do not take issue slots from the loader warps
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU box.  On timeout (2 s of %globaltimer: three orders of magnitude
// beyond any legitimate wait, preempted / time-sliced contexts included) the CTA-wide abort flag is raised so that
// every later wait of this CTA falls through, and the wait's id is written to the model's status word — a word of
// MAPPED HOST memory, so the host sees it without a copy: every API entry point checks it and fails with
// P2M_ERR_CUDA (p2m_api.cu: check_kernel_status), i.e. a timed-out kernel never hands results to the caller silently.
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long MBAR_TIMEOUT_NS = 2000000000ull;
__device__ __noinline__ void mbar_timeout(volatile int* abort_flag, int* status, int code) {
  *abort_flag = 1;
  *reinterpret_cast<volatile int*>(status) = code;
  __threadfence_system();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag, int* status,
                                          int code) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0 = 0;
  for (unsigned spin = 0;; ++spin) {
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
      if (mbar_try_wait(bar, parity)) return;
    }
    if (*abort_flag) return;
    if ((spin & 63u) == 63u) {  // the timer is only consulted every 4096 failed polls (~80 ms of hardware-suspended waits)
      const unsigned long long t = global_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > MBAR_TIMEOUT_NS) break;
    }
  }
  mbar_timeout(abort_flag, status, code);
}
// Same, but with a nanosleep back-off between polls: for the roles whose waits span most of a tile
// (epilogue, loaders) so that their polling does not steal issue slots from the producers.
template <unsigned SLEEP_NS = 200>
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, volatile int* abort_flag, int* status,
                                                  int code) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0 = 0;
  for (unsigned spin = 0;; ++spin) {
#pragma unroll 1
    for (int it = 0; it < 16; ++it) {
      if (mbar_try_wait(bar, parity)) return;
      __nanosleep(SLEEP_NS);
    }
    if (*abort_flag) return;
    if ((spin & 255u) == 255u) {
      const unsigned long long t = global_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > MBAR_TIMEOUT_NS) break;
    }
  }
  mbar_timeout(abort_flag, status, code);
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
// src_bytes = 16: plain copy; 0: the 16 destination bytes are zero-filled and the source is not read
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// One 2-D tiled TMA load (box = 32 floats x box rows of the tensor map) into dense 128-byte rows.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  // the mbarrier gets this thread's arrival once all of its earlier cp.async copies have landed
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void producer_barrier() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void epilogue_barrier() { asm volatile("bar.sync 2, 128;" ::: "memory"); }
// explicit shared-window accesses (32-bit addresses): keeps the hot loops on LDS/STS instead of generic LD/ST
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f4(uint32_t a, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_u2(uint32_t a, const uint2& v) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& x) {
  acc.x = fmaf(w, x.x, acc.x);
  acc.y = fmaf(w, x.y, acc.y);
  acc.z = fmaf(w, x.z, acc.z);
  acc.w = fmaf(w, x.w, acc.w);
}
// One CSR row with four entries in flight: the four (slot, value) pairs are fetched first, then the four
// 16-byte row pieces, then the FMAs on two accumulators — the row's dependent chain is ~2 shared-memory round
// trips per four entries instead of two per entry (rows have <= 14 entries).  (Fetching each entry once per row and
// passing it round the row's 8 lanes with width-8 shuffles was measured in round 2: 15 % SLOWER — shuffles run on
// the same MIO/shared-memory pipe that bounds these kernels.)
// T2 = 2 (L~ T1) - X, one FMA per element (2 g is exact, so this rounds like the subtraction of the doubled value)
__device__ __forceinline__ float4 cheb_t2(const float4& g, const float4& x) {
  return make_float4(fmaf(2.f, g.x, -x.x), fmaf(2.f, g.y, -x.y), fmaf(2.f, g.z, -x.z), fmaf(2.f, g.w, -x.w));
}
__device__ __forceinline__ float4 gather_row4(uint32_t ent, uint32_t e, uint32_t e1, uint32_t rows_q) {
  float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
  for (; e + 3 < e1; e += 4) {
    const uint2 a0 = lds_u2(ent + e * 8), a1 = lds_u2(ent + e * 8 + 8), a2 = lds_u2(ent + e * 8 + 16),
                a3 = lds_u2(ent + e * 8 + 24);
    const float4 x0 = lds_f4(rows_q + a0.x), x1 = lds_f4(rows_q + a1.x), x2 = lds_f4(rows_q + a2.x),
                 x3 = lds_f4(rows_q + a3.x);
    fma4(acc0, __uint_as_float(a0.y), x0);
    fma4(acc1, __uint_as_float(a1.y), x1);
    fma4(acc0, __uint_as_float(a2.y), x2);
    fma4(acc1, __uint_as_float(a3.y), x3);
  }
  if (e + 1 < e1) {
    const uint2 a0 = lds_u2(ent + e * 8), a1 = lds_u2(ent + e * 8 + 8);
    const float4 x0 = lds_f4(rows_q + a0.x), x1 = lds_f4(rows_q + a1.x);
    fma4(acc0, __uint_as_float(a0.y), x0);
    fma4(acc1, __uint_as_float(a1.y), x1);
    e += 2;
  }
  if (e < e1) {
    const uint2 a0 = lds_u2(ent + e * 8);
    fma4(acc0, __uint_as_float(a0.y), lds_f4(rows_q + a0.x));
  }
  return make_float4(acc0.x + acc1.x, acc0.y + acc1.y, acc0.z + acc1.z, acc0.w + acc1.w);
}

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
  int n_tiles;
  const unsigned char* meta;
  const int* meta_bytes;
  int meta_stride, max_h1, max_h2;
  const unsigned char* wpack;
  const float* zero_row;  // 128 bytes of zeros: source of the empty halo slots of ragged tiles
  EpiDev ep;
  int res_identity;       // residual resampling is the identity (Fin_block == Fout): vector path
  const float* t1;        // optional precomputed T1 = L~ x, [rows, fin] at LOGICAL rows (k_cheb_t1): the conv kernel
                          // then stages the tile's own X rows and the T1 rows of its 1-hop halo, and only runs the
                          // second sparse product on chip (no halo recomputation of T1)
  int plain;              // 1: plain GEMM y = x * B^T (no SpMM): one K-block per chunk, rows = the tile's own
  const float* a_scale;   // optional device scalar: x is multiplied by it before the fp16 split (power of two,
                          // chosen from max|x|: gradients are far below fp16's range) and divided out afterwards
  long long ldy;          // row stride of y in floats, and first output column
  int y_col0;
  float* y;
  int* status;
  long long* trace;  // optional [8][512] event log of CTA 0 (debug): (event << 48) | clock
  int own_table;         // 1: the tile's rows are the index list at the head of its metadata blob (TileSet), not
                         //    the consecutive rows [128 pat, 128 pat + 128)
  const float* head_wt;  // optional fused thin head (N == 64): epilogue writes head_z[row][12] = y_row * head_wt[64][12]
  float* head_z;
  // Dense GEMM mode (launch_umma_gemm: PoseNet's Linear layers): the A operand is PRE-PACKED like a weight image
  // (k_pack_plain on the activation matrix: one 16 KB block per (128-row tile, 32-feature chunk)), so both operands of
  // every K-block arrive by cp.async.bulk and no producer warp runs; blockIdx.y selects an N-wide slice of the output
  // columns (its weight image, epilogue vectors and output / residual columns).
  const unsigned char* apack;
  long long wslice_bytes;  // bytes of one N-slice's weight image
  int tma;           // 1: the tile's own rows of x (and t1) arrive by one 2-D TMA load each (T1-given / plain mode on
                     //    levels whose size is a multiple of 128); with in_unpool the x box is the 64 source rows
  CUtensorMap tm_x, tm_t1;
};

// Event timeline of CTA 0 (tools/umma_trace.py): compiled in only with -DP2M_UMMA_TRACE (build.py: P2M_TRACE=1);
// production kernels carry no clock reads.
#ifdef P2M_UMMA_TRACE
__device__ __forceinline__ void trace_ev(const KParams& p, int role, int& n, int ev) {
  if (p.trace != nullptr && blockIdx.x == 0 && n < 512) {
    p.trace[role * 512 + n] = ((long long)ev << 48) | (clock64() & 0xFFFFFFFFFFFFll);
    ++n;
  }
}
#else
__device__ __forceinline__ void trace_ev(const KParams&, int, int&, int) {}
#endif

// Warp roles (24 warps = 6 warpgroups, one persistent CTA per SM):
//   0..15  producers: SpMM out of shared memory + fp16 (hi,lo) split + swizzled A-block stores
//   16,17  loaders: stage the T1 (and, outside the production mode, X) rows of the next chunk; thread 0 also fetches
//          the next tile's metadata (cp.async.bulk), thread 32 issues the TMA boxes
//   18     weight-block loader (one thread, cp.async.bulk)
//   19     MMA issuer (one thread) and TMEM owner
//   20..23 epilogue: TMEM -> registers -> fused epilogue -> HBM, overlapped with the next tile's main loop
constexpr int W_PROD = 16;
constexpr int W_XLOAD = 16, N_XLOAD = 2, W_BLOAD = 18, W_MMA = 19, W_EPI0 = 20;
constexpr int NUM_THREADS2 = 24 * 32;
constexpr int REGS_LAUNCH = 80, REGS_UTIL = 56, REGS_EPI = 104;  // setmaxnreg targets per warpgroup (see the kernel)
// setmaxnreg.inc draws on the pool the CTA's own setmaxnreg.dec filled: requests beyond it would spin forever
static_assert(128 * (REGS_EPI - REGS_LAUNCH) <= 128 * (REGS_LAUNCH - REGS_UTIL), "register pool balance");
static_assert(W_XLOAD % 4 == 0 && W_EPI0 % 4 == 0 && W_EPI0 - W_XLOAD == 4, "setmaxnreg works on aligned warpgroups");

// RM = 1: the residual is the 2:1 channel resampling of a 2N-wide tensor (F.interpolate(linear, align_corners=False)
// from 2N to N channels is the mean of channel pairs: src = 2j + 0.5), fetched as two 16-byte loads per four outputs
// instead of the generic 2-tap gather; a separate instantiation so that the other layers keep their register budget.
// MODE 1: the production configuration — T1 given, not the plain-GEMM mode — fixed at compile time, so the on-chip first
// sparse product, the plain path and their per-tile state drop out of the producers (registers for a deeper gather);
// MODE 0 decides both at run time (plain GEMM, dense GEMM, the split_t1 = 0 ablation).
template <int N, int NS, int XS, int RM = 0, int MODE = 0>
__global__ void __launch_bounds__(NUM_THREADS2, 1) k_cheb_conv_umma(const __grid_constant__ KParams p) {
  constexpr bool PAIR = (RM == 1);
  constexpr bool KT1 = (MODE == 1);
  // X of the own rows straight from global memory (production mode, deep ring): the only reader of an own X row is the
  // producer thread that emits it, so staging it costs a shared-memory write plus a read back (8 % of the kernel's
  // shared-memory traffic, and more than half of the loaders' copies on index-list tiles) for nothing
  constexpr bool XDIRECT = KT1 && NS >= 3;
  const bool plain = KT1 ? false : (p.plain != 0);
  constexpr int B_BLOCK_BYTES = N * 128;
  constexpr int SLOT_BYTES = A_BLOCK_BYTES + B_BLOCK_BYTES;
  constexpr uint32_t IDESC = make_idesc_f16(TILE_M, N);
  constexpr int TMEM_COLS = 2 * N;  // double-buffered accumulator (N = 64/128/256 -> 128/256/512 columns)

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* ring = smem_raw;  // 128B-swizzled blocks need 1024-byte alignment (checked below)
  const bool t1g = KT1 ? true : (p.t1 != nullptr);
  float* Xs = reinterpret_cast<float*>(ring + NS * SLOT_BYTES);                 // [XS][max_h2 | 128][32]
  const size_t xs_stage_floats = (size_t)((t1g || plain) ? TILE_M : p.max_h2) * FC;
  float* T1s = Xs + XS * xs_stage_floats;                                       // [1 | XS | 0][max_h1][32]
  const size_t t1_stage_floats = (size_t)p.max_h1 * FC;
  unsigned char* meta_s =
      reinterpret_cast<unsigned char*>(T1s + (plain ? 0 : (t1g ? XS : 1)) * t1_stage_floats);  // [2][meta_stride]
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta_s + 2 * (size_t)p.meta_stride);
  // barrier map
  uint64_t* b_ab_full = bars;                // [NS]
  uint64_t* b_ab_empty = b_ab_full + NS;     // [NS]
  uint64_t* b_x_full = b_ab_empty + NS;      // [XS]
  uint64_t* b_x_empty = b_x_full + XS;       // [XS]
  uint64_t* b_m_full = b_x_empty + XS;       // [2]
  uint64_t* b_m_empty = b_m_full + 2;        // [2]
  uint64_t* b_acc_full = b_m_empty + 2;      // [2]
  uint64_t* b_acc_empty = b_acc_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_acc_empty + 2);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* ep_mul = reinterpret_cast<float*>(tmem_slot + 4);  // [N] acc * mul + add  (weight scale, bias, folded BN)
  float* ep_add = ep_mul + N;
  // epilogue transpose staging: [4 warps][32 rows][EC floats], 16-byte chunks XOR-swizzled by the row
  constexpr int EC = (N == 256) ? 16 : 32;
  unsigned char* epi_stage =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ep_add + N) + 127) & ~(uintptr_t)127);
  int* own_s = reinterpret_cast<int*>(epi_stage + 4 * 32 * EC * 4);  // [4 warps][32] vertex id of each epilogue row
  float* head_w_s = reinterpret_cast<float*>(own_s + 4 * 32);        // [64][12] (N == 64 with a fused head)

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int n_chunk = p.fin / FC;
  const int n_use = 3 * n_chunk;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      // one elected arrive per producer warp + the weight loader (dense GEMM mode: the loader alone)
      mbar_init(smem_u32(b_ab_full + s), p.apack != nullptr ? 1 : W_PROD + 1);
      mbar_init(smem_u32(b_ab_empty + s), 1);
    }
    for (int s = 0; s < XS; ++s) {
      mbar_init(smem_u32(b_x_full + s), N_XLOAD * 32 + 1);  // every loader thread (cp.async.mbarrier.arrive.noinc)
                                                            // + one expect_tx (TMA) / plain arrival of loader thread 32
      mbar_init(smem_u32(b_x_empty + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(b_m_full + s), 1);
      mbar_init(smem_u32(b_m_empty + s), 1);
      mbar_init(smem_u32(b_acc_full + s), 1);
      mbar_init(smem_u32(b_acc_empty + s), 4);
    }
    *abort_flag = (smem_u32(ring) & 1023u) ? 1 : 0;
    if (*abort_flag) mbar_timeout(abort_flag, p.status, 100);
    fence_barrier_init();
  }
  const float a_scale = p.a_scale ? *p.a_scale : 1.f;
  const int ecol0 = (int)blockIdx.y * N;  // first output column of this CTA's slice (0 outside the dense GEMM mode)
  for (int n = threadIdx.x; n < N; n += NUM_THREADS2) {
    const float sc = (p.ep.scale ? p.ep.scale[ecol0 + n] : 1.f) / a_scale;
    const float sh = p.ep.scale ? p.ep.shift[ecol0 + n] : 0.f;
    const float bi = p.ep.bias ? p.ep.bias[ecol0 + n] : 0.f;
    ep_mul[n] = W_INV_SCALE * sc;
    ep_add[n] = fmaf(bi, p.ep.scale ? p.ep.scale[ecol0 + n] : 1.f, sh);
  }
  if (N == 64 && p.head_z != nullptr)
    for (int i = threadIdx.x; i < 64 * 12; i += NUM_THREADS2) head_w_s[i] = p.head_wt[i];
  if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Register split by warpgroup (the kernel is launched with 80 per thread: 768 x 80 = 60 K of the SM's 64 K): the
  // utility warpgroup (loaders, weight loader, MMA issuer) hands back half of its share and the epilogue warpgroup takes
  // exactly that (120 per thread) — enough to keep a sub-slab of residual pieces, its vertex ids and affine
  // coefficients in registers instead of re-reading them from shared memory per piece.  The producers stay at 80.
  // (each budget is set at the top of its own region: after a join ptxas has to assume the smallest one)
  if (warp >= W_XLOAD && warp < W_EPI0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_UTIL));
  if (warp >= W_XLOAD && warp < W_XLOAD + N_XLOAD) {
    // ------------------------------------------------------------ loaders (two warps): tile metadata, own rows, halo rows
    // Per stage (tile, 32-feature chunk) they bring in what the producers read: the tile's metadata blob (thread 0,
    // cp.async.bulk, one tile ahead), the own rows of X / T1 (one 2-D TMA box each where the tile is a run of
    // consecutive rows — thread 32), and every other staged row with 16-byte cp.async copies (8 lanes per 128-byte row,
    // all 64 threads): the T1 rows of the 1-hop halo, or all rows where the tile is an index list / no T1 is given.
    // A stage needs ~60 warp-level copies; issued by the 16 producer warps (round 2 until this change) every warp paid
    // the whole preamble for its one to four rows — a fifth of the producers' instruction stream per chunk.
    if (p.apack == nullptr) {
      const int lt = tid - W_XLOAD * 32;  // 0..63
      const int lq = lt & 7, lrg = lt >> 3;
      const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      if (lt == 32 && p.tma) {
        if (!XDIRECT) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tm_x)) : "memory");
        if (t1g) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tm_t1)) : "memory");
      }
      auto fetch_meta = [&](int itf) {  // thread 0: blob of this CTA's tile number itf into buffer itf & 1
        const int pat = (blockIdx.x + itf * gridDim.x) % p.P;
        const int m = itf & 1;
        mbar_wait_relaxed(smem_u32(b_m_empty + m), ((itf >> 1) & 1) ^ 1, abort_flag, p.status, 1);
        const int mbytes = p.meta_bytes[pat];
        mbar_arrive_expect_tx(smem_u32(b_m_full + m), mbytes);
        bulk_g2s(smem_u32(meta_s + (size_t)m * p.meta_stride), p.meta + (size_t)pat * p.meta_stride, mbytes,
                 smem_u32(b_m_full + m));
      };
      if (lt == 0 && my_tiles > 0) fetch_meta(0);
      const uint32_t fin_bytes = (uint32_t)p.fin * 4u;
      const int sh = p.in_unpool ? 1 : 0;
      int g2 = 0;
      int ltn = 0;
      for (int it2 = 0; it2 < my_tiles; ++it2) {
        const int tile2 = blockIdx.x + it2 * gridDim.x;
        const long long mesh_row0 = (long long)(tile2 / p.P) * p.V;
        const int m2 = it2 & 1;
        mbar_wait_relaxed(smem_u32(b_m_full + m2), (it2 >> 1) & 1, abort_flag, p.status, 8);
        const unsigned char* mb2 = meta_s + (size_t)m2 * p.meta_stride;
        const TileHeader* hdr2 = reinterpret_cast<const TileHeader*>(mb2);
        const int* halo = reinterpret_cast<const int*>(mb2 + hdr2->off_halo);
        const int h1 = hdr2->h1, h2 = hdr2->h2;
        for (int c2 = 0; c2 < n_chunk; ++c2, ++g2) {
          const int xs2 = g2 % XS;
          mbar_wait_relaxed(smem_u32(b_x_empty + xs2), ((g2 / XS) & 1) ^ 1, abort_flag, p.status, 11);
          if (lt == 0) trace_ev(p, 4, ltn, 1);
          const uint32_t xbar = smem_u32(b_x_full + xs2);
          // a row's byte offset inside its mesh fits 32 bits: one 32 x 32 -> 64-bit multiply-add per row; empty slots
          // (-1) are zero-filled by the copy itself (src-size 0)
          auto stage_rows = [&](uint32_t dbase, const char* sbase, int first, int n_rows, int shift) {
            for (int i0 = first + lrg; i0 < n_rows; i0 += 32) {
              int v[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) v[u] = (i0 + 8 * u < n_rows) ? halo[i0 + 8 * u] : -2;
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (v[u] != -2) {
                  const uint32_t srow = (uint32_t)max(v[u], 0) >> shift;
                  cp_async16_zfill(dbase + (i0 + 8 * u) * 128, sbase + (uint64_t)srow * (uint64_t)fin_bytes,
                                   v[u] >= 0 ? 16u : 0u);
                }
              }
            }
          };
          const char* t1_mesh =
              reinterpret_cast<const char*>(t1g ? p.t1 + mesh_row0 * p.fin + c2 * FC + lq * 4 : nullptr);
          const char* x_mesh = reinterpret_cast<const char*>(p.x + (mesh_row0 >> sh) * p.fin + c2 * FC + lq * 4);
          const uint32_t t1_dst = smem_u32(T1s + xs2 * t1_stage_floats) + lq * 16;
          const uint32_t x_dst = smem_u32(Xs + xs2 * xs_stage_floats) + lq * 16;
          if (p.tma) {
            if (lt == 32) {
              const int own0 = tile2 * TILE_M;  // V is a multiple of 128: tiles never straddle meshes
              mbar_arrive_expect_tx(xbar, (XDIRECT ? 0 : (p.in_unpool ? TILE_M / 2 : TILE_M) * 128) + (t1g ? TILE_M * 128 : 0));
              if (!XDIRECT)
                tma_load_2d(smem_u32(Xs + xs2 * xs_stage_floats), &p.tm_x, c2 * FC, p.in_unpool ? own0 >> 1 : own0, xbar);
              if (t1g) tma_load_2d(smem_u32(T1s + xs2 * t1_stage_floats), &p.tm_t1, c2 * FC, own0, xbar);
            }
            if (t1g) stage_rows(t1_dst, t1_mesh, TILE_M, h1, 0);  // only the halo rows are left
          } else {
            if (t1g) stage_rows(t1_dst, t1_mesh, 0, h1, 0);
            if (!XDIRECT) stage_rows(x_dst, x_mesh, 0, (plain || t1g) ? TILE_M : h2, sh);
            if (lt == 32) mbar_arrive(xbar);
          }
          cp_async_arrive_noinc(xbar);  // this thread's arrival once its copies have landed
          if (lt == 0) trace_ev(p, 4, ltn, 2);
          // next tile's metadata: its buffer was last read for tile it2 - 1, which the producers have left by the time
          // the second chunk of this tile could be staged (single-chunk layers: by the time its only chunk could)
          if (lt == 0 && c2 == (n_chunk > 1 ? 1 : 0) && it2 + 1 < my_tiles) fetch_meta(it2 + 1);
        }
      }
    }
  } else if (warp == W_BLOAD) {
    // ------------------------------------------------------------ weight-block loader (one thread)
    if (lane == 0) {
      uint32_t ucnt = 0;
      int tn = 0;
      const int uses = plain ? n_chunk : n_use;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int u = 0; u < uses; ++u, ++ucnt) {
          const int s = ucnt % NS;
          const uint32_t par = (ucnt / NS) & 1;
          mbar_wait_relaxed(smem_u32(b_ab_empty + s), par ^ 1, abort_flag, p.status, 4);
          trace_ev(p, 1, tn, 10 + u);
          const unsigned char* wsl = p.wpack + (size_t)blockIdx.y * (size_t)p.wslice_bytes;
          if (p.apack != nullptr) {  // dense GEMM mode: the tile's pre-packed A block of chunk u rides along
            mbar_arrive_expect_tx(smem_u32(b_ab_full + s), B_BLOCK_BYTES + A_BLOCK_BYTES);
            bulk_g2s(smem_u32(ring + s * SLOT_BYTES), p.apack + ((size_t)tile * uses + u) * A_BLOCK_BYTES, A_BLOCK_BYTES,
                     smem_u32(b_ab_full + s));
          } else {
            mbar_arrive_expect_tx(smem_u32(b_ab_full + s), B_BLOCK_BYTES);
          }
          bulk_g2s(smem_u32(ring + s * SLOT_BYTES + A_BLOCK_BYTES), wsl + (size_t)u * B_BLOCK_BYTES, B_BLOCK_BYTES,
                   smem_u32(b_ab_full + s));
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      uint32_t ucnt = 0;
      int it = 0;
      int tn = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        mbar_wait(smem_u32(b_acc_empty + as), ((it >> 1) & 1) ^ 1, abort_flag, p.status, 5);
        trace_ev(p, 2, tn, 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * N);
        const int uses = plain ? n_chunk : n_use;
        for (int u = 0; u < uses; ++u, ++ucnt) {
          const int s = ucnt % NS;
          mbar_wait(smem_u32(b_ab_full + s), (ucnt / NS) & 1, abort_flag, p.status, 6);
          trace_ev(p, 2, tn, 10 + u);
          tc_fence_after();
          const uint32_t a0 = smem_u32(ring + s * SLOT_BYTES);
          const uint64_t da = make_desc_sw128(a0), db = make_desc_sw128(a0 + A_BLOCK_BYTES);
          // A block columns: [hi 0..31 | lo 32..63], B block columns: [Whi 0..31 | Wlo 32..63] (fp16);
          // a 16-element K step is 32 bytes = +2 in the descriptor's start-address field.
          umma_f16(d_tmem, da + 0, db + 0, IDESC, u > 0);  // hi * Whi
          umma_f16(d_tmem, da + 2, db + 2, IDESC, 1);
          umma_f16(d_tmem, da + 4, db + 0, IDESC, 1);      // lo * Whi
          umma_f16(d_tmem, da + 6, db + 2, IDESC, 1);
          umma_f16(d_tmem, da + 0, db + 4, IDESC, 1);      // hi * Wlo
          umma_f16(d_tmem, da + 2, db + 6, IDESC, 1);
          umma_commit(smem_u32(b_ab_empty + s));  // frees the slot when these MMAs have read it
        }
        umma_commit(smem_u32(b_acc_full + as));
      }
    }
  }
  } else if (warp >= W_EPI0) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
    // ------------------------------------------------------------ epilogue: TMEM -> registers -> (transpose) -> HBM
    // tcgen05.ld hands every thread one accumulator ROW; storing that directly makes each warp-wide 16-byte store
    // touch 32 different lines.  The rows are therefore transposed through a small per-warp staging buffer so that
    // a warp-wide access covers whole 128-byte (N = 256: 64-byte) row pieces of a few rows; the fused epilogue
    // (affine, ReLU, residual) runs in the transposed layout, where the residual reads coalesce as well.
    const int lane_base = (warp & 3) * 32;  // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    constexpr int CPR = EC / 4;             // 16-byte chunks per staged row
    constexpr int RPI = 32 / CPR;           // rows covered by one warp-wide 16-byte access
    const uint32_t stg = smem_u32(epi_stage) + (uint32_t)(warp - W_EPI0) * (32 * EC * 4);
    const int prow = lane / CPR, pc = lane % CPR;
    constexpr int NP = 32 / RPI;              // phase-2 rows (pieces) per thread and sub-slab
    constexpr int NSW = (EC == 32) ? 2 : 1;   // distinct column chunks a thread touches per sub-slab (swizzle: row & 7)
    const bool res_reg = (PAIR || (p.ep.res != nullptr && p.res_identity)) && !(N == 64 && p.head_z != nullptr);
    int colk[NSW];  // column (inside a sub-slab) of the 16-byte chunk this thread handles for rows of swizzle class k
#pragma unroll
    for (int k = 0; k < NSW; ++k) {
      const int rr = k * RPI + prow;
      const uint32_t sw2 = (EC == 32) ? (uint32_t)(rr & 7) : (uint32_t)((rr >> 1) & 3);
      colk[k] = (int)(((uint32_t)pc ^ sw2) << 2);
    }
    const uint32_t sw1 = (EC == 32) ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
    int it = 0;
    int etn = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int b = tile / p.P, pat = tile - b * p.P;
      const int as = it & 1;
      // vertex id of each of this warp's 32 accumulator rows (-1: the slot holds no vertex)
      const long long mesh0 = (long long)b * p.V;
      int* own_w = own_s + (warp - W_EPI0) * 32;
      {
        int v;
        if (p.own_table) {
          v = __ldg(reinterpret_cast<const int*>(p.meta + (size_t)pat * p.meta_stride + 64) + lane_base + lane);
        } else {
          v = pat * TILE_M + lane_base + lane;
          if (v >= p.V) v = -1;
        }
        __syncwarp();  // the previous tile's readers of own_w are done
        own_w[lane] = v;
        __syncwarp();
      }
      // vertex id of each of this thread's NP phase-2 rows, kept in registers for the whole tile (every shared-memory
      // instruction saved here matters: the epilogue warps queue behind the producers on the same LSU/MIO pipe, and on
      // the layers with a residual their ~250 instructions per tile made them the slowest role — tools/umma_trace_model.py)
      int own_v[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) own_v[i] = own_w[i * RPI + prow];
      if (p.ep.res != nullptr) {
        // pull this tile's residual rows into L2 while its main loop is still running: the reads below then pay
        // an L2 hit instead of a DRAM round trip per batch
        // 128-byte lines per residual row (dense GEMM mode: only this CTA's N-column slice of the row)
        const int lpr = ((p.apack != nullptr ? N : p.ep.res_F) * 4 + 127) >> 7;
        for (int j = lane; j < 32 * lpr; j += 32) {
          const int rr = j / lpr, ln = j - rr * lpr;
          const int vtx = own_w[rr];
          if (vtx >= 0) {
            const long long r = mesh0 + vtx;
            prefetch_l2(p.ep.res + (p.ep.res_unpool ? (r >> 1) : r) * p.ep.res_F + (p.apack != nullptr ? ecol0 : 0) + ln * 32);
          }
        }
      }
      // Residual pieces that are added straight from registers (identity / pair-mean): rv[][] always holds the pieces
      // of the NEXT 32 columns — the first 32 are issued here, before the accumulator wait, and each piece is re-issued
      // for the column 32 further on as soon as it has been consumed, so a piece has a whole 32-column slab of time to
      // arrive (one batch of four at a time exposed ~8 L2 round trips per tile).
      uint32_t res_off[NP];  // float index of the residual row of piece i (row * res_F < 2^32)
      float4 rv[32 / EC][NP];
      auto res_piece = [&](int i, int n) -> float4 {
        if (own_v[i] < 0) return make_float4(0.f, 0.f, 0.f, 0.f);
        const float* row = p.ep.res + (size_t)res_off[i];
        if (PAIR) {
          const float4* src = reinterpret_cast<const float4*>(row + 2 * n);
          const float4 a = __ldg(src), b = __ldg(src + 1);
          return make_float4(0.5f * a.x + 0.5f * a.y, 0.5f * a.z + 0.5f * a.w, 0.5f * b.x + 0.5f * b.y,
                             0.5f * b.z + 0.5f * b.w);
        }
        return __ldg(reinterpret_cast<const float4*>(row + ecol0 + n));
      };
      if (res_reg) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const long long r = mesh0 + max(own_v[i], 0);
          res_off[i] = (uint32_t)((p.ep.res_unpool ? (r >> 1) : r) * p.ep.res_F);
#pragma unroll
          for (int h = 0; h < 32 / EC; ++h) rv[h][i] = res_piece(i, h * EC + colk[i & (NSW - 1)]);
        }
      }
      // only the first epilogue warp polls the mbarrier, the other three sleep in a hardware barrier until it has seen
      // the accumulator (four polling warps were 12 % of the kernel's executed instructions: a suspended try_wait wakes
      // on every mbarrier event of the CTA, i.e. every ~100 cycles here)
      bool released = false;
      if (warp == W_EPI0) mbar_wait_relaxed(smem_u32(b_acc_full + as), (it >> 1) & 1, abort_flag, p.status, 7);
      epilogue_barrier();
      if (warp == W_EPI0 && lane == 0) trace_ev(p, 3, etn, 1);
      tc_fence_after();
      if (N == 64 && p.head_z != nullptr) {
        // fused thin head: thread = row; y_n = act(acc_n * mul_n + add_n) never leaves the registers, only the
        // 12 projections Z = y W' (W' = [W0 - W2 | W1 | W2] of the 64 -> 3 layer, 4-padded) are written
        float z[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) z[j] = 0.f;
#pragma unroll 1
        for (int cb = 0; cb < N; cb += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(as * N + cb), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = cb + j;
            float t = fmaf(__uint_as_float(v[j]), ep_mul[n], ep_add[n]);
            if (p.ep.relu) t = fmaxf(t, 0.f);
            const float4 w0 = *reinterpret_cast<const float4*>(head_w_s + n * 12);
            const float4 w1 = *reinterpret_cast<const float4*>(head_w_s + n * 12 + 4);
            const float4 w2 = *reinterpret_cast<const float4*>(head_w_s + n * 12 + 8);
            z[0] = fmaf(t, w0.x, z[0]); z[1] = fmaf(t, w0.y, z[1]); z[2] = fmaf(t, w0.z, z[2]); z[3] = fmaf(t, w0.w, z[3]);
            z[4] = fmaf(t, w1.x, z[4]); z[5] = fmaf(t, w1.y, z[5]); z[6] = fmaf(t, w1.z, z[6]); z[7] = fmaf(t, w1.w, z[7]);
            z[8] = fmaf(t, w2.x, z[8]); z[9] = fmaf(t, w2.y, z[9]); z[10] = fmaf(t, w2.z, z[10]); z[11] = fmaf(t, w2.w, z[11]);
          }
        }
        if (own_w[lane] >= 0) {
          float4* zr = reinterpret_cast<float4*>(p.head_z + (mesh0 + own_w[lane]) * 12);
          zr[0] = make_float4(z[0], z[1], z[2], z[3]);
          zr[1] = make_float4(z[4], z[5], z[6], z[7]);
          zr[2] = make_float4(z[8], z[9], z[10], z[11]);
        }
      } else
#pragma unroll 1
      for (int cb = 0; cb < N; cb += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(as * N + cb), v);
        if (cb == N - 32) {  // the whole accumulator has been read: hand the TMEM buffer back before the last slab's stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(b_acc_empty + as));
          released = true;
        }
#pragma unroll
        for (int h = 0; h < 32 / EC; ++h) {
          // phase 1: lane = row
#pragma unroll
          for (int c = 0; c < CPR; ++c)
            sts_f4(stg + lane * (EC * 4) + (((uint32_t)c ^ sw1) << 4),
                   make_float4(__uint_as_float(v[h * EC + 4 * c]), __uint_as_float(v[h * EC + 4 * c + 1]),
                               __uint_as_float(v[h * EC + 4 * c + 2]), __uint_as_float(v[h * EC + 4 * c + 3])));
          __syncwarp();
          // phase 2: lane = (row inside an RPI-row group, 16-byte chunk).  The residual pieces of the whole slice
          // are fetched up front (read-only path): behind the y stores the compiler could not batch them.
          // the thread's output columns of this sub-slab (two alternating 16-byte chunks for EC = 32, one for EC = 16)
          // and their affine coefficients: fetched once per sub-slab instead of once per piece
          const int cbh = cb + h * EC;
          float4 mu_k[NSW], ad_k[NSW];
#pragma unroll
          for (int k = 0; k < NSW; ++k) {
            mu_k[k] = *reinterpret_cast<const float4*>(ep_mul + cbh + colk[k]);
            ad_k[k] = *reinterpret_cast<const float4*>(ep_add + cbh + colk[k]);
          }
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const int rr = i * RPI + prow;
            const int n = cbh + colk[i & (NSW - 1)];
            const float4 a = lds_f4(stg + rr * (EC * 4) + (pc << 4));
            const int vtx = own_v[i];
            if (vtx >= 0) {
              const float4 mu = mu_k[i & (NSW - 1)], ad = ad_k[i & (NSW - 1)];
              float o[4] = {fmaf(a.x, mu.x, ad.x), fmaf(a.y, mu.y, ad.y), fmaf(a.z, mu.z, ad.z), fmaf(a.w, mu.w, ad.w)};
              if (p.ep.relu) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
              }
              const long long r = mesh0 + vtx;
              if (res_reg) {
                o[0] += rv[h][i].x; o[1] += rv[h][i].y; o[2] += rv[h][i].z; o[3] += rv[h][i].w;
              } else if (p.ep.res != nullptr) {
                const float* res_row = p.ep.res + (p.ep.res_unpool ? (r >> 1) : r) * p.ep.res_F;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float l = __ldg(p.ep.lam + n + e);
                  o[e] += (1.f - l) * __ldg(res_row + __ldg(p.ep.i0 + n + e)) + l * __ldg(res_row + __ldg(p.ep.i1 + n + e));
                }
              }
              *reinterpret_cast<float4*>(p.y + r * p.ldy + p.y_col0 + ecol0 + n) = make_float4(o[0], o[1], o[2], o[3]);
            }
            if (res_reg && cb + 32 < N) rv[h][i] = res_piece(i, n + 32);  // same chunk, 32 columns further on
          }
          __syncwarp();
        }
      }
      if (warp == W_EPI0 && lane == 0) trace_ev(p, 3, etn, 2);
      if (!released) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(b_acc_empty + as));
      }
    }
  } else {
    if (p.apack == nullptr) {
    // ------------------------------------------------------------ producers (16 warps)
    const int q = tid & 7;     // float4 lane inside the 32-feature chunk
    const int rg = tid >> 3;   // row group 0..63
    const uint32_t t1s_a = smem_u32(T1s);
    const uint32_t ring_a = smem_u32(ring);
    const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_stage = my_tiles * n_chunk;  // flat sequence of (tile, chunk) stages of this CTA

    constexpr int T1_ROWS = 4;  // max_h1 <= 256 rows over 64 row groups (checked on the host)
    uint32_t t1_row[T1_ROWS], t1_e[T1_ROWS];
    uint32_t row0 = 0, row1 = 0, r0e = 0, r1e = 0, ent_a = 0;
    int ptn = 0;
    // A/B ring cursor (slot, phase parity) kept incrementally, and the four store offsets of this thread's two rows
    // inside an A block (first / second 8-byte store of each row: odd row groups store lo first, see emit()) —
    // per-tile constants: the hot loop carries no modulo, division or swizzle arithmetic
    uint32_t slot = 0, spar = 0;
    uint32_t so_f[2] = {0, 0}, so_s[2] = {0, 0};
    const bool odd = (rg & 1) != 0;
    auto next_slot = [&]() {
      if (++slot == (uint32_t)NS) {
        slot = 0;
        spar ^= 1u;
      }
    };

    const int xsh = (p.tma && p.in_unpool) ? 1 : 0;  // TMA-staged unpooled input: staged row = tile row >> 1
    const float* xp0 = nullptr;  // XDIRECT: this thread's two own X rows (nullptr: empty slot), at its 16-byte column
    const float* xp1 = nullptr;
    int it = 0, c = -1;
    for (int g = 0; g < n_stage; ++g) {
      if (++c == n_chunk) {
        c = 0;
        ++it;
      }
      const int m = it & 1;
      const int xs = g % XS;
      float4 xd0 = make_float4(0.f, 0.f, 0.f, 0.f), xd1 = xd0;
      if (XDIRECT && c != 0) {  // in flight while the stage wait and the gather run (first chunk: below, after the tile setup)
        if (xp0) xd0 = __ldg(reinterpret_cast<const float4*>(xp0 + c * FC));
        if (xp1) xd1 = __ldg(reinterpret_cast<const float4*>(xp1 + c * FC));
      }
      if (tid == 0) trace_ev(p, 0, ptn, 1);
      mbar_wait(smem_u32(b_x_full + xs), (g / XS) & 1, abort_flag, p.status, 9);
      if (tid == 0) trace_ev(p, 0, ptn, 2);

      if (c == 0) {  // per-tile bookkeeping: which rows this thread owns and where their CSR rows start/end
        mbar_wait(smem_u32(b_m_full + m), (it >> 1) & 1, abort_flag, p.status, 8);  // (long complete: the loaders read it)
        const unsigned char* mb = meta_s + (size_t)m * p.meta_stride;
        const TileHeader* hdr = reinterpret_cast<const TileHeader*>(mb);
        const uint32_t mb_a = smem_u32(mb);
        const uint32_t rp_a = mb_a + hdr->off_rp, ord1_a = mb_a + hdr->off_ord1, ord2_a = mb_a + hdr->off_ord2;
        ent_a = mb_a + hdr->off_ent;
        const int h1 = (t1g || plain) ? 0 : hdr->h1;  // the trimmed metadata has no T1 row order
#pragma unroll
        for (int t = 0; t < T1_ROWS; ++t) {
          const int j = rg + 64 * t;
          t1_row[t] = 0xFFFFu;
          if (j < h1) {
            const uint32_t i = lds_u16(ord1_a + 2 * j);
            t1_row[t] = i;
            t1_e[t] = lds_u16(rp_a + 2 * i) | (lds_u16(rp_a + 2 * i + 2) << 16);
          }
        }
        row0 = lds_u16(ord2_a + 2 * rg);
        row1 = lds_u16(ord2_a + 2 * (64 + rg));
        r0e = lds_u16(rp_a + 2 * row0) | (lds_u16(rp_a + 2 * row0 + 2) << 16);
        r1e = lds_u16(rp_a + 2 * row1) | (lds_u16(rp_a + 2 * row1 + 2) << 16);
        if (plain) {  // plain GEMM: the thread's rows are the consecutive slots rg and 64 + rg
          row0 = rg;
          row1 = 64 + rg;
        }
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          const uint32_t i = ps ? row1 : row0;
          const uint32_t a_hi = sw128_off(i, q >> 1) + (q & 1) * 8, a_lo = sw128_off(i, 4 + (q >> 1)) + (q & 1) * 8;
          so_f[ps] = odd ? a_lo : a_hi;
          so_s[ps] = odd ? a_hi : a_lo;
        }
        if (XDIRECT) {
          const int* halo = reinterpret_cast<const int*>(mb + hdr->off_halo);  // slots 0..127 = the tile's own rows
          const int v0 = halo[row0], v1 = halo[row1];
          const int sh = p.in_unpool ? 1 : 0;
          const long long mesh_row0 = (long long)((blockIdx.x + (unsigned)it * gridDim.x) / (unsigned)p.P) * p.V;
          const float* xm = p.x + (mesh_row0 >> sh) * p.fin + q * 4;
          xp0 = (v0 >= 0) ? xm + (size_t)((uint32_t)v0 >> sh) * (uint32_t)p.fin : nullptr;
          xp1 = (v1 >= 0) ? xm + (size_t)((uint32_t)v1 >> sh) * (uint32_t)p.fin : nullptr;
          if (xp0) xd0 = __ldg(reinterpret_cast<const float4*>(xp0));
          if (xp1) xd1 = __ldg(reinterpret_cast<const float4*>(xp1));
        }
      }
      const uint32_t xs_q = smem_u32(Xs + xs * xs_stage_floats) + q * 16;
      const uint32_t t1s_q = t1s_a + (t1g ? (uint32_t)(xs * t1_stage_floats * 4) : 0u) + q * 16;
      if (plain) {
        // plain GEMM: the staged rows ARE the A operand (scaled into fp16 range if a_scale is given)
        const uint32_t s = slot;
        mbar_wait(smem_u32(b_ab_empty + s), spar ^ 1u, abort_flag, p.status, 10);
        const uint32_t ablk = ring_a + s * SLOT_BYTES;
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          const uint32_t i = ps * 64 + rg;
          float4 v = lds_f4(xs_q + (i >> xsh) * 128);
          v.x *= a_scale; v.y *= a_scale; v.z *= a_scale; v.w *= a_scale;
          uint2 hi, lo;
          split4(v, hi, lo);
          // the four consecutive rows of a warp share (row & 4), i.e. the 64-byte half their hi parts go to: odd row
          // groups store lo first (selects, see emit()), so that both rows of a half-warp cover different bank halves
          sts_u2(ablk + so_f[ps], odd ? lo : hi);
          sts_u2(ablk + so_s[ps], odd ? hi : lo);
        }
        fence_async_proxy();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(smem_u32(b_ab_full + s));
        next_slot();
        producer_barrier();
        if (tid == 0) mbar_arrive(smem_u32(b_x_empty + xs));  // stage free: the loaders may refill it
        if (tid == 0 && c == n_chunk - 1) mbar_arrive(smem_u32(b_m_empty + m));
        continue;
      }
      // (1) T1 = L~ X on the tile rows and their 1-hop halo (local CSR columns = staged X rows); two rows per
      //     thread are gathered together for memory-level parallelism.
      if (!t1g) {
#pragma unroll
        for (int t = 0; t < T1_ROWS; ++t) {
          if (t1_row[t] != 0xFFFFu)
            sts_f4(t1s_q + t1_row[t] * 128, gather_row4(ent_a, t1_e[t] & 0xFFFFu, t1_e[t] >> 16, xs_q));
        }
        if (tid == 0) trace_ev(p, 0, ptn, 4);
        producer_barrier();
        if (tid == 0) trace_ev(p, 0, ptn, 5);
      }
      // (2) split to fp16 (hi, lo) and write the three K-blocks of the two tile rows this thread finishes.  The X and
      //     T1 blocks go first: they need no gather, so the tensor core starts on them while (3) the second sparse
      //     product T2 = 2 L~ T1 - X is still being gathered (with a 2-deep ring, N = 256, the T2 block re-uses the X
      //     block's slot and would otherwise wait for its MMAs).  NS >= 3: ONE generic->async proxy fence for all
      //     three blocks (the fence drains the thread's outstanding shared stores and is expensive); NS < 3: one per
      //     block.  Odd row groups store lo first: a warp then covers both 64-byte halves of its rows per store.
      const uint32_t slot0 = slot;
      auto emit = [&](const float4& v0, const float4& v1) {
        const uint32_t s = slot;
        if (NS < 3) mbar_wait(smem_u32(b_ab_empty + s), spar ^ 1u, abort_flag, p.status, 10);
        const uint32_t ablk = ring_a + s * SLOT_BYTES;
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          uint2 hi, lo;
          float4 v = ps ? v1 : v0;
          if (p.a_scale != nullptr) {  // backward-data pass: gradients are scaled into fp16's range (power of two)
            v.x *= a_scale; v.y *= a_scale; v.z *= a_scale; v.w *= a_scale;
          }
          split4(v, hi, lo);
          // 64-bit shared stores are served per HALF-warp (two rows here).  Odd row groups store lo first — by
          // selects, not branches: a predicated store would leave each half-warp's wavefront half empty — and the
          // row order pairs rows so that the two 64-byte pieces of a half-warp fall into different bank halves
          sts_u2(ablk + so_f[ps], odd ? lo : hi);
          sts_u2(ablk + so_s[ps], odd ? hi : lo);
        }
        if (NS < 3) {
          fence_async_proxy();
          __syncwarp();
          if ((tid & 31) == 0) mbar_arrive(smem_u32(b_ab_full + s));
        }
        next_slot();
      };
      if (NS < 3) {
        const float4 x0 = lds_f4(xs_q + (row0 >> xsh) * 128), x1 = lds_f4(xs_q + (row1 >> xsh) * 128);
        emit(x0, x1);
        {
          const float4 t10 = lds_f4(t1s_q + row0 * 128), t11 = lds_f4(t1s_q + row1 * 128);
          emit(t10, t11);
        }
        const float4 g0 = gather_row4(ent_a, r0e & 0xFFFFu, r0e >> 16, t1s_q);
        const float4 g1 = gather_row4(ent_a, r1e & 0xFFFFu, r1e >> 16, t1s_q);
        if (tid == 0) trace_ev(p, 0, ptn, 6);
        emit(cheb_t2(g0, x0),
             cheb_t2(g1, x1));
      } else {
        // deep ring: gather first, then the three blocks back to back (measured faster than the early X/T1 emit:
        // the gather then overlaps the previous chunk's tail instead of this chunk's own stores)
        const float4 g0 = gather_row4(ent_a, r0e & 0xFFFFu, r0e >> 16, t1s_q);
        const float4 g1 = gather_row4(ent_a, r1e & 0xFFFFu, r1e >> 16, t1s_q);
        const float4 x0 = XDIRECT ? xd0 : lds_f4(xs_q + (row0 >> xsh) * 128);
        const float4 x1 = XDIRECT ? xd1 : lds_f4(xs_q + (row1 >> xsh) * 128);
        const float4 t10 = lds_f4(t1s_q + row0 * 128), t11 = lds_f4(t1s_q + row1 * 128);
        if (tid == 0) trace_ev(p, 0, ptn, 6);
        {
          // one wait for the chunk's three slots: the MMA issuer commits them in order, so the last one being free
          // implies the other two (each mbarrier wait is a ~200-cycle round trip on the critical path of the chunk)
          uint32_t s2 = slot + 2, p2 = spar;
          if (s2 >= (uint32_t)NS) {
            s2 -= (uint32_t)NS;
            p2 ^= 1u;
          }
          mbar_wait(smem_u32(b_ab_empty + s2), p2 ^ 1u, abort_flag, p.status, 10);
        }
        emit(x0, x1);
        emit(t10, t11);
        emit(cheb_t2(g0, x0),
             cheb_t2(g1, x1));
      }
      if (NS >= 3) {
        fence_async_proxy();
        __syncwarp();
        if ((tid & 31) == 0) {
#pragma unroll
          for (int k = 0; k < 3; ++k) mbar_arrive(smem_u32(b_ab_full + (slot0 + k) % NS));  // one arrival per warp
        }
      }
      if (tid == 0) trace_ev(p, 0, ptn, 7);
      producer_barrier();  // everybody is done with Xs[xs] and T1s
      if (tid == 0) trace_ev(p, 0, ptn, 8);
      if (tid == 0) mbar_arrive(smem_u32(b_x_empty + xs));  // stage free: the loaders may refill it
      if (tid == 0 && c == n_chunk - 1) mbar_arrive(smem_u32(b_m_empty + m));
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, TMEM_COLS);
}

// =====================================================================================
// dW on tensor cores:  dW[o, (f,k)] = sum_rows dz[row, o] * T_k[row, f]
//
// The reduction runs over the ROWS, so both operands are "MN-major" for the tensor core (the K index of the
// MMA is the mesh row).  The 128B-swizzled row-major blocks the forward kernel already builds — 128 rows x
// [hi 32 | lo 32] fp16 of T_k for one 32-feature chunk — are exactly a canonical MN-major SWIZZLE_128B tile
// (K = 128 rows of 128 bytes, MN = 64 elements), so the producers are shared with the forward and T is
// recomputed on chip instead of being materialised (the SIMT path writes and re-reads 3x the activations).
// A = dz tile, scaled by a power of two into fp16 range and split (hi, lo), stored the same way
// ([128 rows] x 64 channels per block, two blocks per 128 channels).  Per T block: 8 K-steps x (hi, lo) MMAs
// with M = 128 channels, N = 64 ([T_hi | T_lo] columns), accumulated in TMEM across ALL tiles of the CTA
// (6 accumulators of 64 columns = two feature chunks per launch); one atomicAdd pass per CTA at the end.
// =====================================================================================
struct DwParams {
  const float* x;
  int in_unpool;
  int V, P, fin;
  int n_tiles;
  const unsigned char* meta;
  const int* meta_bytes;
  int meta_stride, max_h1, max_h2;
  const float* g;        // dz [rows, fout_total]
  int fout_total, m_off, m_cols;
  int chunk0, n_chunk;   // feature chunks [chunk0, chunk0 + n_chunk), n_chunk <= 2 (n32: <= 4)
  const float* a_scale;  // device scalar (power of two) applied to the gradient tensor before the fp16 split
  float* dw;             // [fout_total, 3*fin], column = f*3 + k  (reference layout), accumulated atomically
  int* status;
  // Swapped roles (L~ symmetric:  sum_rows dz (x) T_k(X) = sum_rows T_k(dz) (x) X): `x` is the GRADIENT dz
  // [rows, fin := layer Fout] whose basis the producers build (scaled by a_scale), `t1` = L~ dz as left behind by
  // the backward-data pass (launch_cheb_t1), `g` the layer INPUT [rows(/2), fout_total := layer Fin] (plain tile,
  // unscaled, read at row >> 1 under the virtual unpool).  The accumulator rows are then input features and the
  // columns output channels: dw[o][f*3+k] with o from the gathered side.  No 2-hop halo, no on-chip T1.
  const float* t1;
  int g_unpool;
  int swap;
  int tma;                 // T1 given, consecutive tiles, V % 128 == 0: the own rows of x and t1 arrive by one 2-D TMA box each
  int n32;                 // 1: the hi and lo halves of a T block accumulate into the SAME 32 TMEM columns (three N = 32
                           //    MMAs per K step: g_hi T_hi + g_lo T_hi + g_hi T_lo; the lo half is addressed by starting
                           //    the MN-major descriptor 64 bytes into the 128-byte swizzle row) -> 12 accumulators of 32
                           //    columns = FOUR feature chunks per launch: the plain-side tile is split half as often
  CUtensorMap tm_x, tm_t1;
};

__device__ __forceinline__ uint64_t make_desc_sw128_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // LBO: distance between 64-element MN groups
  d |= (uint64_t)(1024 >> 4) << 32;                  // SBO: distance between 8-row K groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

constexpr int DW_NS = 3;
constexpr int DW_G_BYTES = 4 * A_BLOCK_BYTES;  // dz tile: (hi, lo) x two 64-channel groups

template <int XS>
__global__ void __launch_bounds__(NUM_THREADS2, 1) k_cheb_dw_umma(const __grid_constant__ DwParams p) {
  constexpr uint32_t IDESC = make_idesc_f16(TILE_M, 64) | (1u << 15) | (1u << 16);  // A and B MN-major
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* ring = smem_raw;                      // [DW_NS] T blocks
  unsigned char* gblk = ring + DW_NS * A_BLOCK_BYTES;  // dz tile blocks: hi g0, hi g1, lo g0, lo g1
  float* Xs = reinterpret_cast<float*>(gblk + DW_G_BYTES);
  const bool t1g = (p.t1 != nullptr);
  const size_t xs_stage_floats = (size_t)(t1g ? TILE_M : p.max_h2) * FC;
  float* T1s = Xs + XS * xs_stage_floats;                   // [XS | 1][max_h1][32]
  const size_t t1_stage_floats = (size_t)p.max_h1 * FC;
  unsigned char* meta_s = reinterpret_cast<unsigned char*>(T1s + (t1g ? XS : 1) * t1_stage_floats);
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta_s + 2 * (size_t)p.meta_stride);
  uint64_t* b_t_full = bars;                 // [DW_NS]
  uint64_t* b_t_empty = b_t_full + DW_NS;    // [DW_NS]
  uint64_t* b_x_full = b_t_empty + DW_NS;    // [XS]
  uint64_t* b_x_empty = b_x_full + XS;       // [XS]
  uint64_t* b_m_full = b_x_empty + XS;       // [2]
  uint64_t* b_m_empty = b_m_full + 2;        // [2]
  uint64_t* b_g_full = b_m_empty + 2;        // [1]
  uint64_t* b_g_empty = b_g_full + 1;        // [1]
  uint64_t* b_done = b_g_empty + 1;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_done + 1);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_chunk = p.n_chunk;
  const int n_use = 3 * n_chunk;
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < DW_NS; ++s) {
      mbar_init(smem_u32(b_t_full + s), W_PROD);
      mbar_init(smem_u32(b_t_empty + s), 1);
    }
    for (int s = 0; s < XS; ++s) {
      mbar_init(smem_u32(b_x_full + s), N_XLOAD * 32 + 1);  // the loader threads' cp.async arrivals + one expect_tx / plain arrival
      mbar_init(smem_u32(b_x_empty + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(b_m_full + s), 1);
      mbar_init(smem_u32(b_m_empty + s), 1);
    }
    mbar_init(smem_u32(b_g_full), W_PROD);
    mbar_init(smem_u32(b_g_empty), 1);
    mbar_init(smem_u32(b_done), 1);
    *abort_flag = (smem_u32(ring) & 1023u) ? 1 : 0;
    if (*abort_flag) mbar_timeout(abort_flag, p.status, 100);
    fence_barrier_init();
  }
  const float a_scale = p.a_scale ? *p.a_scale : 1.f;
  if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= W_XLOAD && warp < W_XLOAD + N_XLOAD) {
    // ------------------------------------------------------------ halo loaders + tile metadata (as in the forward)
    const int lt = tid - W_XLOAD * 32;
    const int q = lt & 7, rg = lt >> 3;
    auto fetch_meta = [&](int it2) {
      const int pat = (blockIdx.x + it2 * gridDim.x) % p.P;
      const int m2 = it2 & 1;
      mbar_wait_relaxed(smem_u32(b_m_empty + m2), ((it2 >> 1) & 1) ^ 1, abort_flag, p.status, 21);
      const int mbytes = p.meta_bytes[pat];
      mbar_arrive_expect_tx(smem_u32(b_m_full + m2), mbytes);
      bulk_g2s(smem_u32(meta_s + (size_t)m2 * p.meta_stride), p.meta + (size_t)pat * p.meta_stride, mbytes,
               smem_u32(b_m_full + m2));
    };
    if (lt == 0 && my_tiles > 0) fetch_meta(0);
    uint32_t g = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int b = tile / p.P;
      const int m = it & 1;
      mbar_wait_relaxed(smem_u32(b_m_full + m), (it >> 1) & 1, abort_flag, p.status, 22);
      const unsigned char* mb = meta_s + (size_t)m * p.meta_stride;
      const TileHeader* hdr = reinterpret_cast<const TileHeader*>(mb);
      const int h2 = t1g ? TILE_M : hdr->h2;  // T1 given: only the tile's own rows of x are staged ...
      const int h1 = hdr->h1;
      const int* halo = reinterpret_cast<const int*>(mb + hdr->off_halo);
      const long long mesh_row0 = (long long)b * p.V;
      if (it + 1 < my_tiles) {
        // the producers read the NEXT tile's plain-side rows straight from global memory at the start of that tile:
        // pull them into L2 now (consecutive tiles only; index-list tiles would need the next blob first)
        const int tile2 = blockIdx.x + (it + 1) * gridDim.x;
        const long long r2 = (long long)(tile2 / p.P) * p.V + (long long)(tile2 % p.P) * TILE_M;
        const int lpr = (p.m_cols * 4 + 127) >> 7;  // 128-byte lines per row of this launch's channel slice
        for (int j = lt; j < TILE_M * lpr; j += N_XLOAD * 32) {
          const long long rr = r2 + j / lpr;
          if (rr < (long long)(tile2 / p.P + 1) * p.V)
            prefetch_l2(p.g + (p.g_unpool ? (rr >> 1) : rr) * p.fout_total + p.m_off + (j % lpr) * 32);
        }
      }
      for (int c = 0; c < n_chunk; ++c, ++g) {
        const int xs = g % XS;
        mbar_wait_relaxed(smem_u32(b_x_empty + xs), ((g / XS) & 1) ^ 1, abort_flag, p.status, 23);
        const uint32_t dst0 = smem_u32(Xs + xs * xs_stage_floats) + q * 16;
        const float* src0 = p.x + (p.chunk0 + c) * FC + q * 4;
        const uint32_t xbar = smem_u32(b_x_full + xs);
        if (lt == 0) {
          if (p.tma) {  // own rows of x and t1: one TMA box each, landing asynchronously
            const int own0 = tile * TILE_M;  // V is a multiple of 128: tiles never straddle meshes
            mbar_arrive_expect_tx(xbar, 2 * TILE_M * 128);
            tma_load_2d(smem_u32(Xs + xs * xs_stage_floats), &p.tm_x, (p.chunk0 + c) * FC, own0, xbar);
            tma_load_2d(smem_u32(T1s + xs * t1_stage_floats), &p.tm_t1, (p.chunk0 + c) * FC, own0, xbar);
          } else {
            mbar_arrive(xbar);
          }
        }
        if (!p.tma)
        for (int i = rg; i < h2; i += 8) {
          const int v = halo[i];
          if (v >= 0) {
            long long r = mesh_row0 + v;
            if (p.in_unpool) r >>= 1;
            cp_async16(dst0 + i * 128, src0 + r * p.fin);
          } else {
            sts_f4(dst0 + i * 128, make_float4(0.f, 0.f, 0.f, 0.f));
          }
        }
        if (t1g) {  // ... plus the T1 rows of the tile and its 1-hop halo
          const uint32_t dst1 = smem_u32(T1s + xs * t1_stage_floats) + q * 16;
          const float* src1 = p.t1 + (p.chunk0 + c) * FC + q * 4;
          for (int i = (p.tma ? TILE_M : 0) + rg; i < h1; i += 8) {  // (TMA: only the halo rows are left)
            const int v = halo[i];
            if (v >= 0)
              cp_async16(dst1 + i * 128, src1 + (mesh_row0 + v) * p.fin);
            else
              sts_f4(dst1 + i * 128, make_float4(0.f, 0.f, 0.f, 0.f));
          }
        }
        cp_async_arrive_noinc(xbar);
        if (c == 0 && lt == 0 && it + 1 < my_tiles) fetch_meta(it + 1);
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t ucnt = 0;
      const uint32_t g_hi = smem_u32(gblk), g_lo = g_hi + 2 * A_BLOCK_BYTES;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(smem_u32(b_g_full), it & 1, abort_flag, p.status, 24);
        tc_fence_after();
        for (int u = 0; u < n_use; ++u, ++ucnt) {
          const int s = ucnt % DW_NS;
          mbar_wait(smem_u32(b_t_full + s), (ucnt / DW_NS) & 1, abort_flag, p.status, 25);
          tc_fence_after();
          const uint64_t dt = make_desc_sw128_mn(smem_u32(ring + s * A_BLOCK_BYTES), A_BLOCK_BYTES);
          const uint64_t dh = make_desc_sw128_mn(g_hi, A_BLOCK_BYTES), dl = make_desc_sw128_mn(g_lo, A_BLOCK_BYTES);
          if (p.n32) {
            constexpr uint32_t IDESC32 = make_idesc_f16(TILE_M, 32) | (1u << 15) | (1u << 16);
            const uint32_t d_tmem = tmem_base + (uint32_t)(u * 32);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              umma_f16(d_tmem, dh + ks * 128, dt + ks * 128, IDESC32, (it > 0 || ks > 0) ? 1u : 0u);  // g_hi T_hi
              umma_f16(d_tmem, dl + ks * 128, dt + ks * 128, IDESC32, 1u);                              // g_lo T_hi
              umma_f16(d_tmem, dh + ks * 128, dt + ks * 128 + 4, IDESC32, 1u);                          // g_hi T_lo (+64 B)
            }
          } else {
            const uint32_t d_tmem = tmem_base + (uint32_t)(u * 64);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {  // 16 mesh rows per K step = 2048 bytes = +128 in the address field
              umma_f16(d_tmem, dh + ks * 128, dt + ks * 128, IDESC, (it > 0 || ks > 0) ? 1u : 0u);
              umma_f16(d_tmem, dl + ks * 128, dt + ks * 128, IDESC, 1u);
            }
          }
          umma_commit(smem_u32(b_t_empty + s));
        }
        umma_commit(smem_u32(b_g_empty));  // the dz blocks may be overwritten once these MMAs have read them
      }
      umma_commit(smem_u32(b_done));
    }
  } else if (warp >= W_EPI0) {
    // ------------------------------------------------------------ final reduction: TMEM -> atomicAdd into dW
    const int lane_base = (warp & 3) * 32;
    const int o_local = lane_base + lane;
    mbar_wait_relaxed(smem_u32(b_done), 0, abort_flag, p.status, 26);
    tc_fence_after();
    const float inv = 1.f / a_scale;
    const bool valid = (o_local < p.m_cols) && (my_tiles > 0);
    float* drow = p.dw + (size_t)(p.m_off + o_local) * 3 * p.fin;
    for (int u = 0; u < n_use; ++u) {
      const int c = u / 3, k = u - 3 * c;
      uint32_t hi[32], lo[32];
      if (p.n32) {
        tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(u * 32), hi);
#pragma unroll
        for (int j = 0; j < 32; ++j) lo[j] = 0u;
      } else {
        tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(u * 64), hi);
        tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(u * 64 + 32), lo);
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int f = (p.chunk0 + c) * FC + j;  // channel of the gathered side
          const float v = (__uint_as_float(hi[j]) + __uint_as_float(lo[j])) * inv;
          if (p.swap)  // accumulator row = input feature (plain side), column = output channel
            atomicAdd(p.dw + (size_t)f * 3 * p.fout_total + (size_t)(p.m_off + o_local) * 3 + k, v);
          else
            atomicAdd(drow + f * 3 + k, v);
        }
      }
    }
  } else if (warp < W_PROD) {
    // ------------------------------------------------------------ producers
    const int q = tid & 7, rg = tid >> 3;
    const uint32_t t1s_a = smem_u32(T1s), ring_a = smem_u32(ring), g_a = smem_u32(gblk);
    uint32_t ucnt = 0, gcnt = 0;
    constexpr int T1_ROWS = 4;
    uint32_t t1_row[T1_ROWS], t1_e[T1_ROWS];
    uint32_t row0 = 0, row1 = 0, r0e = 0, r1e = 0, ent_a = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int b = tile / p.P, pat = tile - b * p.P;
      const int m = it & 1;
      // The plain-side tile goes global -> registers -> fp16 blocks.  Its loads are issued FIRST, so that their
      // latency overlaps the waits below (metadata, and above all g_empty: the previous tile's MMAs still read the
      // single-buffered plain blocks) instead of following them.
      float4 pv[2][4];
      {
        const int n_rows = min(TILE_M, p.V - pat * TILE_M);
        const long long r_base = (long long)b * p.V + (long long)pat * TILE_M;
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          const int i = ps * 64 + rg;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int col = q * 4 + 32 * (jj ^ (rg & 1));
            pv[ps][jj] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n_rows && col < p.m_cols) {
              const long long rr = p.g_unpool ? ((r_base + i) >> 1) : (r_base + i);
              pv[ps][jj] = __ldg(reinterpret_cast<const float4*>(p.g + rr * p.fout_total + p.m_off + col));
            }
          }
        }
      }
      mbar_wait(smem_u32(b_m_full + m), (it >> 1) & 1, abort_flag, p.status, 27);
      {
        const unsigned char* mb = meta_s + (size_t)m * p.meta_stride;
        const TileHeader* hdr = reinterpret_cast<const TileHeader*>(mb);
        const uint32_t mb_a = smem_u32(mb);
        const uint32_t rp_a = mb_a + hdr->off_rp, ord1_a = mb_a + hdr->off_ord1, ord2_a = mb_a + hdr->off_ord2;
        ent_a = mb_a + hdr->off_ent;
        const int h1 = t1g ? 0 : hdr->h1;  // the trimmed metadata has no T1 row order
#pragma unroll
        for (int t = 0; t < T1_ROWS; ++t) {
          const int j = rg + 64 * t;
          t1_row[t] = 0xFFFFu;
          if (j < h1) {
            const uint32_t i = lds_u16(ord1_a + 2 * j);
            t1_row[t] = i;
            t1_e[t] = lds_u16(rp_a + 2 * i) | (lds_u16(rp_a + 2 * i + 2) << 16);
          }
        }
        row0 = lds_u16(ord2_a + 2 * rg);
        row1 = lds_u16(ord2_a + 2 * (64 + rg));
        r0e = lds_u16(rp_a + 2 * row0) | (lds_u16(rp_a + 2 * row0 + 2) << 16);
        r1e = lds_u16(rp_a + 2 * row1) | (lds_u16(rp_a + 2 * row1 + 2) << 16);
      }
      // plain-side tile (dz, or the layer input in swapped mode) -> (hi, lo) fp16 blocks, MN-major [row][channel]
      mbar_wait(smem_u32(b_g_empty), (it & 1) ^ 1, abort_flag, p.status, 28);
      {
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          const int i = ps * 64 + rg;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            // odd row groups take the 32-channel pieces in the order 1,0,3,2: the two rows of a half-warp then store
            // into different 64-byte bank halves (their swizzle bits agree: consecutive rows)
            const int col = q * 4 + 32 * (jj ^ (rg & 1));  // channel inside this launch's 128-channel slice
            float4 v = pv[ps][jj];
            if (!p.swap) {  // legacy roles: this side is the gradient
              v.x *= a_scale; v.y *= a_scale; v.z *= a_scale; v.w *= a_scale;
            }
            uint2 hi, lo;
            split4(v, hi, lo);
            const uint32_t off = (uint32_t)(col >> 6) * A_BLOCK_BYTES + sw128_off(i, (col & 63) >> 3) + ((col >> 2) & 1) * 8;
            sts_u2(g_a + off, hi);  // (rows i = ps*64 + rg: 8 lanes x 8 B = 64 B per row and store, hi and lo blocks apart)
            sts_u2(g_a + 2 * A_BLOCK_BYTES + off, lo);
          }
        }
        fence_async_proxy();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(smem_u32(b_g_full));
      }
      for (int c = 0; c < n_chunk; ++c, ++gcnt) {
        const int xs = gcnt % XS;
        mbar_wait(smem_u32(b_x_full + xs), (gcnt / XS) & 1, abort_flag, p.status, 29);
        const uint32_t xs_q = smem_u32(Xs + xs * xs_stage_floats) + q * 16;
        const uint32_t t1s_q = t1s_a + (t1g ? (uint32_t)(xs * t1_stage_floats * 4) : 0u) + q * 16;
        if (!t1g) {
#pragma unroll
          for (int t = 0; t < T1_ROWS; ++t) {
            if (t1_row[t] != 0xFFFFu)
              sts_f4(t1s_q + t1_row[t] * 128, gather_row4(ent_a, t1_e[t] & 0xFFFFu, t1_e[t] >> 16, xs_q));
          }
          producer_barrier();
        }
        float4 tv[3][2];
        {
          const float4 g0 = gather_row4(ent_a, r0e & 0xFFFFu, r0e >> 16, t1s_q);
          const float4 g1 = gather_row4(ent_a, r1e & 0xFFFFu, r1e >> 16, t1s_q);
          tv[0][0] = lds_f4(xs_q + row0 * 128);
          tv[0][1] = lds_f4(xs_q + row1 * 128);
          tv[1][0] = lds_f4(t1s_q + row0 * 128);
          tv[1][1] = lds_f4(t1s_q + row1 * 128);
          const float4 a = tv[0][0], c2 = tv[0][1];
          tv[2][0] = cheb_t2(g0, a);
          tv[2][1] = cheb_t2(g1, c2);
          if (p.swap) {  // the gathered side is the gradient: into fp16's range before the split
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
              for (int ps = 0; ps < 2; ++ps) {
                tv[k][ps].x *= a_scale; tv[k][ps].y *= a_scale; tv[k][ps].z *= a_scale; tv[k][ps].w *= a_scale;
              }
          }
        }
        const uint32_t u0 = ucnt;
#pragma unroll
        for (int k = 0; k < 3; ++k, ++ucnt) {
          const int s = ucnt % DW_NS;
          mbar_wait(smem_u32(b_t_empty + s), ((ucnt / DW_NS) & 1) ^ 1, abort_flag, p.status, 30);
          const uint32_t blk = ring_a + s * A_BLOCK_BYTES + (q & 1) * 8;
#pragma unroll
          for (int ps = 0; ps < 2; ++ps) {
            const uint32_t i = ps ? row1 : row0;
            uint2 hi, lo;
            split4(tv[k][ps], hi, lo);
            const uint32_t a_hi = blk + sw128_off(i, q >> 1), a_lo = blk + sw128_off(i, 4 + (q >> 1));
            const bool odd = (rg & 1) != 0;  // see emit() of the conv kernel
            sts_u2(odd ? a_lo : a_hi, odd ? lo : hi);
            sts_u2(odd ? a_hi : a_lo, odd ? hi : lo);
          }
        }
        fence_async_proxy();
        __syncwarp();
        if ((tid & 31) == 0) {
#pragma unroll
          for (int k = 0; k < 3; ++k) mbar_arrive(smem_u32(b_t_full + (u0 + k) % DW_NS));
        }
        producer_barrier();
        if (tid == 0) {
          mbar_arrive(smem_u32(b_x_empty + xs));
          if (c == n_chunk - 1) mbar_arrive(smem_u32(b_m_empty + m));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================
// k_cheb_t1 — T1 = L~ X for every row, written once to HBM (fp32, logical rows).  With it the conv kernel only
// needs the tile's own X rows and the T1 rows of its 1-hop halo: the 35 % of the first sparse product that the
// fused kernel spends re-computing T1 on halo rows disappears, and so does its 2-hop X staging.
// Simple kernel: one CTA (512 threads, two per SM) per 128-row tile; the 1-hop halo of X is staged chunk by chunk
// through a 4-deep cp.async ring (one barrier per chunk), the gather runs out of shared memory.
// =====================================================================================
struct T1Params {
  const float* x;
  int in_unpool;
  int V, P, fin;
  const unsigned char* meta;
  const int* meta_bytes;
  int meta_stride, max_h1;
  float* t1;
};

// T1_STAGES = cp.async ring depth (chunks in flight per CTA): 4 when two CTAs of that size fit an SM, else 3 or 2
template <int T1_STAGES>
__global__ void __launch_bounds__(512, 2) k_cheb_t1(const T1Params p) {
  extern __shared__ __align__(16) unsigned char smem_t1[];
  unsigned char* meta_s = smem_t1;
  float* Xs = reinterpret_cast<float*>(smem_t1 + p.meta_stride);  // [T1_STAGES][max_h1][32]
  const size_t stage_floats = (size_t)p.max_h1 * FC;
  const int tid = threadIdx.x, q = tid & 7, rg = tid >> 3;
  const int tile = blockIdx.x;
  const int b = tile / p.P, pat = tile - b * p.P;
  const long long mesh_row0 = (long long)b * p.V;
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.meta + (size_t)pat * p.meta_stride);
    uint4* dst = reinterpret_cast<uint4*>(meta_s);
    const int n16 = p.meta_bytes[pat] >> 4;
    for (int i = tid; i < n16; i += 512) dst[i] = src[i];
  }
  __syncthreads();
  const TileHeader* hdr = reinterpret_cast<const TileHeader*>(meta_s);
  const int h1 = hdr->h1;
  const int* halo = reinterpret_cast<const int*>(meta_s + hdr->off_halo);
  const uint32_t mb_a = smem_u32(meta_s);
  const uint32_t rp_a = mb_a + hdr->off_rp, ent_a = mb_a + hdr->off_ent, ord2_a = mb_a + hdr->off_ord2;
  const int n_chunk = p.fin / FC;
  // source rows of the staged slots this thread copies (same for every chunk): slots rg, rg + 64, ...
  long long srow[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = rg + 64 * u;
    const int v = (i < h1) ? halo[i] : -2;
    long long r = mesh_row0 + v;
    if (p.in_unpool) r >>= 1;
    srow[u] = (v >= 0) ? r * p.fin : (long long)v;  // -1: empty slot (zero-filled), -2: beyond the halo
  }
  auto stage = [&](int c) {
    if (c < n_chunk) {
      const uint32_t dst0 = smem_u32(Xs + (c % T1_STAGES) * stage_floats) + q * 16;
      const float* src0 = p.x + c * FC + q * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = rg + 64 * u;
        if (srow[u] >= 0)
          cp_async16(dst0 + i * 128, src0 + srow[u]);
        else if (srow[u] == -1)
          sts_f4(dst0 + i * 128, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // (possibly empty: keeps the group count uniform)
  };
  // rows this thread produces, taken in length-sorted order: the four rows of a warp then have similar lengths
  int vtx[2];
  uint32_t re[2];
#pragma unroll
  for (int ps = 0; ps < 2; ++ps) {
    const int i = lds_u16(ord2_a + 2 * (ps * 64 + rg));
    vtx[ps] = halo[i];
    re[ps] = lds_u16(rp_a + 2 * i) | (lds_u16(rp_a + 2 * i + 2) << 16);
  }
#pragma unroll
  for (int c = 0; c < T1_STAGES - 1; ++c) stage(c);
  for (int c = 0; c < n_chunk; ++c) {
    asm volatile("cp.async.wait_group %0;" ::"n"(T1_STAGES - 2) : "memory");  // chunk c has landed (this thread's part)
    __syncthreads();  // ... and everybody's; also: everybody is done with chunk c - 1, whose stage is refilled next
    stage(c + T1_STAGES - 1);
    const uint32_t xs_q = smem_u32(Xs + (c % T1_STAGES) * stage_floats) + q * 16;
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      if (vtx[ps] >= 0) {
        const float4 acc = gather_row4(ent_a, re[ps] & 0xFFFFu, re[ps] >> 16, xs_q);
        *reinterpret_cast<float4*>(p.t1 + (mesh_row0 + vtx[ps]) * p.fin + c * FC + q * 4) = acc;
      }
    }
  }
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

// debug: P2M_UMMA_TMA=0 stages every row with cp.async (A/B measurements of the TMA own-row loads)
// debug: P2M_DW_N32=0 keeps separate hi / lo accumulator columns in the swapped dW kernel (two chunks per launch)
const bool g_dw_n32 = [] { const char* e = std::getenv("P2M_DW_N32"); return !(e && e[0] == '0'); }();
const bool g_umma_tma = [] { const char* e = std::getenv("P2M_UMMA_TMA"); return !(e && e[0] == '0'); }();

// mode: 0 = fused (X with its 2-hop halo staged, T1 recomputed on chip), 1 = T1 given, 2 = plain GEMM
inline int epi_stage_bytes(int N) { return 4 * 32 * (N == 256 ? 16 : 32) * 4; }  // per-warp transpose staging
size_t smem_bytes_dims(int N, int NS, int XS, int max_h1, int max_h2, int meta_stride, int mode) {
  const size_t fixed = 1024 + (size_t)NS * (A_BLOCK_BYTES + N * 128) + 8 * (2 * NS + 2 * XS + 8) + 16 +
                       2 * (size_t)N * 4 + 16 + 128 + (size_t)epi_stage_bytes(N) + 4 * 32 * 4 +
                       (N == 64 ? 64 * 12 * 4 : 0);
  if (mode == 1) return fixed + (size_t)XS * TILE_M * FC * 4 + (size_t)XS * max_h1 * FC * 4 + 2 * (size_t)meta_stride;
  if (mode == 2) return fixed + (size_t)XS * TILE_M * FC * 4 + 2 * (size_t)meta_stride;
  return fixed + (size_t)XS * max_h2 * FC * 4 + (size_t)max_h1 * FC * 4 + 2 * (size_t)meta_stride;
}
size_t smem_bytes_for(int N, int NS, int XS, const DevLevel& g, int mode = 0) {
  return smem_bytes_dims(N, NS, XS, g.max_h1, g.max_h2, mode ? g.meta1_stride : g.meta_stride, mode);
}
// the same for an explicit launch: the level's consecutive tiles, or the tile family the caller selected
size_t smem_bytes_args(int N, int NS, int XS, const UmmaConvArgs& a) {
  const int mode = a.plain ? 2 : (a.t1 != nullptr ? 1 : 0);
  if (a.tiles != nullptr) return smem_bytes_dims(N, NS, XS, a.tiles->max_h1, a.tiles->max_h1, a.tiles->stride, mode);
  return smem_bytes_for(N, NS, XS, *a.g, mode);
}
constexpr size_t SMEM_LIMIT = 227 * 1024;
inline int ring_stages(int N) { return N == 256 ? 2 : 3; }
// X staging depth: 2 (prefetch the next chunk's halo during the current chunk) when it fits, else 1
inline int x_stages(int N, const DevLevel& g) {
  return smem_bytes_for(N, ring_stages(N), 2, g) <= SMEM_LIMIT ? 2 : 1;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda)
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<TmapEncodeFn>(f);
  }();
  return fn;
}
// [rows, fin] fp32 row-major matrix, box = 32 features x box_rows rows, dense (unswizzled) 128-byte rows in shared memory
bool make_row_tmap(CUtensorMap* tm, const float* base, long long rows, int fin, int box_rows) {
  TmapEncodeFn enc = tmap_encoder();
  if (enc == nullptr || (reinterpret_cast<uintptr_t>(base) & 15u) != 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)fin, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)fin * 4};
  const cuuint32_t box[2] = {(cuuint32_t)FC, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The conv kernel's setmaxnreg split is balanced for a launch allocation of REGS_LAUNCH registers per thread: a build
// that ends up with another count would leave the epilogue's setmaxnreg.inc spinning on an empty pool.
template <int N, int NS, int XS, int RM, int MODE>
int check_launch_regs() {
  static int state = 0;  // per instantiation; racing first calls all compute the same value
  if (state == 0) {
    cudaFuncAttributes fa;
    P2M_CUDA_OK(cudaFuncGetAttributes(&fa, k_cheb_conv_umma<N, NS, XS, RM, MODE>));
    state = (fa.numRegs == REGS_LAUNCH) ? 1 : -1;
  }
  if (state < 0) {
    set_error("k_cheb_conv_umma was compiled with a register count other than the one its setmaxnreg split assumes");
    return P2M_ERR_CUDA;
  }
  return P2M_OK;
}

template <int N, int NS, int XS, int RM = 0, int MODE = 0>
int launch_cfg(const UmmaConvArgs& a, int* status, const float* zero_row, int sm_count, cudaStream_t s) {
  if (MODE == 0 && a.t1 != nullptr && !a.plain)  // the production configuration has its own instantiation
    return launch_cfg<N, NS, XS, RM, 1>(a, status, zero_row, sm_count, s);
  const DevLevel& g = *a.g;
  const int mode = a.plain ? 2 : (a.t1 != nullptr ? 1 : 0);
  const size_t smem = smem_bytes_args(N, NS, XS, a);
  auto kern = k_cheb_conv_umma<N, NS, XS, RM, MODE>;
  P2M_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  P2M_TRY((check_launch_regs<N, NS, XS, RM, MODE>()));
  KParams p;
  p.x = a.x;
  p.in_unpool = a.in_unpool;
  p.V = g.V;
  p.P = g.n_pattern;
  p.fin = a.fin;
  p.meta = mode ? g.tile_meta1 : g.tile_meta;
  p.meta_bytes = mode ? g.tile_meta1_bytes : g.tile_meta_bytes;
  p.meta_stride = mode ? g.meta1_stride : g.meta_stride;
  p.max_h1 = g.max_h1;
  p.max_h2 = g.max_h2;
  p.own_table = 0;
  if (a.tiles != nullptr) {  // index-list tiles (mode 1 / 2 only, checked by the caller)
    p.P = a.tiles->n_pattern;
    p.meta = a.tiles->meta;
    p.meta_bytes = a.tiles->bytes;
    p.meta_stride = a.tiles->stride;
    p.max_h1 = a.tiles->max_h1;
    p.max_h2 = a.tiles->max_h1;
    p.own_table = 1;
  }
  p.n_tiles = a.batch * p.P;
  p.wpack = static_cast<const unsigned char*>(a.wpack);
  p.apack = nullptr;
  p.wslice_bytes = 0;
  p.zero_row = zero_row;
  p.ep = to_dev(a.ep);
  p.res_identity = (a.ep.res != nullptr && a.ep.res_F == a.fout) ? 1 : 0;
  p.y = a.y;
  p.t1 = a.t1;
  p.plain = a.plain;
  p.a_scale = a.a_scale;
  p.ldy = a.ldy > 0 ? a.ldy : a.fout;
  p.y_col0 = a.y_col0;
  p.status = status;
  p.trace = a.trace;
  p.head_wt = (N == 64) ? a.head_wt : nullptr;
  p.head_z = (N == 64) ? a.head_z : nullptr;
  p.tma = 0;
  std::memset(&p.tm_x, 0, sizeof(p.tm_x));
  std::memset(&p.tm_t1, 0, sizeof(p.tm_t1));
  if ((a.t1 != nullptr || a.plain) && a.tiles == nullptr && g.V % TILE_M == 0 && g_umma_tma) {
    const long long rows = (long long)a.batch * g.V;
    bool ok = make_row_tmap(&p.tm_x, a.x, a.in_unpool ? rows / 2 : rows, a.fin, a.in_unpool ? TILE_M / 2 : TILE_M);
    if (ok && a.t1 != nullptr) ok = make_row_tmap(&p.tm_t1, a.t1, rows, a.fin, TILE_M);
    p.tma = ok ? 1 : 0;
  }
  const int grid = std::min(p.n_tiles, sm_count);
  kern<<<grid, NUM_THREADS2, smem, s>>>(p);
  P2M_LAUNCH_OK();
  return P2M_OK;
}

template <int N>
int launch_n(const UmmaConvArgs& a, int* status, const float* zero_row, int sm_count, cudaStream_t s) {
  constexpr int NS = (N == 256) ? 2 : 3;
  if (N == 128 && a.ep.res != nullptr && a.ep.res_F == 2 * N && (a.t1 != nullptr || a.plain) && a.head_z == nullptr &&
      smem_bytes_args(N, NS, 2, a) <= SMEM_LIMIT)
    return launch_cfg<N, NS, 2, (N == 128 ? 1 : 0)>(a, status, zero_row, sm_count, s);  // pair-mean residual (256 -> 128)
  if (a.t1 != nullptr || a.plain) {
    if (smem_bytes_args(N, NS, 2, a) <= SMEM_LIMIT) return launch_cfg<N, NS, 2>(a, status, zero_row, sm_count, s);
    if (smem_bytes_args(N, NS, 1, a) > SMEM_LIMIT) {
      set_error("umma_conv: tile family does not fit shared memory");
      return P2M_ERR_INVALID;
    }
    return launch_cfg<N, NS, 1>(a, status, zero_row, sm_count, s);
  }
  if (x_stages(N, *a.g) == 2) return launch_cfg<N, NS, 2>(a, status, zero_row, sm_count, s);
  return launch_cfg<N, NS, 1>(a, status, zero_row, sm_count, s);
}

}  // namespace

// =====================================================================================
// host side
// =====================================================================================
namespace {
// The producers store a tile row's fp16 hi part into one 64-byte half of its 128-byte A-block row and the lo part
// into the other; which half is which follows the row's swizzle bit (row & 4).  A warp stores four rows per
// instruction (row groups 4w .. 4w+3 of the order below, odd groups lo first) and 64-bit shared stores are served per
// half-warp, i.e. per PAIR of rows (positions 0,1 and 2,3 of an aligned group of four): a pair is free of bank
// replays iff its two 64-byte pieces fall into different halves — with the odd group storing the other part first,
// iff both rows of the pair have the SAME half class.  Re-deal a length-sorted order accordingly: [c0 c0 c1 c1] per
// group, taken in sorted order (a group's rows still have similar lengths).  Pure re-ordering of which thread
// produces which row: results are unchanged.
void balance_store_halves(std::vector<unsigned short>* ord) {
  std::vector<unsigned short> q0, q1;
  for (unsigned short r : *ord) ((r & 4) ? q1 : q0).push_back(r);
  if (q0.size() != q1.size() || (q0.size() & 1)) return;
  size_t o = 0;
  for (size_t g = 0; g + 1 < q0.size(); g += 2) {
    (*ord)[o++] = q0[g];
    (*ord)[o++] = q0[g + 1];
    (*ord)[o++] = q1[g];
    (*ord)[o++] = q1[g + 1];
  }
}

// Trimmed blob of one tile whose 128 own rows are given by an index list (-1 = empty slot): own rows, their 1-hop
// halo, the CSR of the own rows with staged-row slots as columns, the own rows in length-sorted order.
bool make_indexed_blob(const std::vector<int>& own, const int* rowptr, const int* colidx, const float* val,
                       std::vector<int>* slot_of, std::vector<unsigned char>* blob, int* h1_out) {
  std::vector<int> halo(own);
  int n_rows = 0;
  for (int i = 0; i < TILE_M; ++i)
    if (own[i] >= 0) {
      (*slot_of)[own[i]] = i;
      ++n_rows;
    }
  for (int i = 0; i < TILE_M; ++i) {
    if (own[i] < 0) continue;
    for (int e = rowptr[own[i]]; e < rowptr[own[i] + 1]; ++e) {
      const int v = colidx[e];
      if ((*slot_of)[v] < 0) {
        (*slot_of)[v] = (int)halo.size();
        halo.push_back(v);
      }
    }
  }
  const int h1 = (int)halo.size();
  std::vector<unsigned short> rp(TILE_M + 1, 0);
  std::vector<unsigned int> ent;
  for (int i = 0; i < TILE_M; ++i) {
    if (own[i] >= 0)
      for (int e = rowptr[own[i]]; e < rowptr[own[i] + 1]; ++e) {
        unsigned int bits;
        std::memcpy(&bits, &val[e], 4);
        ent.push_back((unsigned int)(*slot_of)[colidx[e]] * 128u);
        ent.push_back(bits);
      }
    rp[i + 1] = (unsigned short)(ent.size() / 2);
  }
  for (int v : halo)
    if (v >= 0) (*slot_of)[v] = -1;
  const int nnz = (int)(ent.size() / 2);
  if (h1 > 512 || nnz > 65535) return false;
  std::vector<unsigned short> ord2(TILE_M);
  for (int i = 0; i < TILE_M; ++i) ord2[i] = (unsigned short)i;
  std::stable_sort(ord2.begin(), ord2.end(), [&](unsigned short a, unsigned short b2) {
    return (int)rp[a + 1] - (int)rp[a] > (int)rp[b2 + 1] - (int)rp[b2];
  });
  balance_store_halves(&ord2);
  TileHeader t{};
  t.n_rows = n_rows;
  t.h1 = h1;
  t.h2 = h1;
  t.nnz = nnz;
  int o1 = 64;
  t.off_halo = o1; o1 += up16(h1 * 4);
  t.off_rp = o1;   o1 += up16((TILE_M + 1) * 2);
  t.off_ent = o1;  o1 += up16(nnz * 8);
  t.off_ord2 = o1; o1 += up16(TILE_M * 2);
  t.off_ord1 = t.off_ord2;
  t.bytes = o1;
  blob->assign(o1, 0);
  std::memcpy(blob->data(), &t, sizeof(t));
  std::memcpy(blob->data() + t.off_halo, halo.data(), (size_t)h1 * 4);
  std::memcpy(blob->data() + t.off_rp, rp.data(), (TILE_M + 1) * 2);
  if (nnz) std::memcpy(blob->data() + t.off_ent, ent.data(), (size_t)nnz * 8);
  std::memcpy(blob->data() + t.off_ord2, ord2.data(), TILE_M * 2);
  *h1_out = h1;
  return true;
}

// Tiles of 128 consecutive entries of `rows` (ascending vertex ids), uploaded as one TileSet.
int build_tileset(const std::vector<int>& rows, const int* rowptr, const int* colidx, const float* val, int V,
                  TileSet* ts, std::vector<void*>* owned) {
  const int P = ((int)rows.size() + TILE_M - 1) / TILE_M;
  std::vector<std::vector<unsigned char>> blobs(P);
  std::vector<int> slot_of(V, -1);
  int stride = 0, max_h1 = 0;
  for (int pt = 0; pt < P; ++pt) {
    std::vector<int> own(TILE_M, -1);
    for (int i = 0; i < TILE_M && pt * TILE_M + i < (int)rows.size(); ++i) own[i] = rows[pt * TILE_M + i];
    int h1 = 0;
    if (!make_indexed_blob(own, rowptr, colidx, val, &slot_of, &blobs[pt], &h1)) return P2M_ERR_INVALID;
    stride = std::max(stride, (int)blobs[pt].size());
    max_h1 = std::max(max_h1, h1);
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
  ts->meta = d_meta;
  ts->bytes = d_bytes;
  ts->stride = stride;
  ts->n_pattern = P;
  ts->max_h1 = max_h1;
  return P2M_OK;
}
}  // namespace

int build_index_tiles(const std::vector<int>& rows, const int* rowptr, const int* colidx, const float* val, int V,
                      TileSet* ts, std::vector<void*>* owned) {
  return build_tileset(rows, rowptr, colidx, val, V, ts, owned);
}

int build_umma_level_meta(const int* rowptr, const int* colidx, const float* val, int V, DevLevel* out,
                          std::vector<void*>* owned) {
  const int P = (V + TILE_M - 1) / TILE_M;
  std::vector<std::vector<unsigned char>> blobs(P), blobs1(P);
  int max_h1 = 0, max_h2 = 0, stride = 0, stride1 = 0;
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
    // One local CSR serves both products: row i < h1 lists (staged-row slot, value) of vertex halo[i].
    // T1 rows read X slots (< h2); the T2 pass only walks the 128 tile rows, whose columns are < h1,
    // i.e. valid T1 slots (slot numbering of X and T1 coincides below h1).
    std::vector<unsigned short> rp(h1 + 1, 0);
    std::vector<unsigned int> ent;  // pairs {slot*128, float bits}
    for (int i = 0; i < h1; ++i) {
      if (halo[i] >= 0)
        for (int e = rowptr[halo[i]]; e < rowptr[halo[i] + 1]; ++e) {
          unsigned int bits;
          std::memcpy(&bits, &val[e], 4);
          ent.push_back((unsigned int)slot_of[colidx[e]] * 128u);
          ent.push_back(bits);
        }
      rp[i + 1] = (unsigned short)(ent.size() / 2);
    }
    const int nnz = (int)(ent.size() / 2);
    if (nnz > 65535) {
      set_error("umma meta: too many entries in a tile");
      return P2M_ERR_INVALID;
    }
    auto row_len = [&](int i) { return (int)rp[i + 1] - (int)rp[i]; };
    std::vector<unsigned short> ord1(h1), ord2(TILE_M);
    for (int i = 0; i < h1; ++i) ord1[i] = (unsigned short)i;
    for (int i = 0; i < TILE_M; ++i) ord2[i] = (unsigned short)i;
    std::stable_sort(ord1.begin(), ord1.end(), [&](unsigned short a, unsigned short b2) { return row_len(a) > row_len(b2); });
    std::stable_sort(ord2.begin(), ord2.end(), [&](unsigned short a, unsigned short b2) { return row_len(a) > row_len(b2); });
    balance_store_halves(&ord2);
    TileHeader h{};
    h.n_rows = n_rows;
    h.h1 = h1;
    h.h2 = h2;
    h.nnz = nnz;
    int off = 64;
    h.off_halo = off; off += up16(h2 * 4);
    h.off_rp = off;   off += up16((h1 + 1) * 2);
    h.off_ent = off;  off += up16(nnz * 8);
    h.off_ord1 = off; off += up16(h1 * 2);
    h.off_ord2 = off; off += up16(TILE_M * 2);
    h.bytes = off;
    std::vector<unsigned char>& blob = blobs[pt];
    blob.assign(off, 0);
    std::memcpy(blob.data(), &h, sizeof(h));
    std::memcpy(blob.data() + h.off_halo, halo.data(), h2 * 4);
    std::memcpy(blob.data() + h.off_rp, rp.data(), (h1 + 1) * 2);
    if (nnz) std::memcpy(blob.data() + h.off_ent, ent.data(), (size_t)nnz * 8);
    std::memcpy(blob.data() + h.off_ord1, ord1.data(), h1 * 2);
    std::memcpy(blob.data() + h.off_ord2, ord2.data(), TILE_M * 2);
    max_h1 = std::max(max_h1, h1);
    max_h2 = std::max(max_h2, h2);
    stride = std::max(stride, off);
    {
      // trimmed variant: staged rows = own + 1-hop (h2 := h1), CSR rows of the 128 own rows (they come first in ent)
      const int nnz1 = rp[TILE_M <= h1 ? TILE_M : h1];
      TileHeader t{};
      t.n_rows = n_rows;
      t.h1 = h1;
      t.h2 = h1;
      t.nnz = nnz1;
      int o1 = 64;
      t.off_halo = o1; o1 += up16(h1 * 4);
      t.off_rp = o1;   o1 += up16((TILE_M + 1) * 2);
      t.off_ent = o1;  o1 += up16(nnz1 * 8);
      t.off_ord2 = o1; o1 += up16(TILE_M * 2);
      t.off_ord1 = t.off_ord2;  // not used by the consumers of this variant
      t.bytes = o1;
      std::vector<unsigned char>& b1 = blobs1[pt];
      b1.assign(o1, 0);
      std::memcpy(b1.data(), &t, sizeof(t));
      std::memcpy(b1.data() + t.off_halo, halo.data(), h1 * 4);
      std::memcpy(b1.data() + t.off_rp, rp.data(), (TILE_M + 1) * 2);
      if (nnz1) std::memcpy(b1.data() + t.off_ent, ent.data(), (size_t)nnz1 * 8);
      std::memcpy(b1.data() + t.off_ord2, ord2.data(), TILE_M * 2);
      stride1 = std::max(stride1, o1);
    }
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
  {
    stride1 = (stride1 + 127) & ~127;
    std::vector<unsigned char> all1((size_t)P * stride1, 0);
    std::vector<int> bytes1(P);
    for (int pt = 0; pt < P; ++pt) {
      std::memcpy(all1.data() + (size_t)pt * stride1, blobs1[pt].data(), blobs1[pt].size());
      bytes1[pt] = (int)blobs1[pt].size();
    }
    unsigned char* d_meta1 = nullptr;
    int* d_bytes1 = nullptr;
    P2M_CUDA_OK(cudaMalloc(&d_meta1, all1.size()));
    owned->push_back(d_meta1);
    P2M_CUDA_OK(cudaMalloc(&d_bytes1, sizeof(int) * P));
    owned->push_back(d_bytes1);
    P2M_CUDA_OK(cudaMemcpy(d_meta1, all1.data(), all1.size(), cudaMemcpyHostToDevice));
    P2M_CUDA_OK(cudaMemcpy(d_bytes1, bytes1.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
    out->tile_meta1 = d_meta1;
    out->tile_meta1_bytes = d_bytes1;
    out->meta1_stride = stride1;
  }
  {
    // padding-vertex elision: rows whose only entry is the diagonal, all with the same value
    std::vector<int> real_rows, iso_rows;
    float c = 0.f;
    bool uniform = true;
    for (int v = 0; v < V; ++v) {
      const bool iso = (rowptr[v + 1] - rowptr[v] == 1) && colidx[rowptr[v]] == v;
      if (iso) {
        if (iso_rows.empty()) c = val[rowptr[v]];
        if (val[rowptr[v]] != c) uniform = false;
        iso_rows.push_back(v);
      } else {
        real_rows.push_back(v);
      }
    }
    out->n_iso = 0;
    if (uniform && (int)iso_rows.size() >= TILE_M && (int)real_rows.size() >= TILE_M) {
      TileSet rt, it;
      if (build_tileset(real_rows, rowptr, colidx, val, V, &rt, owned) == P2M_OK &&
          build_tileset(iso_rows, rowptr, colidx, val, V, &it, owned) == P2M_OK && rt.max_h1 <= 256) {
        out->real_tiles = rt;
        out->iso_tiles = it;
        out->n_iso = (int)iso_rows.size();
        out->iso_diag = c;
      }
    }
  }
  out->n_pattern = P;
  out->tile_meta = d_meta;
  out->tile_meta_bytes = d_bytes;
  out->meta_stride = stride;
  out->max_h1 = max_h1;
  out->max_h2 = max_h2;
  return P2M_OK;
}


size_t dw_smem_bytes(int XS, const DevLevel& g, bool t1_given = false) {
  const size_t stage = t1_given ? (size_t)XS * (TILE_M + g.max_h1) * FC * 4
                                : (size_t)XS * g.max_h2 * FC * 4 + (size_t)g.max_h1 * FC * 4;
  return 1024 + (size_t)DW_NS * A_BLOCK_BYTES + DW_G_BYTES + stage + 2 * (size_t)(t1_given ? g.meta1_stride : g.meta_stride) +
         8 * (2 * DW_NS + 2 * XS + 8) + 32;
}

bool umma_dw_supported(const DevLevel& g, int fin, int fout) {
  if (g.tile_meta == nullptr || g.n_pattern <= 0 || g.max_h1 > 256) return false;
  if (fin % FC != 0 || fin < FC || fin > 256) return false;
  if (fout != 64 && fout != 128 && fout != 256) return false;
  return dw_smem_bytes(1, g) <= SMEM_LIMIT;
}

namespace {
int launch_dw_kernels(const DevLevel& g, DwParams p, int gathered_width, int plain_width, bool t1_given, int sm_count,
                      cudaStream_t s) {
  const int xs = dw_smem_bytes(2, g, t1_given) <= SMEM_LIMIT ? 2 : 1;
  const size_t smem = dw_smem_bytes(xs, g, t1_given);
  if (smem > SMEM_LIMIT) {
    set_error("umma_dw: does not fit shared memory");
    return P2M_ERR_INVALID;
  }
  auto kern = (xs == 2) ? k_cheb_dw_umma<2> : k_cheb_dw_umma<1>;
  P2M_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(p.n_tiles, sm_count);
  const int total_chunks = gathered_width / FC;
  for (int m_off = 0; m_off < plain_width; m_off += 128) {
    p.m_off = m_off;
    p.m_cols = std::min(128, plain_width - m_off);
    const int per_launch = p.n32 ? 4 : 2;
    for (int c0 = 0; c0 < total_chunks; c0 += per_launch) {
      p.chunk0 = c0;
      p.n_chunk = std::min(per_launch, total_chunks - c0);
      kern<<<grid, NUM_THREADS2, smem, s>>>(p);
      P2M_LAUNCH_OK();
    }
  }
  return P2M_OK;
}
}  // namespace

int launch_umma_dw(const DevLevel& g, const float* x, int in_unpool, int batch, int fin, int fout, const float* dz,
                   const float* a_scale, float* dw_ref, int* status, int sm_count, cudaStream_t s) {
  if (!umma_dw_supported(g, fin, fout)) {
    set_error("umma_dw: unsupported shape");
    return P2M_ERR_INVALID;
  }
  DwParams p;
  p.x = x;
  p.in_unpool = in_unpool;
  p.V = g.V;
  p.P = g.n_pattern;
  p.fin = fin;
  p.n_tiles = batch * g.n_pattern;
  p.meta = g.tile_meta;
  p.meta_bytes = g.tile_meta_bytes;
  p.meta_stride = g.meta_stride;
  p.max_h1 = g.max_h1;
  p.max_h2 = g.max_h2;
  p.g = dz;
  p.fout_total = fout;
  p.a_scale = a_scale;
  p.dw = dw_ref;
  p.status = status;
  p.t1 = nullptr;
  p.g_unpool = 0;
  p.swap = 0;
  p.n32 = 0;
  p.tma = 0;
  std::memset(&p.tm_x, 0, sizeof(p.tm_x));
  std::memset(&p.tm_t1, 0, sizeof(p.tm_t1));
  return launch_dw_kernels(g, p, fin, fout, false, sm_count, s);
}

// dW from the basis of the GRADIENT (see DwParams::swap): dz [rows, fout], t1_dz = L~ dz for EVERY row of the level,
// x the layer input [rows(/2), fin]
bool umma_dw_swapped_supported(const DevLevel& g, int fin, int fout) {
  if (g.tile_meta1 == nullptr || g.n_pattern <= 0 || g.max_h1 > 256) return false;
  if (fout % FC != 0 || fout < FC || fout > 256) return false;       // gathered side: dz
  if (fin != 64 && fin != 128 && fin != 256) return false;           // plain side: x
  return dw_smem_bytes(1, g, true) <= SMEM_LIMIT;
}
int launch_umma_dw_swapped(const DevLevel& g, const float* x, int in_unpool, int batch, int fin, int fout,
                           const float* dz, const float* t1_dz, const float* a_scale, float* dw_ref, int* status,
                           int sm_count, cudaStream_t s) {
  if (!umma_dw_swapped_supported(g, fin, fout) || t1_dz == nullptr) {
    set_error("umma_dw_swapped: unsupported shape");
    return P2M_ERR_INVALID;
  }
  DwParams p;
  p.tma = 0;
  p.n32 = g_dw_n32 ? 1 : 0;
  std::memset(&p.tm_x, 0, sizeof(p.tm_x));
  std::memset(&p.tm_t1, 0, sizeof(p.tm_t1));
  p.x = dz;
  p.in_unpool = 0;
  p.V = g.V;
  p.P = g.n_pattern;
  p.fin = fout;                 // width of the gathered tensor
  p.n_tiles = batch * g.n_pattern;
  p.meta = g.tile_meta1;
  p.meta_bytes = g.tile_meta1_bytes;
  p.meta_stride = g.meta1_stride;
  p.max_h1 = g.max_h1;
  p.max_h2 = g.max_h1;
  p.g = x;
  p.fout_total = fin;           // width of the plain tensor
  p.a_scale = a_scale;
  p.dw = dw_ref;
  p.status = status;
  p.t1 = t1_dz;
  p.g_unpool = in_unpool;
  p.swap = 1;
  if (g.V % TILE_M == 0 && g_umma_tma) {
    const long long rows = (long long)batch * g.V;
    p.tma = (make_row_tmap(&p.tm_x, dz, rows, fout, TILE_M) && make_row_tmap(&p.tm_t1, t1_dz, rows, fout, TILE_M)) ? 1 : 0;
  }
  return launch_dw_kernels(g, p, fout, fin, true, sm_count, s);
}

int launch_cheb_t1(const DevLevel& g, const float* x, int in_unpool, int batch, int fin, float* t1, cudaStream_t s,
                   const TileSet* tiles) {
  if (g.tile_meta == nullptr || fin % FC != 0) {
    set_error("cheb_t1: unsupported shape");
    return P2M_ERR_INVALID;
  }
  const int max_h1 = tiles ? tiles->max_h1 : g.max_h1;
  const int stride = tiles ? tiles->stride : g.meta1_stride;
  const int n_pattern = tiles ? tiles->n_pattern : g.n_pattern;
  if (max_h1 > 256) {  // 4 staged slots per row group
    set_error("cheb_t1: halo too large");
    return P2M_ERR_INVALID;
  }
  auto smem_for = [&](int stages) { return (size_t)stride + stages * (size_t)max_h1 * FC * 4 + 16; };
  const size_t half_sm = (228 * 1024) / 2 - 1024;  // two CTAs per SM (1 KB per CTA is reserved by the system)
  const int stages = smem_for(4) <= half_sm ? 4 : (smem_for(3) <= half_sm ? 3 : 2);
  const size_t smem = smem_for(stages);
  auto kern = stages == 4 ? k_cheb_t1<4> : (stages == 3 ? k_cheb_t1<3> : k_cheb_t1<2>);
  P2M_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  T1Params p;
  p.x = x;
  p.in_unpool = in_unpool;
  p.V = g.V;
  p.P = n_pattern;
  p.fin = fin;
  p.meta = tiles ? tiles->meta : g.tile_meta1;
  p.meta_bytes = tiles ? tiles->bytes : g.tile_meta1_bytes;
  p.meta_stride = stride;
  p.max_h1 = max_h1;
  p.t1 = t1;
  kern<<<batch * n_pattern, 512, smem, s>>>(p);
  P2M_LAUNCH_OK();
  return P2M_OK;
}

bool umma_conv_supported(const DevLevel& g, int fin, int fout) {
  if (g.tile_meta == nullptr || g.n_pattern <= 0) return false;
  if (fin % FC != 0 || fin < FC || fin > 256) return false;
  if (g.max_h1 > 256) return false;  // producers keep <= 4 T1 rows per row group
  if (fout != 64 && fout != 128 && fout != 256) return false;
  return smem_bytes_for(fout, ring_stages(fout), 1, g) <= SMEM_LIMIT;
}

__global__ void __launch_bounds__(256) k_pack_plain(const float* __restrict__ Bmat, long long ld_n, long long ld_k, int N,
                                                    int K, unsigned char* __restrict__ out, float W_SCALE = 64.f) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk each
  const int total = (K / FC) * N * 8;
  if (idx >= total) return;
  const int j = idx & 7;
  const int n = (idx >> 3) % N;
  const int c = (idx >> 3) / N;
  const int k0 = c * FC + (j & 3) * 8;
  __align__(16) __half h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float w = Bmat[(long long)n * ld_n + (long long)(k0 + e) * ld_k] * W_SCALE;
    const __half hi = __float2half_rn(w);
    h[e] = (j < 4) ? hi : __float2half_rn(w - __half2float(hi));
  }
  *reinterpret_cast<uint4*>(out + (size_t)c * N * 128 + sw128_off(n, j)) = *reinterpret_cast<const uint4*>(h);
}

// the same image for the combined weights of the isolated rows, straight from the reference layout
__global__ void __launch_bounds__(256) k_pack_iso(const float* __restrict__ W, float c, int fin, int fout,
                                                  unsigned char* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk each
  const int total = (fin / FC) * fout * 8;
  if (idx >= total) return;
  const int j = idx & 7;
  const int n = (idx >> 3) % fout;
  const int cc = (idx >> 3) / fout;
  const int f0 = cc * FC + (j & 3) * 8;
  const float c2 = 2.f * c * c - 1.f;
  __align__(16) __half h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float* wr = W + (size_t)n * fin * 3 + (size_t)(f0 + e) * 3;
    const float w = (wr[0] + c * wr[1] + c2 * wr[2]) * W_SCALE;
    const __half hi = __float2half_rn(w);
    h[e] = (j < 4) ? hi : __float2half_rn(w - __half2float(hi));
  }
  *reinterpret_cast<uint4*>(out + (size_t)cc * fout * 128 + sw128_off(n, j)) = *reinterpret_cast<const uint4*>(h);
}
int launch_umma_pack_iso(const float* W, float c, int fin, int fout, void* wpack, cudaStream_t s) {
  const int total = (fin / FC) * fout * 8;
  k_pack_iso<<<(total + 255) / 256, 256, 0, s>>>(W, c, fin, fout, static_cast<unsigned char*>(wpack));
  P2M_LAUNCH_OK();
  return P2M_OK;
}

// Backward-data as a forward conv (p2m_api.cu): dX = [dz | L~dz | (2L~^2 - I)dz] * W'^T with W'[f][o*3 + k] = W[o][f*3 + k]
// (L~ symmetric).  Same K-block image as k_pack_weights for a layer with Fin' = fout, Fout' = fin.
__global__ void __launch_bounds__(256) k_pack_weights_t(const float* __restrict__ W, int fin, int fout,
                                                        unsigned char* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk each
  const int n_blocks = (fout / FC) * 3;
  const int total = n_blocks * fin * 8;
  if (idx >= total) return;
  const int j = idx & 7;
  const int n = (idx >> 3) % fin;   // output channel of the backward conv = input feature f
  const int u = (idx >> 3) / fin;
  const int c = u / 3, k = u % 3;
  const int o0 = c * FC + (j & 3) * 8;
  __align__(16) __half h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float w = W[(size_t)(o0 + e) * fin * 3 + (size_t)n * 3 + k] * W_SCALE;
    const __half hi = __float2half_rn(w);
    h[e] = (j < 4) ? hi : __float2half_rn(w - __half2float(hi));
  }
  *reinterpret_cast<uint4*>(out + (size_t)u * fin * 128 + sw128_off(n, j)) = *reinterpret_cast<const uint4*>(h);
}
int launch_umma_pack_weights_t(const float* W, int fin, int fout, void* wpack, cudaStream_t s) {
  const int total = (fout / FC) * 3 * fin * 8;
  k_pack_weights_t<<<(total + 255) / 256, 256, 0, s>>>(W, fin, fout, static_cast<unsigned char*>(wpack));
  P2M_LAUNCH_OK();
  return P2M_OK;
}
// ... and the combined weights of the isolated rows, transposed: B[n = f][o] = W[o][3f] + c W[o][3f+1] + (2c^2-1) W[o][3f+2]
__global__ void __launch_bounds__(256) k_pack_iso_t(const float* __restrict__ W, float c, int fin, int fout,
                                                    unsigned char* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = (fout / FC) * fin * 8;
  if (idx >= total) return;
  const int j = idx & 7;
  const int n = (idx >> 3) % fin;
  const int cc = (idx >> 3) / fin;
  const int o0 = cc * FC + (j & 3) * 8;
  const float c2 = 2.f * c * c - 1.f;
  __align__(16) __half h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float* wr = W + (size_t)(o0 + e) * fin * 3 + (size_t)n * 3;
    const float w = (wr[0] + c * wr[1] + c2 * wr[2]) * W_SCALE;
    const __half hi = __float2half_rn(w);
    h[e] = (j < 4) ? hi : __float2half_rn(w - __half2float(hi));
  }
  *reinterpret_cast<uint4*>(out + (size_t)cc * fin * 128 + sw128_off(n, j)) = *reinterpret_cast<const uint4*>(h);
}
int launch_umma_pack_iso_t(const float* W, float c, int fin, int fout, void* wpack, cudaStream_t s) {
  const int total = (fout / FC) * fin * 8;
  k_pack_iso_t<<<(total + 255) / 256, 256, 0, s>>>(W, c, fin, fout, static_cast<unsigned char*>(wpack));
  P2M_LAUNCH_OK();
  return P2M_OK;
}

size_t umma_plain_pack_bytes(int N, int K) { return (size_t)(K / FC) * N * 128; }

// ---------------------------------------------------------------- dense GEMM on the conv kernel's plain mode
// A operand image of a row-major activation matrix X [M, K] (K % 32 == 0): per (128-row tile, 32-column chunk) one
// 16 KB block [128 rows][hi 32 | lo 32] fp16, 128B-swizzled, rows >= M zero — what the producers would have built.
__global__ void __launch_bounds__(256) k_pack_rows(const float* __restrict__ X, int M, int K, unsigned char* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk each
  const int n_chunk = K / FC;
  const long long total = (long long)((M + TILE_M - 1) / TILE_M) * n_chunk * TILE_M * 8;
  if (idx >= total) return;
  const int j = (int)(idx & 7);
  const int row = (int)((idx >> 3) % TILE_M);
  const long long blk = (idx >> 3) / TILE_M;  // tile * n_chunk + c
  const int c = (int)(blk % n_chunk);
  const long long r = (blk / n_chunk) * TILE_M + row;
  __align__(16) __half h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float v = (r < M) ? X[r * K + c * FC + (j & 3) * 8 + e] : 0.f;
    const __half hi = __float2half_rn(v);
    h[e] = (j < 4) ? hi : __float2half_rn(v - __half2float(hi));
  }
  *reinterpret_cast<uint4*>(out + (size_t)blk * A_BLOCK_BYTES + sw128_off(row, j)) = *reinterpret_cast<const uint4*>(h);
}
// Weight images of ALL output-column slices in one launch: slice j (rows [j ns, (j+1) ns) of W [n_real, K], rows >=
// n_real zero) -> K/32 blocks of ns rows x 128 bytes [Whi | Wlo] (scaled by 2^6 like every weight image).
__global__ void __launch_bounds__(256) k_pack_w_sliced(const float* __restrict__ W, int n_real, int N, int K, int ns,
                                                       unsigned char* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk each
  const int n_chunk = K / FC;
  const long long total = (long long)N * n_chunk * 8;
  if (idx >= total) return;
  const int j = (int)(idx & 7);
  const int nl = (int)((idx >> 3) % ns);
  const long long blk = (idx >> 3) / ns;  // slice * n_chunk + c
  const int c = (int)(blk % n_chunk);
  const int n = (int)(blk / n_chunk) * ns + nl;
  __align__(16) __half h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float w = (n < n_real) ? W[(size_t)n * K + c * FC + (j & 3) * 8 + e] * W_SCALE : 0.f;
    const __half hi = __float2half_rn(w);
    h[e] = (j < 4) ? hi : __float2half_rn(w - __half2float(hi));
  }
  *reinterpret_cast<uint4*>(out + (size_t)blk * ns * 128 + sw128_off(nl, j)) = *reinterpret_cast<const uint4*>(h);
}
size_t umma_gemm_apack_bytes(int M, int K) { return (size_t)((M + TILE_M - 1) / TILE_M) * (K / FC) * A_BLOCK_BYTES; }
size_t umma_gemm_wpack_bytes(int N, int K) { return (size_t)(K / FC) * N * 128; }
bool umma_gemm_supported(int M, int N, int K) { return M > 0 && K >= FC && K % FC == 0 && N >= 64 && N % 64 == 0; }

namespace {
template <int N>
int launch_gemm_cfg(KParams p, int n_slices, int sm_count, cudaStream_t s) {
  constexpr int NS = (N == 256) ? 2 : 3;
  const size_t smem = smem_bytes_dims(N, NS, 1, 0, 0, 0, 2);
  auto kern = k_cheb_conv_umma<N, NS, 1, 0>;
  P2M_TRY((check_launch_regs<N, NS, 1, 0, 0>()));
  P2M_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p.wslice_bytes = (long long)umma_gemm_wpack_bytes(N, p.fin);
  const dim3 grid(std::min(p.n_tiles, sm_count), n_slices);
  kern<<<grid, NUM_THREADS2, smem, s>>>(p);
  P2M_LAUNCH_OK();
  return P2M_OK;
}
}  // namespace

// Y [M, N] = epilogue( X [M, K] * W [N, K]^T )  on tcgen05 (fp16x3): X and W are packed into operand images first
// (apack / wpack: caller-provided scratch of umma_gemm_{a,w}pack_bytes), then every K-block of both operands is
// streamed by cp.async.bulk into the conv kernel's A/B ring — no producer warps, one CTA per (128-row tile, output
// column slice).  The epilogue vectors and an identity residual (ep.res with res_F == N) are indexed by output column.
int launch_umma_gemm(const float* X, const float* W, int M, int N, int K, const Epilogue& ep, float* Y, void* apack,
                     void* wpack, int* status, int sm_count, cudaStream_t s, int n_real) {
  if (n_real <= 0 || n_real > N) n_real = N;  // W has n_real rows; output columns >= n_real are zero-weight padding
  if (!umma_gemm_supported(M, N, K) || (ep.res != nullptr && ep.res_F != N)) {
    set_error("umma_gemm: unsupported shape");
    return P2M_ERR_INVALID;
  }
  {
    const long long total = (long long)((M + TILE_M - 1) / TILE_M) * (K / FC) * TILE_M * 8;
    k_pack_rows<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(X, M, K, static_cast<unsigned char*>(apack));
    P2M_LAUNCH_OK();
  }
  const int tiles = (M + TILE_M - 1) / TILE_M;
  const int ns = (tiles * (N / 128) < sm_count || N % 128 != 0) ? 64 : ((tiles * (N / 256) < sm_count || N % 256 != 0) ? 128 : 256);
  {
    const long long total = (long long)N * (K / FC) * 8;
    k_pack_w_sliced<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(W, n_real, N, K, ns, static_cast<unsigned char*>(wpack));
    P2M_LAUNCH_OK();
  }
  KParams p;
  std::memset(&p, 0, sizeof(p));
  p.V = M;                // one "mesh" of M rows: the epilogue masks rows >= V of the last tile
  p.P = tiles;
  p.fin = K;
  p.n_tiles = tiles;
  p.wpack = static_cast<const unsigned char*>(wpack);
  p.apack = static_cast<const unsigned char*>(apack);
  p.ep = to_dev(ep);
  p.res_identity = (ep.res != nullptr) ? 1 : 0;
  p.plain = 1;
  p.ldy = N;
  p.y = Y;
  p.status = status;
  switch (ns) {
    case 64: return launch_gemm_cfg<64>(p, N / ns, sm_count, s);
    case 128: return launch_gemm_cfg<128>(p, N / ns, sm_count, s);
    default: return launch_gemm_cfg<256>(p, N / ns, sm_count, s);
  }
}

int launch_umma_pack_plain(const float* Bmat, long long ld_n, long long ld_k, int N, int K, void* wpack, cudaStream_t s) {
  const int total = (K / FC) * N * 8;
  k_pack_plain<<<(total + 255) / 256, 256, 0, s>>>(Bmat, ld_n, ld_k, N, K, static_cast<unsigned char*>(wpack));
  P2M_LAUNCH_OK();
  return P2M_OK;
}

size_t umma_wpack_bytes(int fin, int fout) { return (size_t)(fin / FC) * 3 * fout * 128; }

int launch_umma_pack_weights(const float* W, int fin, int fout, void* wpack, cudaStream_t s) {
  const int total = (fin / FC) * 3 * fout * 8;
  k_pack_weights<<<(total + 255) / 256, 256, 0, s>>>(W, fin, fout, static_cast<unsigned char*>(wpack));
  P2M_LAUNCH_OK();
  return P2M_OK;
}

int launch_umma_conv(const UmmaConvArgs& a, int* status, const float* zero_row, int sm_count, cudaStream_t s) {
  if (!umma_conv_supported(*a.g, a.fin, a.fout)) {
    set_error("umma_conv: unsupported shape");
    return P2M_ERR_INVALID;
  }
  if (a.head_z != nullptr && (a.fout != 64 || a.head_wt == nullptr || a.ep.res != nullptr)) {
    set_error("umma_conv: the fused head needs fout == 64 and no residual");
    return P2M_ERR_INVALID;
  }
  if (a.tiles != nullptr && ((a.t1 == nullptr && !a.plain) || a.tiles->max_h1 > 256 || a.tiles->n_pattern <= 0)) {
    set_error("umma_conv: index-list tiles need the T1-given or plain mode and at most 256 staged rows");
    return P2M_ERR_INVALID;
  }
  switch (a.fout) {
    case 64: return launch_n<64>(a, status, zero_row, sm_count, s);
    case 128: return launch_n<128>(a, status, zero_row, sm_count, s);
    case 256: return launch_n<256>(a, status, zero_row, sm_count, s);
  }
  return P2M_ERR_INVALID;
}

}  // namespace p2m
