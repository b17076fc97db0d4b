// Internal (C++) interface between the C-ABI layer (p2m_api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/p2m_b200.h"

namespace p2m {

// ---------------------------------------------------------------- error plumbing (thread-local)
void set_error(const std::string& msg);
void count_launch(int n = 1);

#define P2M_CUDA_OK(expr)                                                                     \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      p2m::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" +        \
                     __FILE__ + ":" + std::to_string(__LINE__) + ")");                        \
      return P2M_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define P2M_LAUNCH_OK()                                                                       \
  do {                                                                                        \
    p2m::count_launch();                                                                      \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      p2m::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e) + " (" +  \
                     __FILE__ + ":" + std::to_string(__LINE__) + ")");                        \
      return P2M_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define P2M_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != P2M_OK) return _s; \
  } while (0)

// ---------------------------------------------------------------- device-resident hierarchy level
// L~ of one level in CSR with RELATIVE column offsets: the neighbour of flat activation row
// r = b*V + v is row r + reloff[p], so kernels never need (b, v) separately (block-diagonal I_B (x) L~).
// A family of 128-row tile patterns of one level whose rows are given by index lists (trimmed blobs: the tile's own
// rows, their 1-hop halo, the CSR of the own rows) instead of being the consecutive rows [128 p, 128 p + 128).
struct TileSet {
  const unsigned char* meta = nullptr;  // [n_pattern][stride]
  const int* bytes = nullptr;           // [n_pattern]
  int stride = 0;
  int n_pattern = 0;                    // tiles per mesh
  int max_h1 = 0;                       // largest own + 1-hop row count
};

struct DevLevel {
  int V = 0;
  int nnz = 0;
  int max_row_nnz = 0;
  int* rowptr = nullptr;   // [V+1]
  int* reloff = nullptr;   // [nnz] col - row
  float* val = nullptr;    // [nnz]
  // tcgen05 path: per-tile-pattern halo / local-CSR blobs (cheb_umma.cu)
  int n_pattern = 0;                    // tiles per mesh = ceil(V / 128)
  const unsigned char* tile_meta = nullptr;   // [n_pattern][meta_stride]
  const int* tile_meta_bytes = nullptr;       // [n_pattern] bytes to copy (multiple of 16)
  int meta_stride = 0;
  // trimmed blobs for the kernels that get X / T1 rows instead of recomputing them (k_cheb_t1, conv with T1 given):
  // halo list up to the 1-hop rows and the CSR rows of the tile's own 128 rows only
  const unsigned char* tile_meta1 = nullptr;  // [n_pattern][meta1_stride]
  const int* tile_meta1_bytes = nullptr;
  int meta1_stride = 0;
  int max_h1 = 0, max_h2 = 0;
  // Padding-vertex elision: the fake vertices of the binary-tree reorder (lib/coarsening.py:214-258) are isolated in
  // L~ and share one diagonal value iso_diag, so on them the conv is a dense map with the combined weights
  // W0 + c W1 + (2c^2 - 1) W2.  real_tiles covers the connected rows (conv path), iso_tiles the isolated ones (plain
  // GEMM).  n_iso == 0: not applicable on this level.
  int n_iso = 0;
  float iso_diag = 0.f;
  TileSet real_tiles, iso_tiles;
  // Eval-mode duplicate elimination among the isolated rows (p2m_api.cu: build_padding_classes).  The two children
  // of a fake vertex are fake too and, with BatchNorm folded, carry identical values in every layer of their level
  // (same parent row through the unpool, same dense map): only one REPRESENTATIVE per class is computed
  // (rep_tiles, plain GEMM like iso_tiles) and the network's output rows of the others are filled from it at the
  // end (copy_dst[i] <- copy_src[i], finest level only).  n_rep == 0: not applicable on this level.
  TileSet rep_tiles;
  int n_rep = 0;
  int n_copy = 0;
  int* copy_dst = nullptr;
  int* copy_src = nullptr;
};

// 2-tap channel resampling table (F.interpolate(mode='linear', align_corners=False) along channels,
// meshnet.py:109,114) and its transpose for the backward pass.
struct InterpTable {
  int fin = 0, fout = 0;
  int* i0 = nullptr;      // [fout]
  int* i1 = nullptr;      // [fout]
  float* lam = nullptr;   // [fout]  out[j] = (1-lam)*x[i0] + lam*x[i1]
  int* t_ptr = nullptr;   // [fin+1]   transpose CSR: dx[i] = sum_p t_w[p] * dout[t_idx[p]]
  int* t_idx = nullptr;
  float* t_w = nullptr;
};

// ---------------------------------------------------------------- SIMT kernels (kernels_simt.cu)
// Chebyshev basis T = [T0 | T1 | T2] (each F wide, row stride 3F) of x.  x is [rows_phys, F]; when
// in_unpool, logical row r reads physical row r>>1 (nearest x2 unpool, meshnet.py:71-78).
int launch_cheb_basis(const DevLevel& g, const float* x, int in_unpool, int rows, int F, float* T, cudaStream_t s);

struct Epilogue {
  const float* bias = nullptr;    // [N] added first
  const float* scale = nullptr;   // [N] then v = v*scale + shift
  const float* shift = nullptr;
  int relu = 0;
  const float* res = nullptr;     // residual source [rows(/2), res_F], channel-resampled to N, added last
  int res_F = 0;
  int res_unpool = 0;
  const int* res_i0 = nullptr;
  const int* res_i1 = nullptr;
  const float* res_lam = nullptr;
  // optional fused output gather (the callers' pred[:, perm_reverse[:n_real]], lib/core/base.py:130): row v of a
  // mesh is stored at slot out_map[v] of a [B, out_rows, N] tensor, or dropped when out_map[v] < 0
  const int* out_map = nullptr;
  int out_rows = 0;
  int level_V = 0;
};
#ifdef __CUDACC__
// Device-side view of Epilogue + the per-element epilogue shared by the SIMT GEMM and the tcgen05 conv.
struct EpiDev {
  const float* bias;
  const float* scale;
  const float* shift;
  int relu;
  const float* res;
  int res_F;
  int res_unpool;
  const int* i0;
  const int* i1;
  const float* lam;
  const int* out_map;
  int out_rows;
  int level_V;
};
inline EpiDev to_dev(const Epilogue& e) {
  return EpiDev{e.bias, e.scale, e.shift, e.relu, e.res, e.res_F, e.res_unpool, e.res_i0, e.res_i1, e.res_lam,
                e.out_map, e.out_rows, e.level_V};
}
__device__ __forceinline__ float apply_epilogue(float v, long long r, int n, const EpiDev& ep) {
  if (ep.bias) v += ep.bias[n];
  if (ep.scale) v = fmaf(v, ep.scale[n], ep.shift[n]);
  if (ep.relu) v = fmaxf(v, 0.f);
  if (ep.res) {
    long long pr = ep.res_unpool ? (r >> 1) : r;
    const float* rr = ep.res + pr * ep.res_F;
    if (ep.lam == nullptr) {  // no resampling table: plain residual (res_F == N)
      v += rr[n];
    } else {
      float l = ep.lam[n];
      v += (1.f - l) * rr[ep.i0[n]] + l * rr[ep.i1[n]];
    }
  }
  return v;
}
#endif

// C[M,N] = A[M,K] * op(B) (+ epilogue); b_is_kn: B stored [K,N] row-major, else [N,K] row-major.
int launch_gemm(const float* A, int lda, const float* B, int ldb, int b_is_kn, float* C, int ldc, int M, int N, int K,
                const Epilogue& ep, cudaStream_t s);
// C[N1,N2] (+)= A[M,N1]^T * B[M,N2]  (C must be zeroed by the caller; split over M with atomics)
int launch_gemm_tn_atomic(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N1, int N2,
                          cudaStream_t s);

// Thin-output conv (Fout <= 4), weights first; scratch >= thin_conv_scratch_floats(rows, fin) floats.
bool thin_conv_supported(int fin, int fout);
size_t thin_conv_scratch_floats(long long rows, int fin);
int launch_thin_conv(const DevLevel& g, const float* x, int in_unpool, int rows, int fin, int fout, const float* W,
                     const Epilogue& e, float* scratch, float* y, cudaStream_t s);
// Backward of the thin conv (fin == 64, fout <= 4, no unpool): dx [rows, fin] (may be null), dw [fout, 3 fin] in the
// reference layout (zeroed here); scratch >= thin_conv_bwd_scratch_floats(rows, fin) floats.
bool thin_conv_bwd_supported(int fin, int fout);
size_t thin_conv_bwd_scratch_floats(long long rows, int fin);
int launch_thin_conv_bwd(const DevLevel& g, const float* x, int rows, int fin, int fout, const float* W, const float* dz,
                         float* scratch, float* dx, float* dw, int sm_count, cudaStream_t s);
// The same head in two pieces, for the fused eval path: the 128 -> 64 conv's epilogue produces Z = Y W' itself
// (head_wt / head_z of UmmaConvArgs), the tail applies the two sparse products on the 4-wide rows.
int launch_thin_prep(const float* W, int fin, int fout, float* wt, cudaStream_t s);
int launch_thin_tail(const DevLevel& g, int rows, int fout, const float* Z, float* U, const Epilogue& e, float* y,
                     cudaStream_t s);

int launch_permute_w(const float* W, float* Wp, int fout, int fin, cudaStream_t s);      // [n,f*3+k] -> [n,k*fin+f]
int launch_unpermute_w(const float* Wp, float* W, int fout, int fin, cudaStream_t s);    // inverse
int launch_fill_zero(void* p, size_t bytes, cudaStream_t s);
// y[b, dst[i], :] = y[b, src[i], :]  for i < n, b < batch  (y [batch, V, F])
int launch_copy_rows(float* y, int batch, int V, int F, const int* dst, const int* src, int n, cudaStream_t s);

// BatchNorm1d over rows (cheby_graph_conv.py:38-39; meshnet.py:55): eps 1e-5, momentum 0.1
int launch_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv, const float* bias,
                        float* scale, float* shift, int F, cudaStream_t s);
int launch_col_stats(const float* z, int rows, int F, double* sums /*[2F] zeroed here*/, cudaStream_t s);
int launch_bn_finalize(const double* sums, int rows, int F, const float* gamma, const float* beta, float* rm, float* rv,
                       int64_t* nbt, float* save_mean, float* save_invstd, float* scale, float* shift, cudaStream_t s);
// a = relu?(z*scale+shift) (+ resampled residual)
int launch_affine_act(const float* z, int rows, int F, const float* scale, const float* shift, int relu,
                      const float* res, int res_F, int res_unpool, const InterpTable* it, float* a, cudaStream_t s);
// BN+ReLU backward: g_a (grad wrt a = relu(bn(z)) [+ residual]) -> g_z (in place allowed); dgamma, dbeta
// written.  The ReLU mask is recomputed from z (block-end activations already include the residual).
int launch_bn_relu_bwd(const float* z, const float* g_a, int rows, int F, const float* gamma, const float* scale,
                       const float* shift, const float* mean, const float* invstd, int relu,
                       double* sums /*scratch: 2F doubles + 5F floats, 16-byte aligned*/, float* dgamma, float* dbeta,
                       float* g_z, cudaStream_t s,
                       float* gz_scale_out = nullptr /* optional device scalar: launch_absmax_scale(g_z) fused in */);
int launch_col_sum(const float* g, int rows, int F, double* scratch /*[F]*/, float* out, cudaStream_t s);

// dX of the Chebyshev basis: given dT [rows,3F] (blocks dT0|dT1|dT2):
//   dXl = dT0 - dT2 + L~ (dT1 + 2 L~ dT2)   (L~ symmetric)   [+ resample^T(g_res)]
// written to dx; when out_pairsum, dx has rows/2 rows and dx[p] = dXl[2p] + dXl[2p+1].
int launch_cheb_basis_bwd(const DevLevel& g, const float* dT, int rows, int F, float* U /*[rows,F] scratch*/,
                          const float* g_res, int res_Fout, const InterpTable* it, int out_pairsum, float* dx,
                          cudaStream_t s);

// ---------------------------------------------------------------- tcgen05 path (cheb_umma.cu)
struct UmmaConvArgs {
  const DevLevel* g;
  const float* x;           // [rows(/2), Fin]
  int in_unpool;
  int batch;                // rows = batch * g->V
  int fin, fout;
  const void* wpack;        // packed fp16 hi/lo weights from launch_umma_pack_weights
  Epilogue ep;
  float* y;                 // [rows, fout]
  // plain-GEMM mode (backward dT = dz * W_k): no SpMM, x is [rows, fin] and the K-blocks come from
  // launch_umma_pack_plain; y is written at y[r*ldy + y_col0 + n]
  const float* t1 = nullptr;        // optional precomputed T1 = L~ x [rows, fin] (launch_cheb_t1)
  int plain = 0;
  const float* a_scale = nullptr;   // device scalar from launch_absmax_scale (or null)
  long long ldy = 0;                // 0: fout
  int y_col0 = 0;
  // optional fused thin head (fout == 64 only): instead of storing y, the epilogue writes Z[row][12] = y_row * head_wt
  // (head_wt = k_thin_prep's [64][12] table); y is then not written at all
  const float* head_wt = nullptr;
  float* head_z = nullptr;
  // optional: run on this tile family instead of the level's consecutive 128-row tiles (T1-given and plain mode)
  const TileSet* tiles = nullptr;
  long long* trace = nullptr;       // debug (P2M_UMMA_TRACE builds only): [8][512] event log of CTA 0
};
// Host: build the per-tile halo metadata of one level (uploads; device pointers appended to `owned`).
int build_umma_level_meta(const int* rowptr, const int* colidx, const float* val, int V, DevLevel* out,
                          std::vector<void*>* owned);
// Tiles of 128 consecutive entries of `rows` (ascending vertex ids of one level) as a TileSet (trimmed blobs).
int build_index_tiles(const std::vector<int>& rows, const int* rowptr, const int* colidx, const float* val, int V,
                      TileSet* ts, std::vector<void*>* owned);
bool umma_conv_supported(const DevLevel& g, int fin, int fout);
size_t umma_wpack_bytes(int fin, int fout);
int launch_umma_pack_weights(const float* W /*[fout, fin*3] ref layout*/, int fin, int fout, void* wpack, cudaStream_t s);
// B[n][k] = Bmat[n*ld_n + k*ld_k]  (n < N, k < K, K % 32 == 0) -> fp16 [hi|lo] K-blocks, one per 32 k
size_t umma_plain_pack_bytes(int N, int K);
int launch_umma_pack_plain(const float* Bmat, long long ld_n, long long ld_k, int N, int K, void* wpack, cudaStream_t s);
// scale_out[0] = 2^e with max|x| * 2^e in [2^9, 2^10)  (1 if x is all zero); scratch-free, two tiny launches
int launch_absmax_scale(const float* x, long long n, float* scale_out, cudaStream_t s);
// dW[o, f*3+k] += sum_rows dz[row,o] * T_k(x)[row,f] on tensor cores (dw_ref zeroed by the caller)
bool umma_dw_supported(const DevLevel& g, int fin, int fout);
int launch_umma_dw(const DevLevel& g, const float* x, int in_unpool, int batch, int fin, int fout, const float* dz,
                   const float* a_scale, float* dw_ref, int* status, int sm_count, cudaStream_t s);
// The same sum through the basis of the gradient (L~ symmetric): dW[o, f*3+k] += sum_rows T_k(dz)[row,o] * x[row,f], with
// t1_dz = L~ dz [rows, fout] given for EVERY row (the backward-data pass leaves it behind): no 2-hop halo, no T1 on chip
bool umma_dw_swapped_supported(const DevLevel& g, int fin, int fout);
int launch_umma_dw_swapped(const DevLevel& g, const float* x, int in_unpool, int batch, int fin, int fout,
                           const float* dz, const float* t1_dz, const float* a_scale, float* dw_ref, int* status,
                           int sm_count, cudaStream_t s);
// T1 = L~ x for all rows of a level (tile-staged gather), t1 [batch*V, fin] fp32
int launch_cheb_t1(const DevLevel& g, const float* x, int in_unpool, int batch, int fin, float* t1, cudaStream_t s,
                   const TileSet* tiles = nullptr);
// K-blocks of the combined weights of the isolated rows: B[n][f] = W[n][3f] + c W[n][3f+1] + (2c^2 - 1) W[n][3f+2]
// (W in the reference layout [fout, fin*3]); same image as launch_umma_pack_plain(N = fout, K = fin)
int launch_umma_pack_iso(const float* W, float c, int fin, int fout, void* wpack, cudaStream_t s);
// Backward-data weights: the forward conv run on dz with W'[f][o*3+k] = W[o][f*3+k] gives dX (L~ is symmetric); images
// for a layer with Fin' = fout, Fout' = fin: umma_wpack_bytes(fout, fin) / umma_plain_pack_bytes(fin, fout) bytes
int launch_umma_pack_weights_t(const float* W /*[fout, fin*3]*/, int fin, int fout, void* wpack, cudaStream_t s);
int launch_umma_pack_iso_t(const float* W, float c, int fin, int fout, void* wpack, cudaStream_t s);
int launch_umma_conv(const UmmaConvArgs& a, int* status_flag, const float* zero_row, int sm_count, cudaStream_t s);
// Dense GEMM on tcgen05 (fp16x3): Y [M, N] = epilogue(X [M, K] W [N, K]^T), K % 32 == 0, N % 64 == 0; apack / wpack are
// scratch of umma_gemm_apack_bytes(M, K) / umma_gemm_wpack_bytes(N, K); ep vectors and an identity residual
// (ep.res, res_F == N) are indexed by output column.
bool umma_gemm_supported(int M, int N, int K);
size_t umma_gemm_apack_bytes(int M, int K);
size_t umma_gemm_wpack_bytes(int N, int K);
int launch_umma_gemm(const float* X, const float* W, int M, int N, int K, const Epilogue& ep, float* Y, void* apack,
                     void* wpack, int* status, int sm_count, cudaStream_t s, int n_real = 0 /* rows of W if < N */);
// out[ro, :] = sum over the logical rows r of physical row ro (r = ro, or 2 ro and 2 ro + 1 under the virtual unpool)
// of  dxl[r, :] + resample^T(g_res[r, :])   (g_res may be null)
int launch_dx_finish(const float* dxl, int rows, int F, const float* g_res, int res_Fout, const InterpTable* it,
                     int out_pairsum, float* out, cudaStream_t s);

}  // namespace p2m
