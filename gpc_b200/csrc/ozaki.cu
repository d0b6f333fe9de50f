// ozaki.cu -- fp64 GEMM/SYRK on the 5th-generation tensor cores of sm_100a (tcgen05 + TMEM + TMA).
//
// tcgen05.mma has no f64 kind, and the legacy DMMA path (dense.cu) tops out at 37 TFLOP/s.  The large trailing
// updates of the blocked Cholesky / inverse (the dsyrk_/dgemm_ work below dpotrf_/dpotri_, reference
// lapack.h:59-73, CMatrix.cpp:371-432) are therefore computed with the error-free "Ozaki" splitting on the INT8
// tensor pipe (tcgen05.mma.kind::i8, exact int32 accumulation in TMEM):
//
//   a_ik = 2^ea_i * sum_p A_p[i,k] 2^-(8p-2),  A_p int8   (row-wise exponent; first digit 6 bits + sign, then balanced
//                                                         base-256 digits in [-128, 127])
//   C_ij = 2^(ea_i+eb_j) * sum_{c=2}^{S+1} 2^-(8c-4) * ( sum_{p+q=c} A_p B_q' )_ij
//
// Every int8 product is exact; products of equal weight c = p+q share one int32 TMEM accumulator ("level"), so one
// 128 x 64 output tile owns S levels x 64 TMEM columns = all 512 for S = 8.  The S levels are combined in fp64 (Horner,
// exact power-of-two scalings) by the epilogue warps.  With S = 8 the operands are represented to 6 + 8*7 = 62 bits
// below the row maximum (fp64 has 53) and S(S+1)/2 = 36 int8 MMAs replace one fp64 MMA; S = 7 (54 bits, 28 MMAs) matches
// the fp64 pipe normwise but not for rows with a few dominant entries, where the error is relative to the row maximum.
//
// Kernel anatomy (one CTA per output tile, 6 warps):
//   warp 4 lane 0 : TMA producer -- slices are K-major int8 [slice][row][k]; a k-block (128 B of k) of all S B-slices
//                   (64 rows each, stacked along N) goes to one of two B sets, the A-slices (128 rows) stream through
//                   a ring of 16 KB slots.  128-byte swizzle, mbarrier transaction counts.
//   warp 5 lane 0 : MMA issuer -- for A_p the B slices q = 1..S+1-p are contiguous in smem, so ONE instruction with
//                   N = 64*(S+1-p) (<= 256) feeds S+1-p consecutive levels; tcgen05.commit releases smem slots.
//   warps 0-3     : epilogue -- tcgen05.ld 32x32b (one TMEM lane = one output row per thread), Horner over the levels,
//                   scale by 2^(ea_i+eb_j), C = alpha*acc + beta*C, coalesced column-major stores.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <map>
#include <mutex>
#include <vector>
#include "common.cuh"

#define GPC_CHECK(expr)            \
  do {                             \
    int _rc = (expr);              \
    if (_rc != GPC_OK) return _rc; \
  } while (0)

namespace gpc {

constexpr int OZ_BM = 128;        // output tile rows (= TMEM lanes, UMMA M)
constexpr int OZ_BN = 64;         // output tile columns per level
constexpr int OZ_BK = 128;        // bytes (= int8 elements) of k per k-block: one 128B-swizzle row
constexpr int OZ_UK = 32;         // k per tcgen05.mma.kind::i8
constexpr int OZ_MAXS = 8;        // slices (levels); S*OZ_BN <= 512 TMEM columns
constexpr int OZ_THREADS = 192;
constexpr int OZ_A_BYTES = OZ_BM * OZ_BK;   // 16 KB per A-slice slot
constexpr int OZ_B_BYTES = OZ_BN * OZ_BK;   //  8 KB per B-slice tile
constexpr int OZ_MAXNA = 8;
constexpr long long OZ_SPIN_LIMIT = 4000000000LL;  // ~2 s of clock64: a protocol error traps instead of hanging the GPU

// ------------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void oz_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void oz_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t oz_mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void oz_mbar_wait(uint32_t bar, uint32_t parity, int* errflag, int code) {
  if (oz_mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!oz_mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > OZ_SPIN_LIMIT) {
      if (errflag) atomicExch(errflag, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void oz_tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void oz_prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void oz_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void oz_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void oz_tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// int32 -> fp64 without I2F.F64 (a quarter-rate conversion on the fp64 pipe; the epilogue converts 512 values per
// thread): 2^52 + (x + 2^31) is built exactly from the bits, one full-rate DADD removes the offset
__device__ __forceinline__ double oz_i2d(int x) {
  return __hiloint2double(0x43300000, (int)((unsigned)x ^ 0x80000000u)) - 4503601774854144.0;  // 2^52 + 2^31
}
// two TMEM loads in flight behind ONE wait (halves the exposed TMEM round trips of the epilogue)
__device__ __forceinline__ void oz_tmem_ld16x2(uint32_t taddr0, uint32_t taddr1, int (&r0)[16], int (&r1)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%32];\n"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%33];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]), "=r"(r0[8]),
        "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11]), "=r"(r0[12]), "=r"(r0[13]), "=r"(r0[14]), "=r"(r0[15]), "=r"(r1[0]),
        "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]), "=r"(r1[8]), "=r"(r1[9]),
        "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15])
      : "r"(taddr0), "r"(taddr1)
      : "memory");
}
__device__ __forceinline__ void oz_tmem_ld16(uint32_t taddr, int (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// ------------------------------------------------------------------------------------------------------
// Slicing: fp64 operand (R rows = the m or n index, K deep) -> S int8 slices, K-major [slice][row][k], + row scales
//   kc = 1: element (r, kk) at g[kk + r*ld];   kc = 0: element (r, kk) at g[r + kk*ld]
// ------------------------------------------------------------------------------------------------------
// tri: the operand is triangular in (row, kk) -- +1: zero for kk < row, -1: zero for kk > row.  The zero part is
// never read (it may be uninitialised memory): rows of 128-row block b only look at kk >= 128 b (+1) / kk < 128 (b+1) (-1),
// the same 128-granular ranges the GEMM kernel walks.
__global__ void __launch_bounds__(256) oz_rowmax_kernel(const double* __restrict__ g, int64_t ld, int kc, int64_t K,
                                                       int64_t kchunk, int* __restrict__ emax, int tri) {
  const int tid = threadIdx.x;
  int64_t k0 = (int64_t)blockIdx.y * kchunk;
  int64_t k1 = (k0 + kchunk < K) ? k0 + kchunk : K;
  if (tri > 0 && k0 < (int64_t)blockIdx.x * 128) k0 = (int64_t)blockIdx.x * 128;
  if (tri < 0 && k1 > ((int64_t)blockIdx.x + 1) * 128) k1 = ((int64_t)blockIdx.x + 1) * 128;
  if (k0 >= k1) return;
  if (!kc) {
    // 128 rows per block; thread pair (h = 0/1) strides over kk
    const int64_t r = (int64_t)blockIdx.x * 128 + (tid & 127);
    int e = 0;
    for (int64_t kk = k0 + (tid >> 7); kk < k1; kk += 2) {
      unsigned hi = (unsigned)__double2hiint(g[r + kk * ld]);
      int ex = (int)((hi >> 20) & 0x7ffu);
      e = ex > e ? ex : e;
    }
    atomicMax(&emax[r], e);
  } else {
    // 128 rows per block, 16 per warp, lanes sweep along k
    const int lane = tid & 31, warp = tid >> 5;
    for (int i = 0; i < 16; i++) {
      const int64_t r = (int64_t)blockIdx.x * 128 + warp * 16 + i;
      int e = 0;
      for (int64_t kk = k0 + lane; kk < k1; kk += 32) {
        unsigned hi = (unsigned)__double2hiint(g[kk + r * ld]);
        int ex = (int)((hi >> 20) & 0x7ffu);
        e = ex > e ? ex : e;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        int t = __shfl_xor_sync(0xffffffffu, e, o);
        e = t > e ? t : e;
      }
      if (lane == 0) atomicMax(&emax[r], e);
    }
  }
}

constexpr int OZ_SL_ROWS = 32;            // rows per slicing block
constexpr int OZ_SL_STRIDE = OZ_BK + 4;   // 132-byte smem row stride: conflict-free byte scatter for both layouts

// Destination of a slicing pass in the block-cyclic one-sweep mode (dist.cu): the operand is a stack of nb-row blocks;
// block b goes to slot slot0 + b * slot_stride of the slot buffer (S planes of nb x nb int8, then nb row scales; one
// slot = slot_rows rows of nb bytes).  nb == 0: the plain [slice][row][k] layout.
struct OzSlotDst {
  int nb, slot0, slot_stride, slot_rows;
};

template <int S>
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* __restrict__ g, int64_t ld, int kc, int64_t Rpad,
                                                      int64_t Kpad, const int* __restrict__ emax,
                                                      int8_t* __restrict__ out, double* __restrict__ scale, int tri,
                                                      const OzSlotDst sd) {
  __shared__ __align__(16) int8_t sm[S * OZ_SL_ROWS * OZ_SL_STRIDE];
  __shared__ double s_inv[OZ_SL_ROWS];
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * OZ_SL_ROWS;
  const int64_t k0 = (int64_t)blockIdx.y * OZ_BK;
  // k-blocks in the zero part of a triangular operand are neither read nor written (the GEMM never loads them);
  // the block with blockIdx.y == the row block's own index is always inside the valid range and writes the scales
  const int64_t rb = r0 / OZ_BM;
  const bool scale_writer = tri ? ((int64_t)blockIdx.y == rb) : (blockIdx.y == 0);
  if ((tri > 0 && (int64_t)blockIdx.y < rb) || (tri < 0 && (int64_t)blockIdx.y > rb)) return;
  if (tid < OZ_SL_ROWS) {
    // |x| < 2^(E-1022) for every x of the row (E = largest biased exponent); x' = x * 2^-(E-1022) in (-1, 1)
    int E = emax[r0 + tid];
    double inv, sc;
    if (E >= 2047) {  // Inf / NaN in the row: poison the row's outputs instead of silently slicing garbage
      inv = 0.0;
      sc = __longlong_as_double(0x7ff8000000000000LL);
    } else {
      int Ec = E < 2 ? 2 : (E > 2044 ? 2044 : E);
      inv = __longlong_as_double((long long)(2045 - Ec) << 52);  // 2^-(Ec-1022)
      sc = __longlong_as_double((long long)(Ec + 1) << 52);      // 2^(Ec-1022)
    }
    s_inv[tid] = inv;
    if (sd.nb) {
      const int64_t b = r0 / sd.nb, rr = r0 - b * sd.nb;
      if (blockIdx.y == 0)
        reinterpret_cast<double*>(out + (((int64_t)sd.slot0 + b * sd.slot_stride) * sd.slot_rows + (int64_t)S * sd.nb) * sd.nb)[rr + tid] = sc;
    } else if (scale_writer) {
      scale[r0 + tid] = sc;
    }
  }
  __syncthreads();
#pragma unroll 4
  for (int i = 0; i < (OZ_SL_ROWS * OZ_BK) / 256; i++) {
    int r, kk;
    double x;
    if (!kc) {
      r = tid & 31;
      kk = (tid >> 5) + 8 * i;
      x = g[(r0 + r) + (k0 + kk) * ld];
    } else {
      kk = tid & 127;
      r = (tid >> 7) + 2 * i;
      x = g[(k0 + kk) + (r0 + r) * ld];
    }
    // first digit: 6 bits + sign (|x'| < 1), every further digit a full balanced byte: x' = d_0 2^-6 + d_1 2^-14 + ...
    // rint leaves |remainder| <= 1/2, so a raw digit can reach +128: carry it into the next more significant digit
    double v = x * s_inv[r] * 64.0;
    int8_t* dst = sm + r * OZ_SL_STRIDE + kk;
    int dg[S];
    // round to nearest even by adding 1.5 * 2^52 (|v| <= 128): the integer is the low word of the sum, the rounded value
    // the sum minus the constant -- two full-rate DADDs instead of F2F.RNI + F2I on the conversion unit (16 lanes per clock
    // per SM: with 16 conversions per element the kernel was bound by that unit at half the HBM rate)
    const double RN = 6755399441055744.0;
#pragma unroll
    for (int p = 0; p < S; p++) {
      const double t = v + RN;
      const double d = t - RN;
      dg[p] = __double2loint(t);
      v = (v - d) * 256.0;
    }
#pragma unroll
    for (int p = S - 1; p >= 1; p--) {
      if (dg[p] > 127) {
        dg[p] -= 256;
        dg[p - 1] += 1;
      }
    }
#pragma unroll
    for (int p = 0; p < S; p++) dst[p * (OZ_SL_ROWS * OZ_SL_STRIDE)] = (int8_t)dg[p];
  }
  __syncthreads();
  // write out: S * 32 rows of 128 contiguous bytes, 16 B per thread
  for (int c = tid; c < S * OZ_SL_ROWS * (OZ_BK / 16); c += 256) {
    const int p = c / (OZ_SL_ROWS * 8), rem = c % (OZ_SL_ROWS * 8);
    const int r = rem >> 3, ch = rem & 7;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(sm + p * (OZ_SL_ROWS * OZ_SL_STRIDE) + r * OZ_SL_STRIDE + ch * 16);
    uint4 v4 = make_uint4(src[0], src[1], src[2], src[3]);
    if (sd.nb) {
      const int64_t b = r0 / sd.nb, rr = r0 - b * sd.nb;
      const int64_t row = ((int64_t)sd.slot0 + b * sd.slot_stride) * sd.slot_rows + (int64_t)p * sd.nb + rr + r;
      *reinterpret_cast<uint4*>(out + row * sd.nb + k0 + ch * 16) = v4;
    } else {
      *reinterpret_cast<uint4*>(out + ((int64_t)p * Rpad + r0 + r) * Kpad + k0 + ch * 16) = v4;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// The tcgen05 GEMM
// ------------------------------------------------------------------------------------------------------
struct OzArgs {
  double* C;
  int64_t ldc;
  const double* scaleA;  // 2^ea_i per row of A
  const double* scaleB;  // 2^eb_j per row of B (= column of C)
  double alpha, beta;
  int tiles_m, tiles_n;
  int kblocks;           // K / 128
  int S, NA;
  int lower;             // skip tiles strictly above the diagonal
  int a_tri, b_tri;      // triangular operands: +1 zero for kk < row/col (k loop starts there), -1 zero for kk > row/col
  int kb_lo, kb_hi;      // k-block range of this launch (k is chunked so that the int32 levels cannot overflow)
  int tile_base;         // first linear tile index of this launch (wave-by-wave launches, GemmCall::sm_limit)
  int RpadA, RpadB;      // padded row counts (slice stride in the tensor maps)
  int* errflag;
  // ---- block-cyclic one-sweep mode (dist.cu): C is the LOCAL part of a matrix distributed in nb x nb blocks over a
  // P x Q process grid (local block (il, jl) = global block (il P + p, jl Q + q)); both operands are block rows of ONE
  // pre-sliced panel held in "slots" (slot g = global block g: S planes of nb x nb int8, then nb row scales).  Per tile:
  //   skipped unless global block row >= global block column (and neither lies in [skip_lo, skip_hi]);
  //   C = beta' C + alpha' A B',  alpha' = -1 below the panel's step row (i > kstep), +1 at / above it,
  //                               beta' = 0 in block row kstep and block column kstep, 1 elsewhere.
  int cyc;               // 0: off
  int nb, P, Q, p, q;    // block size (multiple of 128) and grid position
  int kstep;             // global block index of the current panel
  int skip_lo, skip_hi;  // global block rows AND columns in [skip_lo, skip_hi] are left out (-1, -1: none): look-ahead launches
  int bm0, bn0;          // tile offsets of the launched sub-rectangle inside the local matrix
  int slot_rows;         // rows (of nb bytes) per slot = S * nb + 8
  const uint8_t* slots;  // base of the slot buffer (for the row scales)
  long long* dbg;        // non-null: clock64 phase stamps of CTA dbg_cta (tools/oz_stamps.py), 8 entries
  int dbg_cta;
};

__device__ __forceinline__ uint32_t oz_elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred;
}
// Shared-memory operand descriptor of a K-major tile with 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart
// (SBO), descriptor version 1.  Constant high word (SBO >> 4 @bit 32, version @bit 46, layout SWIZZLE_128B = 2 @bit 61)
// | low word (start address >> 4, LBO field = 1).
// Instruction descriptor: D = S32 (2 @bit 4), A = B = signed int8 (1 @bit 7, 1 @bit 10), both K-major, N >> 3 @bit 17,
// M >> 4 @bit 24.
constexpr uint32_t OZ_DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t oz_desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void oz_mma_i8_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(OZ_DESC_HI)
      : "memory");
}
__host__ __device__ constexpr uint32_t oz_idesc_c(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
}

// all MMAs of one A slice (p) against its B slices for one k-block: 4 k-steps x ceil((S-p)/4) instructions
template <int S, int P>
__device__ __forceinline__ void oz_issue_slice(uint32_t tmem, uint32_t a_lo, uint32_t b_lo, uint32_t first_acc) {
#pragma unroll
  for (int ks = 0; ks < OZ_BK / OZ_UK; ks++) {
#pragma unroll
    for (int q0 = 0; q0 < S - P; q0 += 4) {
      const int nq = (S - P - q0) < 4 ? (S - P - q0) : 4;
      const uint32_t acc = (P == 0 && ks == 0) ? first_acc : 1u;  // A slice 0 touches every level first
      oz_mma_i8_lo(tmem + (uint32_t)(P + q0) * OZ_BN, a_lo + ks * (OZ_UK >> 4), b_lo + q0 * (OZ_B_BYTES >> 4) + ks * (OZ_UK >> 4),
                   oz_idesc_c(nq * OZ_BN), acc);
    }
  }
}

template <int S, int NA>
__global__ void __launch_bounds__(OZ_THREADS, 1)
    oz_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const OzArgs a) {
  extern __shared__ uint8_t oz_smem_raw[];
  // ---- which tile (grouped raster: 8 row tiles share the B panels that stream through L2)
  int bm, bn;
  {
    const int GM = 8;
    const int t = (int)blockIdx.x + a.tile_base;
    const int per_group = GM * a.tiles_n;
    const int grp = t / per_group, rem = t % per_group;
    const int first = grp * GM;
    const int gsz = (a.tiles_m - first) < GM ? (a.tiles_m - first) : GM;
    bm = first + rem % gsz;
    bn = rem / gsz;
  }
  const bool dbg_on = a.dbg != nullptr && (int)blockIdx.x == a.dbg_cta;
  if (dbg_on && threadIdx.x == 0) a.dbg[0] = clock64();  // 0: CTA start
  if (a.lower && bn > 2 * bm + 1) return;  // whole CTA, before any barrier / TMEM allocation
  // operand row coordinates of slice 0 in the tensor maps, slice-to-slice stride, scales, and the per-tile alpha / beta
  int a_row0, b_row0, a_ss = a.RpadA, b_ss = a.RpadB;
  const double* scaleA = a.scaleA;
  const double* scaleB = a.scaleB;
  double alpha_t = a.alpha, beta_t = a.beta;
  if (a.cyc) {
    bm += a.bm0;
    bn += a.bn0;
    const int il = (bm * OZ_BM) / a.nb, jl = (bn * OZ_BN) / a.nb;
    const int gi = il * a.P + a.p, gj = jl * a.Q + a.q;
    if (gi < gj || (gi >= a.skip_lo && gi <= a.skip_hi) || (gj >= a.skip_lo && gj <= a.skip_hi)) return;
    const int ri = bm * OZ_BM - il * a.nb, rj = bn * OZ_BN - jl * a.nb;
    a_row0 = gi * a.slot_rows + ri;
    b_row0 = gj * a.slot_rows + rj;
    a_ss = b_ss = a.nb;
    scaleA = reinterpret_cast<const double*>(a.slots + ((int64_t)gi * a.slot_rows + (int64_t)a.S * a.nb) * a.nb) + ri;
    scaleB = reinterpret_cast<const double*>(a.slots + ((int64_t)gj * a.slot_rows + (int64_t)a.S * a.nb) * a.nb) + rj;
    alpha_t = (gi > a.kstep) ? -1.0 : 1.0;
    beta_t = (gi == a.kstep || gj == a.kstep) ? 0.0 : 1.0;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = bm * OZ_BM, n0 = bn * OZ_BN;
  if (!a.cyc) {
    a_row0 = m0;
    b_row0 = n0;
    scaleA += m0;
    scaleB += n0;
  }
  int kb0 = a.kb_lo, KB = a.kb_hi;
  if (a.a_tri > 0 && m0 / OZ_BK > kb0) kb0 = m0 / OZ_BK;
  if (a.b_tri > 0 && n0 / OZ_BK > kb0) kb0 = n0 / OZ_BK;
  if (a.a_tri < 0 && (m0 + OZ_BM + OZ_BK - 1) / OZ_BK < KB) KB = (m0 + OZ_BM + OZ_BK - 1) / OZ_BK;
  if (a.b_tri < 0 && (n0 + OZ_BN + OZ_BK - 1) / OZ_BK < KB) KB = (n0 + OZ_BN + OZ_BK - 1) / OZ_BK;
  if (kb0 >= KB) {
    // nothing to accumulate for this tile in this launch: C = beta * C
    if (beta_t == 1.0) return;
    if (tid < OZ_BM) {
      double* crow = a.C + (int64_t)(m0 + tid) + (int64_t)n0 * a.ldc;
      for (int j = 0; j < OZ_BN; j++) {
        double* p = crow + (int64_t)j * a.ldc;
        *p = (beta_t == 0.0) ? 0.0 : beta_t * (*p);
      }
    }
    return;
  }

  // ---- shared memory carve-up (1024-byte aligned for the 128B swizzle atoms)
  const uint32_t raw = smem_u32(oz_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = oz_smem_raw + (base - raw);
  const uint32_t sB = base;                                   // 2 sets x S x 8 KB
  const uint32_t sA = sB + 2u * S * OZ_B_BYTES;               // NA x 16 KB
  uint8_t* tail = gbase + 2u * S * OZ_B_BYTES + (uint32_t)NA * OZ_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);         // full_a[8] empty_a[8] full_b[2] empty_b[2] tmem_full
  double* s_cscale = reinterpret_cast<double*>(tail + 256);   // 64 doubles
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(tail + 256 + 512);
  const uint32_t bar0 = smem_u32(bars);
  auto FULL_A = [&](uint32_t i) { return bar0 + 8u * i; };
  auto EMPTY_A = [&](uint32_t i) { return bar0 + 8u * (OZ_MAXNA + i); };
  auto FULL_B = [&](uint32_t i) { return bar0 + 8u * (2 * OZ_MAXNA + i); };
  auto EMPTY_B = [&](uint32_t i) { return bar0 + 8u * (2 * OZ_MAXNA + 2 + i); };
  const uint32_t TMEM_FULL = bar0 + 8u * (2 * OZ_MAXNA + 4);

  if (warp == 4 && lane == 0) {
    for (int i = 0; i < OZ_MAXNA; i++) {
      oz_mbar_init(FULL_A(i), 1);
      oz_mbar_init(EMPTY_A(i), 1);
    }
    for (int i = 0; i < 2; i++) {
      oz_mbar_init(FULL_B(i), 1);
      oz_mbar_init(EMPTY_B(i), 1);
    }
    oz_mbar_init(TMEM_FULL, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    oz_prefetch_tmap(&tmA);
    oz_prefetch_tmap(&tmB);
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  oz_tc_fence_before();
  __syncthreads();
  oz_tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (dbg_on && threadIdx.x == 0) a.dbg[1] = clock64();  // 1: barriers initialised, TMEM allocated

  if (warp == 4) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t slot = 0, round = 0;
      for (int kb = kb0; kb < KB; kb++) {
        const int it = kb - kb0, set = it & 1;
        if (it >= 2) oz_mbar_wait(EMPTY_B(set), ((it >> 1) - 1) & 1, a.errflag, 1);
        oz_mbar_expect_tx(FULL_B(set), (uint32_t)S * OZ_B_BYTES);
#pragma unroll
        for (int q = 0; q < S; q++)
          oz_tma_load_2d(sB + (uint32_t)(set * S + q) * OZ_B_BYTES, &tmB, FULL_B(set), kb * OZ_BK, q * b_ss + b_row0);
#pragma unroll
        for (int p = 0; p < S; p++) {
          if (round >= 1) oz_mbar_wait(EMPTY_A(slot), (round - 1) & 1, a.errflag, 2);
          oz_mbar_expect_tx(FULL_A(slot), OZ_A_BYTES);
          oz_tma_load_2d(sA + slot * OZ_A_BYTES, &tmA, FULL_A(slot), kb * OZ_BK, p * a_ss + a_row0);
          if (++slot == NA) {
            slot = 0;
            round++;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===== MMA issuer: the whole warp walks the pipeline, one elected lane issues (keeps UTCIMMA in uniform control flow)
    uint32_t slot = 0, aphase = 0;
    const uint32_t a_lo0 = oz_desc_lo(sA), b_lo0 = oz_desc_lo(sB);
    for (int kb = kb0; kb < KB; kb++) {
      const int it = kb - kb0, set = it & 1;
      oz_mbar_wait(FULL_B(set), (it >> 1) & 1, a.errflag, 3);
      if (dbg_on && lane == 0 && it == 0) a.dbg[2] = clock64();  // 2: first operands have landed
      const uint32_t b_lo = b_lo0 + (uint32_t)set * (S * OZ_B_BYTES >> 4);
      const uint32_t first_acc = it == 0 ? 0u : 1u;
#define OZ_SLICE(P)                                                            \
  if (P < S) {                                                                 \
    oz_mbar_wait(FULL_A(slot), aphase, a.errflag, 4);                          \
    oz_tc_fence_after();                                                       \
    if (oz_elect_one()) {                                                      \
      oz_issue_slice<S, (P < S ? P : 0)>(tmem, a_lo0 + slot * (OZ_A_BYTES >> 4), b_lo, first_acc); \
      oz_tc_commit(EMPTY_A(slot));                                             \
    }                                                                          \
    __syncwarp();                                                              \
    if (++slot == NA) {                                                        \
      slot = 0;                                                                \
      aphase ^= 1;                                                             \
    }                                                                          \
  }
      OZ_SLICE(0) OZ_SLICE(1) OZ_SLICE(2) OZ_SLICE(3) OZ_SLICE(4) OZ_SLICE(5) OZ_SLICE(6) OZ_SLICE(7)
#undef OZ_SLICE
      if (oz_elect_one()) oz_tc_commit(EMPTY_B(set));
      __syncwarp();
    }
    if (oz_elect_one()) oz_tc_commit(TMEM_FULL);
    __syncwarp();
    if (dbg_on && lane == 0) a.dbg[3] = clock64();  // 3: last MMA issued
  } else {
    // ===== epilogue warps 0..3: TMEM lane = output row =====
    if (tid < OZ_BN) s_cscale[tid] = scaleB[tid];
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int row = warp * 32 + lane;
    const double rs = scaleA[row] * (1.0 / 4096.0);  // 2^-12: weight of level c = 2
    const double alpha = alpha_t, beta = beta_t;
    // the whole C row segment of this thread (64 doubles) is fetched while the tile is still being accumulated: a
    // load issued per chunk after the wait costs a DRAM round trip per chunk on every tile
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    double* crow = a.C + (int64_t)(m0 + row) + (int64_t)n0 * a.ldc;
    double cv[OZ_BN];
    if (beta != 0.0) {
#pragma unroll
      for (int j = 0; j < OZ_BN; j++) cv[j] = __ldcg(crow + (int64_t)j * a.ldc);
    } else {
#pragma unroll
      for (int j = 0; j < OZ_BN; j++) cv[j] = 0.0;
    }
    if (dbg_on && tid == 0) a.dbg[4] = clock64();  // 4: C row segment requested
    while (!oz_mbar_try_wait(TMEM_FULL, 0)) __nanosleep(256);  // stay out of the issuer's way while the tile is computed
    oz_tc_fence_after();
    if (dbg_on && tid == 0) a.dbg[5] = clock64();  // 5: accumulators complete
#pragma unroll
    for (int ch = 0; ch < OZ_BN / 16; ch++) {
      double acc[16];
      int r[16], r2[16];
      double* cp = crow + (int64_t)(ch * 16) * a.ldc;
      // Horner over the levels, most significant last: two levels per TMEM round trip
      int l = S - 1;
      if (S & 1) {
        oz_tmem_ld16(trow + (uint32_t)(l * OZ_BN + ch * 16), r);
#pragma unroll
        for (int j = 0; j < 16; j++) acc[j] = oz_i2d(r[j]);
        l--;
      } else {
        oz_tmem_ld16x2(trow + (uint32_t)(l * OZ_BN + ch * 16), trow + (uint32_t)((l - 1) * OZ_BN + ch * 16), r, r2);
#pragma unroll
        for (int j = 0; j < 16; j++) acc[j] = fma(oz_i2d(r[j]), 0.00390625, oz_i2d(r2[j]));
        l -= 2;
      }
#pragma unroll
      for (; l >= 1; l -= 2) {
        oz_tmem_ld16x2(trow + (uint32_t)(l * OZ_BN + ch * 16), trow + (uint32_t)((l - 1) * OZ_BN + ch * 16), r, r2);
#pragma unroll
        for (int j = 0; j < 16; j++) {
          acc[j] = fma(acc[j], 0.00390625, oz_i2d(r[j]));   // 2^-8 per level
          acc[j] = fma(acc[j], 0.00390625, oz_i2d(r2[j]));
        }
      }
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const double v = alpha * (acc[j] * rs * s_cscale[ch * 16 + j]);
        cp[(int64_t)j * a.ldc] = fma(beta, cv[ch * 16 + j], v);
      }
    }
  }
  if (dbg_on && tid == 0) a.dbg[6] = clock64();  // 6: epilogue stores issued (thread 0)
  oz_tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    oz_tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
  if (dbg_on && tid == 0) a.dbg[7] = clock64();  // 7: end
}

static long long* g_oz_dbg = nullptr;
static int g_oz_dbg_cta = 0;
void oz_set_debug(long long* dev_stamps, int cta) {
  g_oz_dbg = dev_stamps;
  g_oz_dbg_cta = cta;
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
constexpr size_t OZ_SMEM_FIXED = 1024 /*alignment slack*/ + 1024 /*barriers, column scales, TMEM address*/;
constexpr int oz_na(int S) {
  return (int)((232448 - OZ_SMEM_FIXED - 2 * (size_t)S * OZ_B_BYTES) / OZ_A_BYTES) > OZ_MAXNA
             ? OZ_MAXNA
             : (int)((232448 - OZ_SMEM_FIXED - 2 * (size_t)S * OZ_B_BYTES) / OZ_A_BYTES);
}
constexpr size_t oz_smem(int S) { return OZ_SMEM_FIXED + 2 * (size_t)S * OZ_B_BYTES + (size_t)oz_na(S) * OZ_A_BYTES; }

typedef CUresult (*PFN_tmap_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmap_encode get_encode() {
  static PFN_tmap_encode fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmap_encode)p;
  }
  return fn;
}

// int8 slices [S*Rpad rows][Kpad] K-major; box = 128 bytes of k x box_rows rows, 128B swizzle
static int make_tmap(CUtensorMap* tm, const int8_t* base, int64_t rows_total, int64_t Kpad, int box_rows) {
  PFN_tmap_encode enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return GPC_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows_total};
  cuuint64_t gstr[1] = {(cuuint64_t)Kpad};
  cuuint32_t box[2] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    return GPC_ERR_CUDA;
  }
  return GPC_OK;
}

// One workspace per device, shared by all streams: Ozaki GEMMs fill the whole GPU, so they are serialised
// through `done` instead of giving every side stream its own multi-GB slice buffers.
struct OzWorkspace {
  int8_t* sl[2] = {nullptr, nullptr};
  size_t cap[2] = {0, 0};
  int* emax[2] = {nullptr, nullptr};
  double* scale[2] = {nullptr, nullptr};
  size_t rcap[2] = {0, 0};
  int* errflag = nullptr;
  cudaEvent_t done = nullptr;
  bool used = false;
};
static std::mutex g_oz_mu;
constexpr int OZ_BIG = 4;  // large calls in flight: main (chain) stream, a side stream (T), the bulk stream of the row-block pipeline
static OzWorkspace g_oz_ws[64][OZ_BIG];
static unsigned g_oz_big_next[64];
// high-water marks (bytes of slices, rows) of the large workspaces of a device: a workspace that has to grow grows to
// the mark at once, so that after the first evaluation no call re-allocates (cudaFree synchronises the device), whichever
// workspace the idle-first selection hands it
static size_t g_oz_big_hw_bytes[64], g_oz_big_hw_rows[64];
// Calls whose slices fit OZ_SMALL_BYTES per operand take one of OZ_POOL fixed-size workspaces round-robin instead (each
// guarded by its own event), so that the concurrent branches of the recursions (fork/join side streams) do not
// serialise on the shared buffers.  The pool is allocated once per device: no allocation in steady state.
constexpr size_t OZ_SMALL_BYTES = (size_t)16 << 20;
constexpr size_t OZ_SMALL_ROWS = 32768;  // S*R*K <= 16 MB with K >= 128, S >= 2
constexpr int OZ_POOL = 8;
static OzWorkspace g_oz_pool[64][OZ_POOL];
static unsigned g_oz_pool_next[64];

static int g_oz_on = -1, g_oz_S = 8;
static int64_t g_oz_min_mn = 256, g_oz_min_k = 512;
static double g_oz_bias = 1.0;  // use the tensor-core engine when its modelled time is below bias x the DMMA time
static void oz_init_settings() {
  if (g_oz_on >= 0) return;
  const char* e = getenv("GPC_OZAKI");
  g_oz_on = e ? atoi(e) : 1;
  if ((e = getenv("GPC_OZAKI_SLICES"))) g_oz_S = atoi(e);
  if ((e = getenv("GPC_OZAKI_MIN_MN"))) g_oz_min_mn = atoll(e);
  if ((e = getenv("GPC_OZAKI_MIN_K"))) g_oz_min_k = atoll(e);
  if ((e = getenv("GPC_OZAKI_BIAS"))) g_oz_bias = atof(e);
  if (g_oz_S < 2) g_oz_S = 2;
  if (g_oz_S > OZ_MAXS) g_oz_S = OZ_MAXS;
}
void oz_configure(int on, int slices, int64_t min_mn, int64_t min_k) {
  oz_init_settings();
  if (on >= 0) g_oz_on = on;
  if (slices >= 2 && slices <= OZ_MAXS) g_oz_S = slices;
  if (min_mn > 0) g_oz_min_mn = min_mn;
  if (min_k > 0) g_oz_min_k = min_k;
}
// Engine choice per call: a wave-quantised cost model of both engines, calibrated on B200 (tools/oz_check.py perf,
// tools/oz_sweep.py): Ozaki = fixed launch cost + slicing traffic + waves of 128x64 tiles at 36 int8 MMAs per fp64 MMA;
// DMMA = waves of 64x64 tiles (small problems) or the measured large-problem rate.
static double oz_cost_us(const GemmCall& c, int S) {
  const double tm = (double)(c.m / OZ_BM), tn = (double)(c.n / OZ_BN);
  const double tiles = c.lower ? tm * (tm + 1.0) : tm * tn;  // lower: row block bm owns 2(bm+1) column tiles
  const double waves = ceil(tiles / 148.0);
  const bool same = (c.A == c.B && c.lda == c.ldb && c.a_kc == c.b_kc && c.m == c.n);
  const double elems = (double)c.k * (same ? (double)c.m : (double)(c.m + c.n));
  const double per_k = 0.0267 * (S * (S + 1)) / 72.0;  // us per unit of k per tile wave (36 products measured)
  const bool tri = c.ktri || c.a_tri || c.b_tri;
  const double kfrac = tri ? (c.lower ? 0.67 : 0.5) : 1.0;  // average share of the k range a tile walks
  return 20.0 + elems * (16.0 + S) / 3.0e6 + waves * ((double)c.k * kfrac * per_k + 5.0);
}
static double dmma_cost_us(const GemmCall& c) {
  const double flops = c.lower ? (double)c.m * (double)(c.m + TILE) * (double)c.k : 2.0 * (double)c.m * (double)c.n * (double)c.k;
  const double t64 = c.lower ? (double)(c.m / 64) * (double)(c.m / 64 + 1) / 2.0 : (double)(c.m / 64) * (double)(c.n / 64);
  const bool tri = c.ktri || c.a_tri || c.b_tri;
  const double kfrac = tri ? (c.lower ? 0.67 : 0.5) : 1.0;
  const double big = kfrac * flops / 33.0e6;                                  // 33 TFLOP/s sustained on many-wave problems
  const double small = ceil(t64 / 296.0) * (double)c.k * kfrac * 0.0655;     // 64x64 tiles, 2 CTAs per SM
  return 4.0 + (big > small ? big : small);
}
int oz_slices() {
  oz_init_settings();
  return g_oz_S;
}
bool oz_wants(const GemmCall& c) {
  oz_init_settings();
  if (!g_oz_on) return false;
  if (c.C == c.A || c.C == c.B) return false;
  if (c.m % OZ_BM || c.n % OZ_BN || c.k % OZ_BK) return false;
  if (c.m < g_oz_min_mn || c.n < g_oz_min_mn || c.k < g_oz_min_k) return false;
  if (c.m * (int64_t)OZ_MAXS >= (1LL << 31) || c.n * (int64_t)OZ_MAXS >= (1LL << 31)) return false;
  return oz_cost_us(c, g_oz_S) < g_oz_bias * dmma_cost_us(c);
}

static int ensure_ws(OzWorkspace& w, int which, size_t bytes, size_t rows) {
  if (!w.done) {
    GPC_CUDA_CHECK(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
    GPC_CUDA_CHECK(cudaMalloc(&w.errflag, sizeof(int)));
    GPC_CUDA_CHECK(cudaMemset(w.errflag, 0, sizeof(int)));
  }
  if (w.cap[which] < bytes) {
    if (w.sl[which]) GPC_CUDA_CHECK(cudaFree(w.sl[which]));  // implicit device synchronisation
    w.sl[which] = nullptr;
    w.cap[which] = 0;
    GPC_CUDA_CHECK(cudaMalloc(&w.sl[which], bytes));
    w.cap[which] = bytes;
  }
  if (w.rcap[which] < rows) {
    if (w.emax[which]) GPC_CUDA_CHECK(cudaFree(w.emax[which]));
    if (w.scale[which]) GPC_CUDA_CHECK(cudaFree(w.scale[which]));
    w.emax[which] = nullptr;
    w.scale[which] = nullptr;
    w.rcap[which] = 0;
    GPC_CUDA_CHECK(cudaMalloc(&w.emax[which], rows * sizeof(int)));
    GPC_CUDA_CHECK(cudaMalloc(&w.scale[which], rows * sizeof(double)));
    w.rcap[which] = rows;
  }
  return GPC_OK;
}

template <int S>
static void launch_slice_t(const double* g, int64_t ld, int kc, int64_t R, int64_t K, const int* emax, int8_t* out,
                           double* scale, cudaStream_t s, int tri, OzSlotDst sd = OzSlotDst{0, 0, 0, 0}) {
  dim3 grid((unsigned)(R / OZ_SL_ROWS), (unsigned)(K / OZ_BK));
  oz_slice_kernel<S><<<grid, 256, 0, s>>>(g, ld, kc, R, K, emax, out, scale, tri, sd);
}

static int slice_operand(OzWorkspace& w, int which, const double* g, int64_t ld, bool kc, int64_t R, int64_t K, int S,
                         cudaStream_t s, int64_t* launches, int tri = 0) {
  GPC_CHECK(ensure_ws(w, which, (size_t)S * R * K, (size_t)R));
  GPC_CUDA_CHECK(cudaMemsetAsync(w.emax[which], 0, R * sizeof(int), s));
  // enough blocks to cover the GPU a few times: (R/128) x kchunks >= ~600, chunks of >= 64 columns
  int64_t kchunks = (600 + R / 128 - 1) / (R / 128);
  if (kchunks > K / 64) kchunks = K / 64;
  if (kchunks < 1) kchunks = 1;
  int64_t kchunk = (K + kchunks - 1) / kchunks;
  dim3 g1((unsigned)(R / 128), (unsigned)kchunks);
  oz_rowmax_kernel<<<g1, 256, 0, s>>>(g, ld, kc ? 1 : 0, K, kchunk, w.emax[which], tri);
  GPC_CUDA_CHECK(cudaGetLastError());
  switch (S) {
    case 2: launch_slice_t<2>(g, ld, kc, R, K, w.emax[which], w.sl[which], w.scale[which], s, tri); break;
    case 3: launch_slice_t<3>(g, ld, kc, R, K, w.emax[which], w.sl[which], w.scale[which], s, tri); break;
    case 4: launch_slice_t<4>(g, ld, kc, R, K, w.emax[which], w.sl[which], w.scale[which], s, tri); break;
    case 5: launch_slice_t<5>(g, ld, kc, R, K, w.emax[which], w.sl[which], w.scale[which], s, tri); break;
    case 6: launch_slice_t<6>(g, ld, kc, R, K, w.emax[which], w.sl[which], w.scale[which], s, tri); break;
    case 7: launch_slice_t<7>(g, ld, kc, R, K, w.emax[which], w.sl[which], w.scale[which], s, tri); break;
    default: launch_slice_t<8>(g, ld, kc, R, K, w.emax[which], w.sl[which], w.scale[which], s, tri); break;
  }
  GPC_CUDA_CHECK(cudaGetLastError());
  if (launches) (*launches) += 2;
  if (trace_sync("oz_slice_kernel", s) != GPC_OK) return GPC_ERR_CUDA;
  return GPC_OK;
}

// Wave-by-wave launches (GemmCall::sm_limit): cut the linear tile range of oz_gemm_kernel (grouped raster: groups of 8
// row tiles, column by column; tile t -> group t / (8 tiles_n), column (t % (8 tiles_n)) / gsz, row first + rem % gsz) at
// column boundaries so that no launch holds more than `limit` NON-EMPTY tiles (lower mode: a tile (bm, bn) exists iff
// bn <= 2 bm + 1; the others exit at once).  Returns the launch boundaries, first 0, last tiles_m * tiles_n.
std::vector<int64_t> oz_wave_cuts(int tiles_m, int tiles_n, bool lower, int limit) {
  std::vector<int64_t> cuts;
  cuts.push_back(0);
  const int64_t ntiles = (int64_t)tiles_m * tiles_n;
  if (limit > 0 && ntiles > limit) {
    const int GM = 8;
    int64_t live = 0;
    for (int first = 0; first < tiles_m; first += GM) {
      const int gsz = (tiles_m - first) < GM ? (tiles_m - first) : GM;
      const int64_t gbase = (int64_t)(first / GM) * GM * tiles_n;
      for (int bn = 0; bn < tiles_n; bn++) {
        int cnt = gsz;
        if (lower) {
          cnt = 0;
          for (int r = 0; r < gsz; r++) cnt += (bn <= 2 * (first + r) + 1) ? 1 : 0;
        }
        if (live + cnt > limit && live > 0) {
          cuts.push_back(gbase + (int64_t)bn * gsz);
          live = 0;
        }
        live += cnt;
      }
    }
  }
  cuts.push_back(ntiles);
  return cuts;
}

int launch_gemm_ozaki(const GemmCall& c, cudaStream_t s, int64_t* launches, int slices) {
  oz_init_settings();
  int dev = 0;
  GPC_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) {
    set_error("launch_gemm_ozaki: device index out of range");
    return GPC_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(g_oz_mu);
  const int S = (slices >= 2 && slices <= OZ_MAXS) ? slices : g_oz_S;
  const size_t op_bytes = (size_t)S * (size_t)(c.m > c.n ? c.m : c.n) * (size_t)c.k;
  const bool shared_ws = op_bytes > OZ_SMALL_BYTES;
  // the second large workspace only for operands up to 1 GB of slices: beyond that (N >= 32768) large calls are
  // serialised through one workspace, which keeps N = 65536 within one GPU's memory
  const bool rotate_big = op_bytes <= ((size_t)1 << 30);
  // a workspace whose previous user has finished is preferred over plain rotation: a call on the (high-priority) chain
  // stream must not queue behind a long product of the bulk stream just because the rotation pointed at its buffers
  // ... and among the idle ones a workspace that is already large enough (growing one is a cudaFree + cudaMalloc, i.e.
  // a device synchronisation in the middle of an evaluation)
  auto pick = [&](OzWorkspace* ws, int count, unsigned& next) -> OzWorkspace& {
    int idle_any = -1;
    for (int t = 0; t < count; t++) {
      const int i = (int)((next + t) % count);
      OzWorkspace& c = ws[i];
      const bool idle = !c.used || cudaEventQuery(c.done) == cudaSuccess;
      if (!idle) {
        cudaGetLastError();  // cudaErrorNotReady is not an error
        continue;
      }
      if (c.cap[0] >= op_bytes && c.cap[1] >= op_bytes) {
        next = (unsigned)((i + 1) % count);
        return c;
      }
      if (idle_any < 0) idle_any = i;
    }
    if (idle_any >= 0) {
      next = (unsigned)((idle_any + 1) % count);
      return ws[idle_any];
    }
    return ws[next++ % count];
  };
  OzWorkspace& w = shared_ws ? (rotate_big ? pick(g_oz_ws[dev], OZ_BIG, g_oz_big_next[dev]) : g_oz_ws[dev][0])
                             : pick(g_oz_pool[dev], OZ_POOL, g_oz_pool_next[dev]);
  if (!shared_ws && !w.sl[0]) {  // first small call on this device: allocate the whole pool now, not over 8 calls
    for (int q = 0; q < OZ_POOL; q++)
      for (int i = 0; i < 2; i++) GPC_CHECK(ensure_ws(g_oz_pool[dev][q], i, OZ_SMALL_BYTES, OZ_SMALL_ROWS));
  }
  if (c.m % OZ_BM || c.n % OZ_BN || c.k % OZ_BK || c.C == c.A || c.C == c.B) {
    set_error("launch_gemm_ozaki: needs m % 128 == n % 64 == k % 128 == 0 and C distinct from A, B");
    return GPC_ERR_ARG;
  }
  // serialise against the previous Ozaki GEMM (possibly on another stream): the slice buffers are shared
  if (w.used) GPC_CUDA_CHECK(cudaStreamWaitEvent(s, w.done, 0));
  if (shared_ws && rotate_big) {
    const size_t rows = (size_t)(c.m > c.n ? c.m : c.n);
    if (op_bytes > g_oz_big_hw_bytes[dev]) g_oz_big_hw_bytes[dev] = op_bytes;
    if (rows > g_oz_big_hw_rows[dev]) g_oz_big_hw_rows[dev] = rows;
    for (int i = 0; i < 2; i++) GPC_CHECK(ensure_ws(w, i, g_oz_big_hw_bytes[dev], g_oz_big_hw_rows[dev]));
  }
  const bool same = (c.A == c.B && c.lda == c.ldb && c.a_kc == c.b_kc && c.m == c.n);
  GPC_CHECK(slice_operand(w, 0, c.A, c.lda, c.a_kc, c.m, c.k, S, s, launches, c.ktri ? 1 : c.a_tri));
  if (!same) GPC_CHECK(slice_operand(w, 1, c.B, c.ldb, c.b_kc, c.n, c.k, S, s, launches, c.b_tri));
  const int wb = same ? 0 : 1;
  CUtensorMap tmA, tmB;
  GPC_CHECK(make_tmap(&tmA, w.sl[0], (int64_t)S * c.m, c.k, OZ_BM));
  GPC_CHECK(make_tmap(&tmB, w.sl[wb], (int64_t)S * c.n, c.k, OZ_BN));
  OzArgs a;
  memset(&a, 0, sizeof(a));
  a.C = c.C;
  a.ldc = c.ldc;
  a.scaleA = w.scale[0];
  a.scaleB = w.scale[wb];
  a.alpha = c.alpha;
  a.beta = c.beta;
  a.tiles_m = (int)(c.m / OZ_BM);
  a.tiles_n = (int)(c.n / OZ_BN);
  a.kblocks = (int)(c.k / OZ_BK);
  a.S = S;
  a.NA = oz_na(S);
  a.lower = c.lower ? 1 : 0;
  a.a_tri = c.ktri ? 1 : c.a_tri;
  a.b_tri = c.b_tri;
  a.RpadA = (int)c.m;
  a.RpadB = (int)c.n;
  a.errflag = w.errflag;
  a.dbg = g_oz_dbg;
  a.dbg_cta = g_oz_dbg_cta;
  const int64_t ntiles = (int64_t)a.tiles_m * a.tiles_n;
  // k chunks: per unit of k a level sums <= 8 products, two of them with a first digit (|d_0| <= 65), the others with
  // |d| <= 128: (6 * 2^14 + 2 * 65 * 128) * k < 2^31  ->  k <= 18683: chunks of 128 k-blocks (16384)
  const int kchunk = 128;
  const double beta0 = c.beta;
  const std::vector<int64_t> cuts = oz_wave_cuts(a.tiles_m, a.tiles_n, a.lower != 0, c.sm_limit);
  for (int kb = 0; kb < a.kblocks; kb += kchunk) {
    a.kb_lo = kb;
    a.kb_hi = (kb + kchunk < a.kblocks) ? kb + kchunk : a.kblocks;
    a.beta = (kb == 0) ? beta0 : 1.0;
   for (size_t ci = 0; ci + 1 < cuts.size(); ci++) {
    a.tile_base = (int)cuts[ci];
    const int64_t nlaunch = cuts[ci + 1] - cuts[ci];
    if (nlaunch <= 0) continue;
    switch (S) {
#define OZ_CASE(SS)                                                                                              \
  case SS: {                                                                                                     \
    static bool configured_dev[64] = {false};                                                                              \
    auto kern = oz_gemm_kernel<SS, oz_na(SS)>;                                                                   \
    if (!configured_dev[cur_device() & 63]) {                                                                                           \
      GPC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oz_smem(SS))); \
      configured_dev[cur_device() & 63] = true;                                                                                         \
    }                                                                                                            \
    kern<<<(unsigned)nlaunch, OZ_THREADS, oz_smem(SS), s>>>(tmA, tmB, a);                                        \
  } break;
      OZ_CASE(2) OZ_CASE(3) OZ_CASE(4) OZ_CASE(5) OZ_CASE(6) OZ_CASE(7) OZ_CASE(8)
#undef OZ_CASE
      default:
        set_error("launch_gemm_ozaki: slices out of range");
        return GPC_ERR_ARG;
    }
    if (launches) (*launches)++;
   }
  }
  GPC_CUDA_CHECK(cudaGetLastError());
  GPC_CUDA_CHECK(cudaEventRecord(w.done, s));
  w.used = true;
  if (trace_sync("oz_gemm_kernel", s) != GPC_OK) return GPC_ERR_CUDA;
  return GPC_OK;
}

// ------------------------------------------------------------------------------------------------------
// Block-cyclic one-sweep mode (dist.cu).  The panel of a step lives in a slot buffer: slot g (global block g) =
// S planes of nb x nb int8 (k contiguous) followed by nb row scales = slot_rows = S nb + 8 rows of nb bytes.
// ------------------------------------------------------------------------------------------------------
size_t oz_slot_bytes(int nb, int S) { return (size_t)(S * nb + 8) * (size_t)nb; }

// slices the stacked operand g (R = nblocks * nb rows, nb deep; kc as in slice_operand) into slots slot0 + b * stride
int oz_slice_to_slots(const double* g, int64_t ld, bool kc, int64_t R, int nb, int S, int* emax_scratch, uint8_t* slots,
                      int slot0, int slot_stride, cudaStream_t s, int64_t* launches) {
  if (R <= 0) return GPC_OK;
  if (R % nb || nb % OZ_BK || S < 2 || S > OZ_MAXS) {
    set_error("oz_slice_to_slots: bad shape");
    return GPC_ERR_ARG;
  }
  GPC_CUDA_CHECK(cudaMemsetAsync(emax_scratch, 0, R * sizeof(int), s));
  int64_t kchunks = (600 + R / 128 - 1) / (R / 128);
  if (kchunks > nb / 64) kchunks = nb / 64;
  if (kchunks < 1) kchunks = 1;
  const int64_t kchunk = (nb + kchunks - 1) / kchunks;
  dim3 g1((unsigned)(R / 128), (unsigned)kchunks);
  oz_rowmax_kernel<<<g1, 256, 0, s>>>(g, ld, kc ? 1 : 0, nb, kchunk, emax_scratch, 0);
  GPC_CUDA_CHECK(cudaGetLastError());
  OzSlotDst sd{nb, slot0, slot_stride, S * nb + 8};
  int8_t* out = reinterpret_cast<int8_t*>(slots);
  switch (S) {
    case 2: launch_slice_t<2>(g, ld, kc, R, nb, emax_scratch, out, nullptr, s, 0, sd); break;
    case 3: launch_slice_t<3>(g, ld, kc, R, nb, emax_scratch, out, nullptr, s, 0, sd); break;
    case 4: launch_slice_t<4>(g, ld, kc, R, nb, emax_scratch, out, nullptr, s, 0, sd); break;
    case 5: launch_slice_t<5>(g, ld, kc, R, nb, emax_scratch, out, nullptr, s, 0, sd); break;
    case 6: launch_slice_t<6>(g, ld, kc, R, nb, emax_scratch, out, nullptr, s, 0, sd); break;
    case 7: launch_slice_t<7>(g, ld, kc, R, nb, emax_scratch, out, nullptr, s, 0, sd); break;
    default: launch_slice_t<8>(g, ld, kc, R, nb, emax_scratch, out, nullptr, s, 0, sd); break;
  }
  GPC_CUDA_CHECK(cudaGetLastError());
  if (launches) (*launches) += 2;
  if (trace_sync("oz_slice_kernel(slots)", s) != GPC_OK) return GPC_ERR_CUDA;
  return GPC_OK;
}

int oz_cyc_maps(OzCycMaps* out, const uint8_t* slots, int nslots, int nb, int S) {
  CUtensorMap* tm = reinterpret_cast<CUtensorMap*>(out->opaque);
  static_assert(sizeof(out->opaque) >= 2 * sizeof(CUtensorMap) + 64, "OzCycMaps too small");
  // 64-byte aligned placement inside the opaque storage
  uintptr_t a = (reinterpret_cast<uintptr_t>(out->opaque) + 63) & ~(uintptr_t)63;
  tm = reinterpret_cast<CUtensorMap*>(a);
  const int64_t rows = (int64_t)nslots * (S * nb + 8);
  GPC_CHECK(make_tmap(&tm[0], reinterpret_cast<const int8_t*>(slots), rows, nb, OZ_BM));
  GPC_CHECK(make_tmap(&tm[1], reinterpret_cast<const int8_t*>(slots), rows, nb, OZ_BN));
  out->slots = slots;
  out->nb = nb;
  out->S = S;
  return GPC_OK;
}

// C (local part, ld ldc) updated with the panel in `maps` over the tile sub-rectangle rows [r0, r0 + m), columns
// [c0, c0 + n) of the local matrix (multiples of 128 / 64); see OzArgs for the per-tile rule.
int launch_oz_cyc_update(const OzCycMaps& maps, const OzCycGrid& gr, int kstep, int skip_lo, int skip_hi, double* C,
                         int64_t ldc, int64_t r0, int64_t m, int64_t c0, int64_t n, int* errflag, cudaStream_t s,
                         int64_t* launches) {
  if (m <= 0 || n <= 0) return GPC_OK;
  const int S = maps.S, nb = maps.nb;
  if (r0 % OZ_BM || m % OZ_BM || c0 % OZ_BN || n % OZ_BN || nb % OZ_BK || (int64_t)nb / OZ_BK > 128) {
    set_error("launch_oz_cyc_update: bad tile range / block size");
    return GPC_ERR_ARG;
  }
  uintptr_t al = (reinterpret_cast<uintptr_t>(maps.opaque) + 63) & ~(uintptr_t)63;
  const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(al);
  OzArgs a;
  memset(&a, 0, sizeof(a));
  a.C = C;
  a.ldc = ldc;
  a.alpha = -1.0;
  a.beta = 1.0;
  a.tiles_m = (int)(m / OZ_BM);
  a.tiles_n = (int)(n / OZ_BN);
  a.kblocks = nb / OZ_BK;
  a.S = S;
  a.NA = oz_na(S);
  a.kb_lo = 0;
  a.kb_hi = a.kblocks;
  a.errflag = errflag;
  a.cyc = 1;
  a.nb = nb;
  a.P = gr.P;
  a.Q = gr.Q;
  a.p = gr.p;
  a.q = gr.q;
  a.kstep = kstep;
  a.skip_lo = skip_lo < 0 ? 0x7fffffff : skip_lo;  // empty range when there is nothing to leave out
  a.skip_hi = skip_lo < 0 ? -1 : skip_hi;
  a.bm0 = (int)(r0 / OZ_BM);
  a.bn0 = (int)(c0 / OZ_BN);
  a.slot_rows = S * nb + 8;
  a.slots = maps.slots;
  const int64_t ntiles = (int64_t)a.tiles_m * a.tiles_n;
  switch (S) {
#define OZ_CASE(SS)                                                                                              \
  case SS: {                                                                                                     \
    static bool configured_dev[64] = {false};                                                                    \
    auto kern = oz_gemm_kernel<SS, oz_na(SS)>;                                                                   \
    if (!configured_dev[cur_device() & 63]) {                                                                    \
      GPC_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oz_smem(SS))); \
      configured_dev[cur_device() & 63] = true;                                                                  \
    }                                                                                                            \
    kern<<<(unsigned)ntiles, OZ_THREADS, oz_smem(SS), s>>>(tm[0], tm[1], a);                                     \
  } break;
    OZ_CASE(2) OZ_CASE(3) OZ_CASE(4) OZ_CASE(5) OZ_CASE(6) OZ_CASE(7) OZ_CASE(8)
#undef OZ_CASE
    default:
      set_error("launch_oz_cyc_update: slices out of range");
      return GPC_ERR_ARG;
  }
  if (launches) (*launches)++;
  GPC_CUDA_CHECK(cudaGetLastError());
  if (trace_sync("oz_gemm_kernel(cyc)", s) != GPC_OK) return GPC_ERR_CUDA;
  return GPC_OK;
}

// ------------------------------------------------------------------------------------------------------
// Measured peak of the INT8 tensor pipe (the denominator of the engine's roofline): one CTA per SM keeps issuing
// tcgen05.mma.kind::i8 M = 128, N = 256, K = 32 on shared-memory-resident operands (no TMA, no epilogue) into its 512
// TMEM columns.  Two variants: every instruction reads fresh A and B from shared memory exactly like the GEMM kernel's
// widest instruction (12 KB per 1.05 M MACs); ops = CTAs x instructions x 2 x 128 x 256 x 32.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OZ_THREADS, 1) oz_imma_peak_kernel(int iters, int nwide, int* errflag) {
  extern __shared__ uint8_t oz_smem_raw[];
  const uint32_t raw = smem_u32(oz_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = oz_smem_raw + (base - raw);
  const uint32_t sA = base;                         // 16 KB: 128 rows x 128 B
  const uint32_t sB = base + OZ_A_BYTES;            // 32 KB: 256 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbase + OZ_A_BYTES + 4 * OZ_B_BYTES);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(gbase + OZ_A_BYTES + 4 * OZ_B_BYTES + 64);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // pseudo-random operand bytes: switching activity (and therefore power and the clock the power cap allows) like real slices
  for (int i = tid; i < (OZ_A_BYTES + 4 * OZ_B_BYTES) / 4; i += OZ_THREADS) {
    uint32_t x = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    x ^= x >> 15;
    x *= 2246822519u;
    x ^= x >> 13;
    reinterpret_cast<uint32_t*>(gbase)[i] = x;
  }
  const uint32_t bar0 = smem_u32(bars);
  if (warp == 4 && lane == 0) {
    oz_mbar_init(bar0, 1);
    oz_mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  oz_tc_fence_before();
  __syncthreads();
  oz_tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (warp == 5) {
    const uint32_t a_lo = oz_desc_lo(sA), b_lo = oz_desc_lo(sB);
    const int n = nwide * OZ_BN;  // 64 .. 256
    uint32_t ph[2] = {0, 0};
    for (int it = 0; it < iters; it++) {
      const int half = it & 1;
      if (it >= 2) {  // at most two batches in flight
        oz_mbar_wait(bar0 + 8 * half, ph[half], errflag, 9);
        ph[half] ^= 1;
      }
      oz_tc_fence_after();
      if (oz_elect_one()) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++)
#pragma unroll
          for (int ks = 0; ks < OZ_BK / OZ_UK; ks++)
            oz_mma_i8_lo(tmem + (uint32_t)half * 256, a_lo + ks * (OZ_UK >> 4), b_lo + ks * (OZ_UK >> 4), oz_idesc_c(n),
                         (it >= 2 || rep || ks) ? 1u : 0u);
        oz_tc_commit(bar0 + 8 * half);
      }
      __syncwarp();
    }
    for (int half = 0; half < 2; half++)
      if (iters > half) oz_mbar_wait(bar0 + 8 * half, ph[half], errflag, 10);
  }
  oz_tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    oz_tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

int oz_bench_imma_peak_sustained(int device, int nwide, double seconds, double* tops);
int oz_bench_imma_peak(int device, int nwide, double* tops) {
  GPC_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  GPC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (nwide < 1 || nwide > 4) nwide = 4;
  const size_t smem = 1024 + OZ_A_BYTES + 4 * OZ_B_BYTES + 256;
  GPC_CUDA_CHECK(cudaFuncSetAttribute(oz_imma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int* err = nullptr;
  GPC_CUDA_CHECK(cudaMalloc(&err, sizeof(int)));
  GPC_CUDA_CHECK(cudaMemset(err, 0, sizeof(int)));
  cudaStream_t s;
  GPC_CUDA_CHECK(cudaStreamCreate(&s));
  cudaEvent_t e0, e1;
  GPC_CUDA_CHECK(cudaEventCreate(&e0));
  GPC_CUDA_CHECK(cudaEventCreate(&e1));
  const int ctas = prop.multiProcessorCount, iters = 4096;
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    GPC_CUDA_CHECK(cudaEventRecord(e0, s));
    oz_imma_peak_kernel<<<ctas, OZ_THREADS, smem, s>>>(iters, nwide, err);
    GPC_CUDA_CHECK(cudaEventRecord(e1, s));
    GPC_CUDA_CHECK(cudaStreamSynchronize(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)ctas * iters * 8.0 * (OZ_BK / OZ_UK) * 2.0 * OZ_BM * (double)(nwide * OZ_BN) * OZ_UK;
    const double t = ops / (ms * 1e-3) / 1e12;
    if (rep > 0 && t > best) best = t;
  }
  int herr = 0;
  cudaMemcpy(&herr, err, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(err);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaStreamDestroy(s);
  if (herr) {
    set_error("oz_bench_imma_peak: pipeline protocol error");
    return GPC_ERR_CUDA;
  }
  if (tops) *tops = best;
  return GPC_OK;
}

int oz_bench_imma_peak_sustained(int device, int nwide, double seconds, double* tops) {
  GPC_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  GPC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (nwide < 1 || nwide > 4) nwide = 4;
  if (!(seconds > 0.05)) seconds = 0.05;
  if (seconds > 20.0) seconds = 20.0;
  const size_t smem = 1024 + OZ_A_BYTES + 4 * OZ_B_BYTES + 256;
  GPC_CUDA_CHECK(cudaFuncSetAttribute(oz_imma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int* err = nullptr;
  GPC_CUDA_CHECK(cudaMalloc(&err, sizeof(int)));
  GPC_CUDA_CHECK(cudaMemset(err, 0, sizeof(int)));
  cudaStream_t s;
  GPC_CUDA_CHECK(cudaStreamCreate(&s));
  const int ctas = prop.multiProcessorCount, iters = 4096;
  const double ops = (double)ctas * iters * 8.0 * (OZ_BK / OZ_UK) * 2.0 * OZ_BM * (double)(nwide * OZ_BN) * OZ_UK;
  const int launches = (int)(seconds / (ops / 4.0e15)) + 4;  // ~9 ms each at the burst rate
  const int half = launches / 2;
  cudaEvent_t e0, e1;
  GPC_CUDA_CHECK(cudaEventCreate(&e0));
  GPC_CUDA_CHECK(cudaEventCreate(&e1));
  for (int i = 0; i < launches; i++) {
    if (i == half) GPC_CUDA_CHECK(cudaEventRecord(e0, s));
    oz_imma_peak_kernel<<<ctas, OZ_THREADS, smem, s>>>(iters, nwide, err);
  }
  GPC_CUDA_CHECK(cudaEventRecord(e1, s));
  GPC_CUDA_CHECK(cudaStreamSynchronize(s));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  int herr = 0;
  cudaMemcpy(&herr, err, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(err);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaStreamDestroy(s);
  if (herr) {
    set_error("oz_bench_imma_peak_sustained: pipeline protocol error");
    return GPC_ERR_CUDA;
  }
  if (tops) *tops = ops * (launches - half) / (ms * 1e-3) / 1e12;
  return GPC_OK;
}

void oz_release_device(int dev) {
  if (dev < 0 || dev >= 64) return;
  std::lock_guard<std::mutex> lk(g_oz_mu);
  g_oz_big_hw_bytes[dev] = g_oz_big_hw_rows[dev] = 0;
  for (int b = 0; b < OZ_BIG; b++) {
    OzWorkspace& w = g_oz_ws[dev][b];
    for (int i = 0; i < 2; i++) {
      if (w.sl[i]) cudaFree(w.sl[i]);
      if (w.emax[i]) cudaFree(w.emax[i]);
      if (w.scale[i]) cudaFree(w.scale[i]);
    }
    if (w.errflag) cudaFree(w.errflag);
    if (w.done) cudaEventDestroy(w.done);
    w = OzWorkspace();
  }
  for (int q = 0; q < OZ_POOL; q++) {
    OzWorkspace& v = g_oz_pool[dev][q];
    for (int i = 0; i < 2; i++) {
      if (v.sl[i]) cudaFree(v.sl[i]);
      if (v.emax[i]) cudaFree(v.emax[i]);
      if (v.scale[i]) cudaFree(v.scale[i]);
    }
    if (v.errflag) cudaFree(v.errflag);
    if (v.done) cudaEventDestroy(v.done);
    v = OzWorkspace();
  }
}

}  // namespace gpc

// measured INT8 tensor-pipe peak of this device in TOP/s (ops = 2 x MACs); nwide = N / 64 of the instruction (1..4)
extern "C" int gpc_bench_imma_peak(int device, int nwide, double* tops) { return gpc::oz_bench_imma_peak(device, nwide, tops); }
// the same loop kept running for `seconds` (back-to-back launches): the SUSTAINED rate over the second half, i.e. under
// whatever clock the power cap leaves (the roofline of a kernel timed inside a long step); *mhz_hint is not measured here
extern "C" int gpc_bench_imma_peak_sustained(int device, int nwide, double seconds, double* tops) {
  return gpc::oz_bench_imma_peak_sustained(device, nwide, seconds, tops);
}

/* host-only: the launch boundaries of a wave-limited product (CPU test of the schedule); returns the number of
 * boundaries written (<= cap), or the number needed when cuts == NULL */
extern "C" int gpc_oz_wave_cuts(int tiles_m, int tiles_n, int lower, int limit, long long* cuts, int cap) {
  if (tiles_m < 1 || tiles_n < 1) return GPC_ERR_ARG;
  const std::vector<int64_t> v = gpc::oz_wave_cuts(tiles_m, tiles_n, lower != 0, limit);
  if (!cuts) return (int)v.size();
  int n = 0;
  for (; n < (int)v.size() && n < cap; n++) cuts[n] = (long long)v[(size_t)n];
  return n;
}

extern "C" int gpc_oz_slice_check(int device, int64_t R, int64_t K, int kc, int S, const double* X, signed char* slices_out,
                                  double* scale_out) {
  using namespace gpc;
  if (R % 128 || K % 128 || R < 128 || K < 128 || S < 2 || S > OZ_MAXS || !X || !slices_out || !scale_out) {
    set_error("gpc_oz_slice_check: R, K multiples of 128, S in 2..8");
    return GPC_ERR_ARG;
  }
  GPC_CUDA_CHECK(cudaSetDevice(device));
  double* dX = nullptr;
  OzWorkspace w;
  cudaStream_t s;
  GPC_CUDA_CHECK(cudaStreamCreate(&s));
  GPC_CUDA_CHECK(cudaMalloc(&dX, (size_t)R * K * sizeof(double)));
  GPC_CUDA_CHECK(cudaMemcpy(dX, X, (size_t)R * K * sizeof(double), cudaMemcpyHostToDevice));
  int rc = slice_operand(w, 0, dX, kc ? K : R, kc != 0, R, K, S, s, nullptr);
  if (rc == GPC_OK) {
    GPC_CUDA_CHECK(cudaStreamSynchronize(s));
    GPC_CUDA_CHECK(cudaMemcpy(slices_out, w.sl[0], (size_t)S * R * K, cudaMemcpyDeviceToHost));
    GPC_CUDA_CHECK(cudaMemcpy(scale_out, w.scale[0], (size_t)R * sizeof(double), cudaMemcpyDeviceToHost));
  }
  cudaFree(dX);
  cudaFree(w.sl[0]);
  cudaFree(w.emax[0]);
  cudaFree(w.scale[0]);
  cudaFree(w.errflag);
  if (w.done) cudaEventDestroy(w.done);
  cudaStreamDestroy(s);
  return rc;
}
