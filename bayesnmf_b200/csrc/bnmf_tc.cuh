// Tensor-core reconstruction Mhat = P diag(A) E (get_Mhat_, R/utils.R:29-49) for the Normal likelihood, on the
// 5th-generation tensor cores of sm_100a: hand-written tcgen05.mma (UMMA) with the accumulators in tensor memory.
//
// The reference's `%*%` is an fp64 dgemm and every conditional of the Normal model reads the residual M - Mhat,
// so the product must hold ~1e-9 of the row/column scale -- far below what one bf16 / tf32 pass gives.  The
// kernel therefore runs an error-free "integer slice" scheme (Ozaki-style) on the bf16 path:
//   * every row of P (with A folded in) and every column of E is scaled by a power of two into [0, 1) and rounded
//     to a 40-bit fixed-point number, cut into five 8-bit digits; a digit (0..255) is exact in bf16;
//   * digit plane i of P times digit plane j of E is ONE tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, M = 128,
//     N = 32, K = 16 per signature chunk).  Every product (< 2^16) and every partial sum of the <= 5 x 48 products
//     that share an accumulator (< 2^24) is an integer that fp32 represents exactly: the tensor core adds nothing
//     but exact integers, whatever its internal rounding mode (N <= 48; more signatures: the fp64 kernel);
//   * the digit-plane products are accumulated by weight s = i + j into six accumulators of 32 columns in tensor
//     memory (s = 0..5, 19 products; the six pairs of weight 6 .. 8 are below 2^-46 of the scale and are not
//     formed), 192 of the 256 columns a CTA allocates -- two CTAs share an SM -- read back with tcgen05.ld and
//     recombined in fp64:   Mhat[k,g] = 2^(eP[k] + eE[g] - 80) * sum_s T_s[k,g] 2^(64 - 8 s).
// The only error is the fixed-point rounding of the inputs: |dMhat| <= N 2^-40 2^(eP[k]+eE[g]), i.e. ~1e-11 relative
// to (largest entry of the row of P) x (largest entry of the column of E) -- an ABSOLUTE bound, which is what
// residuals need (four digits, 2^-32, were measured first: conditional draws of small exposures then left the 1e-6
// parity band); the Poisson paths divide by Mhat and take its logarithm cell by cell and keep the fp64 kernel.
//
// k_tc_prep_P cuts P once per iteration (digit planes already in the shared-memory tile layout, row exponents).
// k_mhat_tc: one CTA (128 threads) per tile of 128 mutation types x 32 genomes copies its planes of P and its
// block of E into shared memory with coalesced 16-byte loads, cuts the columns of E, and has the tiles in the
// canonical K-major no-swizzle UMMA layout (core matrices of 8 rows x 16 bytes); one elected thread issues the
// 19 x ceil(N/16) MMAs, tcgen05.commit signals an mbarrier, and each warp drains its 32 lanes of tensor memory.
// Operand bytes are tiny (N <= 64), so there is no TMA pipeline to overlap: the kernel is bound by the K x G
// doubles it writes.
#pragma once
#include <stdint.h>

namespace bnmf {

constexpr int TC_M = 128, TC_N = 32, TC_COLS = 256, TC_PLANES = 5, TC_GROUPS = 6, TC_BITS = 8 * TC_PLANES;
constexpr int TC_MAX_N = 48;      // 5 pairs x 48 x 255^2 < 2^24: the partial sums of an accumulator stay exact integers in fp32

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading byte
// offset (between the two 8-element K chunks of an MMA), stride byte offset (between 8-row groups), all >> 4;
// version 1 (Blackwell) in bits [46,48), layout type 0 in bits [61,64)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15, 16 = 0), N >> 3 in bits [17,23), M >> 4 in bits [24,29)
__device__ __forceinline__ uint32_t tc_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}

// element (row r, K index c) of a tile of R rows x 16 bf16 in the canonical layout: chunk (c / 8) at LBO = R * 16
// bytes, 8-row group (r / 8) at SBO = 128 bytes, row (r % 8) at 16 bytes, element (c % 8) at 2 bytes
__host__ __device__ __forceinline__ uint32_t tc_off(int R, int r, int c) {
  return (uint32_t)((c >> 3) * (R * 16) + (r >> 3) * 128 + (r & 7) * 16 + (c & 7) * 2);
}
__device__ __forceinline__ unsigned short tc_bf16_of_digit(uint32_t d) { return (unsigned short)(__float_as_uint((float)d) >> 16); }

// scale 2^e with v < 2^e for the largest v of a row / column (0 for an all-zero one) and the five digits of
// rint(v 2^(40 - e)) (clamped below 2^40)
__device__ __forceinline__ int tc_exponent(double mx) { int e = 0; if (mx > 0.0) frexp(mx, &e); return e; }
__device__ __forceinline__ unsigned long long tc_fixed(double v, double up /* 2^(40 - e) */) {
  const double x = rint(v * up);
  return x >= 1099511627775.0 ? 1099511627775ull : (x > 0.0 ? (unsigned long long)x : 0ull);
}
__device__ __forceinline__ uint32_t tc_digit(unsigned long long x, int i) { return (uint32_t)(x >> (8 * (TC_PLANES - 1 - i))) & 255u; }
__device__ __forceinline__ double tc_pow2(int e) { return __longlong_as_double((long long)(e + 1023) << 52); }   // |e| < 1023

// digit planes of P with A folded in, per tile of 128 mutation types, already in the shared-memory layout of the
// A operand: Pd[mtile][plane i][chunk kc][4096 bytes]; eP[k] = the row's exponent
__host__ __device__ inline size_t tc_planes_bytes(int K, int N) { return (size_t)((K + TC_M - 1) / TC_M) * TC_PLANES * ((N + 15) / 16) * (TC_M * 32); }
template <typename T>
__global__ void __launch_bounds__(128) k_tc_prep_P(const T* __restrict__ P, const int32_t* __restrict__ A, unsigned char* __restrict__ Pd,
                                                   int* __restrict__ eP, int K, int N) {
  const int KC = (N + 15) / 16;
  const int k = blockIdx.x * TC_M + threadIdx.x;
  unsigned char* base = Pd + (size_t)blockIdx.x * TC_PLANES * KC * (TC_M * 32);
  double mx = 0.0;
  if (k < K) {
#pragma unroll 8
    for (int n = 0; n < N; ++n) { const double v = A[n] ? (double)P[(long long)k + (long long)K * n] : 0.0; mx = v > mx ? v : mx; }
  }
  const int e = tc_exponent(mx);
  const double up = tc_pow2(TC_BITS - e);
  for (int n = 0; n < 16 * KC; ++n) {
    const double v = (k < K && n < N && A[n]) ? (double)P[(long long)k + (long long)K * n] : 0.0;
    const unsigned long long x = tc_fixed(v, up);
    const uint32_t off = tc_off(TC_M, threadIdx.x, n & 15);
#pragma unroll
    for (int i = 0; i < TC_PLANES; ++i)
      *reinterpret_cast<unsigned short*>(base + ((size_t)i * KC + (n >> 4)) * (TC_M * 32) + off) = tc_bf16_of_digit(tc_digit(x, i));
  }
  if (k < K) eP[k] = e;
}

// shared memory: [5 digit planes][KC chunks] tiles of A (128 x 16 bf16 = 4 KB) and of B (32 x 16 = 1 KB), the block
// of E (32 x N elements of T), the column scales, the mbarrier and the tensor-memory base address
template <typename T> __host__ __device__ inline size_t tc_smem_bytes(int N) {
  const int KC = (N + 15) / 16;
  return (size_t)TC_PLANES * KC * (TC_M * 32 + TC_N * 32) + (size_t)TC_N * N * sizeof(T) + TC_N * sizeof(double) + 64;
}

template <typename T>
__global__ void __launch_bounds__(128, 2)
k_mhat_tc(const unsigned char* __restrict__ Pd, const int* __restrict__ ePg, const T* __restrict__ E, T* __restrict__ Mhat, int K, int N, long long G) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  const int KC = (N + 15) / 16;
  unsigned char* sA = tc_smem;                                    // [plane][chunk][4096]
  unsigned char* sB = sA + (size_t)TC_PLANES * KC * (TC_M * 32);  // [plane][chunk][1024]
  T* sE = reinterpret_cast<T*>(sB + (size_t)TC_PLANES * KC * (TC_N * 32));   // [32 genomes][N]
  double* scE = reinterpret_cast<double*>(sE + (size_t)TC_N * N);
  uint64_t* bar = reinterpret_cast<uint64_t*>(scE + TC_N);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const long long g0 = (long long)blockIdx.x * TC_N;
  const int k0 = blockIdx.y * TC_M;

  if (warp == 0) {     // tensor memory: 256 columns (6 accumulators x 32), two CTAs per SM
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(tmem_slot)), "r"(TC_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tc_smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- operands: the planes of P (coalesced 16-byte copies), the block of E, zeroed B tiles ----
  {
    const uint4* src = reinterpret_cast<const uint4*>(Pd + (size_t)blockIdx.y * TC_PLANES * KC * (TC_M * 32));
    uint4* dst = reinterpret_cast<uint4*>(sA);
    const int nA = TC_PLANES * KC * (TC_M * 32) / 16;
    for (int i = tid; i < nA; i += 128) dst[i] = src[i];
    uint4* zb = reinterpret_cast<uint4*>(sB);
    const int nB = TC_PLANES * KC * (TC_N * 32) / 16;
    for (int i = tid; i < nB; i += 128) zb[i] = make_uint4(0u, 0u, 0u, 0u);
    const long long ng = G - g0 < TC_N ? G - g0 : TC_N;            // genomes of this tile: one contiguous block of E
    const T* eblk = E + (long long)N * g0;
    const int ne = (int)(ng * N);
    for (int i = tid; i < ne; i += 128) sE[i] = eblk[i];
  }
  __syncthreads();
  {   // four threads per genome: column maximum (two shuffles), then every thread cuts its quarter of the signatures
    const int gl = tid >> 2, q = tid & 3;
    const long long g = g0 + gl;
    const T* col = sE + (size_t)gl * N;
    double mx = 0.0;
    if (g < G) for (int n = q; n < N; n += 4) { const double v = (double)col[n]; mx = v > mx ? v : mx; }
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const int e = tc_exponent(mx);
    if (g < G) {
      const double up = tc_pow2(TC_BITS - e);
      for (int n = q; n < N; n += 4) {
        const unsigned long long y = tc_fixed((double)col[n], up);
        const uint32_t off = tc_off(TC_N, gl, n & 15);
#pragma unroll
        for (int j = 0; j < TC_PLANES; ++j)
          *reinterpret_cast<unsigned short*>(sB + ((size_t)j * KC + (n >> 4)) * (TC_N * 32) + off) = tc_bf16_of_digit(tc_digit(y, j));
      }
    }
    if (q == 0) scE[gl] = tc_pow2(e - TC_BITS);
  }
  // the tiles were written through the generic proxy, the tensor core reads them through the async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  // ---- 19 digit-plane products per signature chunk, accumulated by weight s = i + j <= 5 ----
  if (tid == 0) {
    const uint32_t idesc = tc_idesc(TC_M, TC_N);
    uint32_t started = 0u;
    for (int kc = 0; kc < KC; ++kc)
      for (int i = 0; i < TC_PLANES; ++i)
        for (int j = 0; j < TC_PLANES && i + j < TC_GROUPS; ++j) {
          const int s = i + j;
          const uint64_t da = tc_desc(tc_smem_u32(sA + ((size_t)i * KC + kc) * (TC_M * 32)), TC_M * 16, 128);
          const uint64_t db = tc_desc(tc_smem_u32(sB + ((size_t)j * KC + kc) * (TC_N * 32)), TC_N * 16, 128);
          tc_mma(tmem + (uint32_t)(s * TC_N), da, db, idesc, (started >> s) & 1u);
          started |= 1u << s;
        }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(tc_smem_u32(bar)) : "memory");
  }
  {   // everybody waits for the accumulators (phase 0 of the barrier)
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(tc_smem_u32(bar)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- epilogue: thread = row (lane of tensor memory); 8 genomes at a time, five accumulators each ----
  {
    const int k = k0 + tid;
    const double scP = tc_pow2((k < K ? ePg[k] : 0) - TC_BITS);    // 2^(eP - 40) x 2^(eE - 40) = 2^(eP + eE - 80)
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const double w[TC_GROUPS] = {18446744073709551616.0, 72057594037927936.0, 281474976710656.0, 1099511627776.0, 4294967296.0, 16777216.0};   // 2^(64 - 8 s)
    for (int c0 = 0; c0 < TC_N; c0 += 8) {
      uint32_t v[TC_GROUPS][8];
#pragma unroll
      for (int s = 0; s < TC_GROUPS; ++s) tc_ld8(tmem + lane_base + (uint32_t)(s * TC_N + c0), v[s]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (k < K) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          double acc = 0.0;
#pragma unroll
          for (int s = TC_GROUPS - 1; s >= 0; --s) acc += (double)__uint_as_float(v[s][c]) * w[s];
          const long long g = g0 + c0 + c;
          if (g < G) Mhat[(long long)k + (long long)K * g] = (T)(acc * scP * scE[c0 + c]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TC_COLS) : "memory");
}

}  // namespace bnmf
