// tcgen05 / TMEM / TMA back-end of the dense-layer / implicit-GEMM problem (sm_100a).
//
//   C[m, n] = prelu( sum_tap sum_k A[pixel m shifted by tap, k] * W[n, tap*kc + k] + bias[n] )
//
// used for d1 (as a 3x3 dilation-4 conv over the NHWC-64 conv5 map), FC1 and fc_2
// (reference: DenseLayer + PReLU at cnn_cort/nets.py:179-180, 217-218, 227-228).
//
// Precision: single-pass TF32 / fp16 misses the 1e-3 softmax tolerance on unsaturated inputs
// (measured: 1.4e-3), so the product is the three-MMA split  xh*wh + xh*wl + xl*wh  of bf16 pairs
// (x = xh + xl to ~2^-17, kind::f16 with fp32 accumulation in TMEM; oracle emulation: 1.8e-5).
// Activations and weights are stored per 64-wide k block as 64 bf16 hi | 64 bf16 lo, i.e. exactly
// the bytes of the fp32 row, so one 128 B swizzle row holds one block half.
//
// One CTA = one 128-row x BN-column output tile, 192 threads:
//   warp 0   TMA producer: per k block four tiled loads into one shared-memory stage -- A hi and
//            A lo (4-D box 64 bf16 x 128 pixels, tap shift applied to the pixel/line coordinates,
//            zero fill outside) and W hi and W lo (2-D box 64 x BN)
//   warp 1   allocates TMEM; one lane issues 12 x tcgen05.mma (M=128, N=BN, K=16) per stage,
//            tcgen05.commit releases the stage / signals the accumulator
//   warps 2-5 epilogue: tcgen05.ld 32x32b.x16 -> bias + PReLU -> plain fp32 or split bf16 rows
// Two CTAs are co-resident per SM (<= 256 TMEM columns and <= 112 KB shared memory each) so that
// one tile's loads / epilogue overlap the other's MMAs.
#include <cuda.h>

#include "common.cuh"

namespace sc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcState {
  EncodeTiledFn encode;
};

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                    // logical k per block (64 bf16 = one 128 B swizzle row)
constexpr int TC_A_HALF = TC_BM * 128;       // bytes of the hi (or lo) half of an A stage
constexpr int TC_THREADS = 192;
constexpr int TC_TMEM_COLS = 256;

struct TcArgs {
  int kpt;            // k-blocks per tap (kc / 64)
  int nkb;            // total k-blocks
  int stages;
  int bn;             // tile columns (multiple of 16, <= 192)
  int nt, mt;         // tiles along n, along m (per line)
  int Y;              // lines per plane
  int M;              // rows per line
  int n_store, npad, c_col0;
  int a_y0, a_z0;     // coordinate offsets of line / plane in the A tensor map
  int tap_dx[9], tap_dy[9];
  float* C;
  long long ldc, c_ys, c_zs;
  const float* bias;
  const float* alpha;
  const float* scale;
  int out_split;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row atoms 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // carve: [stages x (A hi | A lo | W hi | W lo)][barriers][tmem ptr][bias][alpha]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_half = a.bn * 128;
  const int stage_bytes = 2 * TC_A_HALF + 2 * b_half;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + a.stages;
  uint64_t* accum = bars + 2 * a.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);
  float* s_alpha = s_bias + 192;
  float* s_scale = s_alpha + 192;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long bid = blockIdx.x;
  const int n_tile = (int)(bid % a.nt); bid /= a.nt;
  const int m_tile = (int)(bid % a.mt); bid /= a.mt;
  const int y = (int)(bid % a.Y);
  const int z = (int)(bid / a.Y);
  const int m0 = m_tile * TC_BM, n0 = n_tile * a.bn;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < a.bn; i += TC_THREADS) {
    s_bias[i] = n0 + i < a.npad ? __ldg(a.bias + n0 + i) : 0.f;
    s_alpha[i] = n0 + i < a.npad ? __ldg(a.alpha + n0 + i) : 1.f;
    s_scale[i] = (a.scale && n0 + i < a.npad) ? __ldg(a.scale + n0 + i) : 1.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
      for (int kb = 0; kb < a.nkb; ++kb) {
        const int s = kb % a.stages, it = kb / a.stages;
        if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);
        mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
        uint8_t* st = smem + s * stage_bytes;
        const int tap = kb / a.kpt, ka = (kb - tap * a.kpt) * 2 * TC_BK;   // bf16 element offset of the block's hi half
        const int px = m0 + a.tap_dx[tap], ln = y + a.a_y0 + a.tap_dy[tap], pl = z + a.a_z0;
        tma_load_4d(&mapA, &full[s], st, ka, px, ln, pl);
        tma_load_4d(&mapA, &full[s], st + TC_A_HALF, ka + TC_BK, px, ln, pl);
        tma_load_2d(&mapB, &full[s], st + 2 * TC_A_HALF, kb * 2 * TC_BK, n0);
        tma_load_2d(&mapB, &full[s], st + 2 * TC_A_HALF + b_half, kb * 2 * TC_BK + TC_BK, n0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=F32, A=B=BF16, both K-major, N = bn, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < a.nkb; ++kb) {
        const int s = kb % a.stages, it = kb / a.stages;
        mbar_wait(&full[s], it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem_u32(smem + s * stage_bytes);
        const uint64_t ah = umma_desc(st), al = umma_desc(st + TC_A_HALF);
        const uint64_t wh = umma_desc(st + 2 * TC_A_HALF), wl = umma_desc(st + 2 * TC_A_HALF + b_half);
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 4 x K=16 bf16 (32 B) inside the 128 B swizzle row; small terms first
          const uint64_t o = (uint64_t)(j * 2);
          umma_bf16(tmem_base, al + o, wh + o, idesc, (kb | j) != 0);
          umma_bf16(tmem_base, ah + o, wl + o, idesc, 1);
          umma_bf16(tmem_base, ah + o, wh + o, idesc, 1);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(accum);
    }
    __syncwarp();
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int q = warp & 3;
    mbar_wait(accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = m0 + q * 32 + lane;
    float* crow = a.C + (long long)z * a.c_zs + (long long)y * a.c_ys + (long long)m * a.ldc;
    const bool row_ok = m < a.M;
    for (int c0 = 0; c0 < a.bn; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int n = n0 + c0 + g * 4;
          if (n < a.n_store) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int c = c0 + g * 4 + k;
              v[k] = prelu(fmaf(__uint_as_float(r[g * 4 + k]), s_scale[c], s_bias[c]), s_alpha[c]);
            }
            store_row4(crow, a.c_col0 + n, a.out_split, v[0], v[1], v[2], v[3]);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// plain fp32 [rows][576] -> split bf16 hi|lo blocks (test entry sc_dense_layer only)
__global__ void split_rows_kernel(const float* __restrict__ in, int64_t rows, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (kFeatLd / 4)) return;
  const int64_t m = i / (kFeatLd / 4);
  const int n = (int)(i - m * (kFeatLd / 4)) * 4;
  const float4 v = __ldg(reinterpret_cast<const float4*>(in + m * kFeatLd + n));
  store_split4(out + m * kFeatLd, n, v.x, v.y, v.z, v.w);
}
int launch_split_rows(sc_ctx* ctx, const float* in, int64_t rows, float* out, cudaStream_t st) {
  const int64_t work = rows * (kFeatLd / 4);
  split_rows_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(in, rows, out);
  ctx->launches++;
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int tc_init(sc_ctx* ctx) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    cudaGetLastError();
    set_error("tcgen05 back-end: cuTensorMapEncodeTiled unavailable");
    return SC_ERR_UNSUPPORTED;
  }
  TcState* s = new TcState();
  s->encode = reinterpret_cast<EncodeTiledFn>(fn);
  ctx->tc_state = s;
  return SC_OK;
}

void tc_destroy(sc_ctx* ctx) {
  delete reinterpret_cast<TcState*>(ctx->tc_state);
  ctx->tc_state = nullptr;
}

static int pick_bn(int n_store) {
  // widest tile <= 192 (two CTAs x 256 TMEM columns per SM) that wastes the fewest columns
  int best = 16, best_cost = 1 << 30;
  for (int bn = 192; bn >= 32; bn -= 16) {
    const int tiles = (n_store + bn - 1) / bn;
    const int cost = tiles * bn * 4 + tiles * 128;  // padded columns + A re-reads
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

int launch_gemm_tc(sc_ctx* ctx, const GemmProblem& p, const GemmW& w, cudaStream_t st) {
  TcState* s = reinterpret_cast<TcState*>(ctx->tc_state);
  SC_CHECK(s != nullptr, SC_ERR_UNSUPPORTED, "tcgen05 back-end not initialised");
  if (p.M <= 0 || p.Y <= 0 || p.Z <= 0) return SC_OK;
  SC_CHECK(p.kc % TC_BK == 0 && p.ntaps * p.kc == w.Kpad && p.n_store % 4 == 0 && p.n_store <= w.Npad, SC_ERR_ARG,
           "gemm_tc: bad geometry kc=%d ntaps=%d Kpad=%d n_store=%d", p.kc, p.ntaps, w.Kpad, p.n_store);
  TcArgs a;
  a.kpt = p.kc / TC_BK;
  a.c_col0 = p.c_col0;
  a.nkb = p.ntaps * a.kpt;
  a.bn = pick_bn(p.n_store);
  a.nt = (p.n_store + a.bn - 1) / a.bn;
  a.mt = (p.M + TC_BM - 1) / TC_BM;
  a.Y = p.Y; a.M = p.M; a.n_store = p.n_store; a.npad = w.Npad;
  a.a_y0 = p.a_y0; a.a_z0 = p.a_z0;
  for (int t = 0; t < 9; ++t) { a.tap_dx[t] = t < p.ntaps ? p.tap_dx[t] : 0; a.tap_dy[t] = t < p.ntaps ? p.tap_dy[t] : 0; }
  a.C = p.C; a.ldc = p.ldc; a.c_ys = p.c_ys; a.c_zs = p.c_zs;
  a.bias = w.bias; a.alpha = w.alpha; a.scale = w.scale; a.out_split = p.out_split;
  SC_CHECK(p.c_col0 % 4 == 0, SC_ERR_ARG, "gemm_tc: c_col0 must be a multiple of 4");
  const int stage_bytes = 2 * TC_A_HALF + 2 * a.bn * 128;
  a.stages = (108 * 1024) / stage_bytes;
  if (a.stages > 4) a.stages = 4;
  SC_CHECK(a.stages >= 1, SC_ERR_ARG, "gemm_tc: tile too wide for one stage");
  const size_t smem = 1024 + (size_t)a.stages * stage_bytes + (2 * a.stages + 1) * 8 + 16 + 3 * 192 * 4;

  CUtensorMap mapA, mapB;
  {
    // a row of kc logical columns is 2*kc bf16 (hi | lo per 64-wide block)
    cuuint64_t dims[4] = {(cuuint64_t)p.a_dims[0] * 2, (cuuint64_t)p.a_dims[1], (cuuint64_t)p.a_dims[2], (cuuint64_t)p.a_dims[3]};
    cuuint64_t strides[3] = {(cuuint64_t)p.a_strides[0] * 4, (cuuint64_t)p.a_strides[1] * 4, (cuuint64_t)p.a_strides[2] * 4};
    cuuint32_t box[4] = {TC_BK, TC_BM, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = s->encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<float*>(p.a_base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SC_CHECK(r == CUDA_SUCCESS, SC_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled(A) failed with %d (dims %llu %llu %llu %llu)", (int)r,
             (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)dims[3]);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)w.Kpad * 2, (cuuint64_t)w.Npad};
    cuuint64_t strides[1] = {(cuuint64_t)w.Kpad * 4};
    cuuint32_t box[2] = {TC_BK, (cuuint32_t)a.bn};
    cuuint32_t es[2] = {1, 1};
    CUresult r = s->encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w.w_nk, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SC_CHECK(r == CUDA_SUCCESS, SC_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  static bool configured = false;
  if (!configured) {
    SC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    configured = true;
  }
  SC_CHECK(smem <= 112 * 1024, SC_ERR_ARG, "gemm_tc: shared memory budget exceeded (%zu)", smem);
  const long long blocks = (long long)a.mt * a.nt * p.Y * p.Z;
  SC_CHECK(blocks < (1ll << 31), SC_ERR_ARG, "gemm_tc: grid too large");
  ProfScope prof(ctx, p.prof_cls, st);
  gemm_tc_kernel<<<(unsigned)blocks, TC_THREADS, smem, st>>>(mapA, mapB, a);
  ctx->launches++;
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

}  // namespace sc
