// Strip-sweep implicit-GEMM 3x3 dilated convolution on tcgen05 (sm_100a) for the dense whole-volume path
// (layers conv2..conv5 of cnn_cort/nets.py:172-177 evaluated at every pixel of a slice).
//
// Map layout ("wide rows"): all slices of a view lie side by side, position (row r, slice s, col c) is pixel
// r * Pw + s * C1 + c with Pw = ns * C1.  A CTA owns a strip of 128 adjacent wide columns and sweeps down the rows:
// every input row of the strip is TMA-loaded ONCE into a shared-memory ring and serves the three filter rows of
// three consecutive output rows (the flattened one-tile-per-launch kernel loaded every row three times).  With
// dilation d the rows r, r+d, r+2d interact, so a sweep runs over one residue class of rows (r = q + d*n).
//
//   warp 0    TMA producer: one (128 + 2d)-pixel box per input row (two for the 256 B/pixel layout)
//   warp 1    MMA issue, converged, compile-time unrolled: per output row 9 taps x KSTEPS x
//             { xh * [wh | wl] (N = 2*bn),  xl * wh (N = bn) }  -- the bf16x3 split product as two MMAs;
//             column taps are descriptor start offsets inside the row box, weights stay resident (k-step-packed panels)
//   warps 2-9 epilogue: TMEM -> BN scale/shift -> PReLU -> [fused 2x2 stride-1 max-pool: horizontal neighbour by
//             shuffle / a small exchange buffer, vertical neighbour = the previous row kept in registers]
//             -> split bf16 hi|lo -> warp-transposed coalesced stores
// Two TMEM accumulators: the epilogue of row i overlaps the MMAs of row i+1.
//
// Pixel formats: F32CH: 128 B = 32 bf16 hi | 32 bf16 lo (20-channel maps); F64CH: 256 B = 64 hi | 64 lo.
#include "tc_common.cuh"

namespace sc {

constexpr int SW_SLOT_HALF = 17408;   // bytes reserved per TMA box: (128 + 2*4) pixels x 128 B, 1024-aligned

struct SweepArgs {
  int Pw, R;              // wide-row pitch (pixels) and number of rows of the map geometry
  int bn;                 // accumulator columns per half (output channels rounded up to 16)
  int nstrips, strip_w;   // strips start every strip_w wide columns (128, or 128 - pool reach)
  int nseg, L;            // row segments per (strip, class); class rows per segment
  int n_items;
  int stages;
  int in_boxes;           // 1: F32CH input (hi|lo in one 128 B row), 2: F64CH input (hi box, lo box)
  int lo_off;             // byte offset of the lo operand inside a ring slot (64 or SW_SLOT_HALF)
  int out_fmt;            // 1: F32CH, 0: F64CH
  int out_chunks;         // 16-channel chunks written per pixel (chunks beyond bn are zeros)
  int npanels;            // weight panels (4 k-steps each)
  unsigned char* out;
  const float* scale; const float* shift; const float* alpha;
};

__device__ __forceinline__ void sweep_mma(uint32_t acc, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate, uint32_t elected) {
  // descriptors: low word = (addr >> 4) | version/LBO bits, high word constant (SBO 1024 B, SWIZZLE_128B)
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "mov.b64 da, {%1, %6};\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}" ::"r"(acc), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(elected), "r"(0x40004040u) : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// KSTEPS: 16-channel k-steps per tap; DIL: dilation; POOL: fuse the stride-1 max-pool with window {0, DIL}^2;
// NCH: 16-column chunks per epilogue warp; LO64: lo operand 64 B into the row (F32CH) instead of in its own box
template <int KSTEPS, int DIL, int POOL, int NCH, bool LO64>
__global__ void __launch_bounds__(320, 1)
conv_sweep_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW, const SweepArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int BOXPX = 128 + 2 * DIL;
  constexpr int BOX_BYTES = BOXPX * 128;
  constexpr int NBOX = LO64 ? 1 : 2;
  constexpr int SLOT = NBOX * SW_SLOT_HALF;
  constexpr int LO_OFF = LO64 ? 64 : SW_SLOT_HALF;
  const int panel_bytes = 2 * a.bn * 128;
  const int w_bytes = a.npanels * panel_bytes;
  uint8_t* sW = smem;
  uint8_t* sRing = smem + w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + a.stages * SLOT);
  uint64_t* full = bars;
  uint64_t* empty = bars + a.stages;
  uint64_t* tfull = bars + 2 * a.stages;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
  float* s_const = reinterpret_cast<float*>(tmem_slot + 2);            // [scale 64 | shift 64 | alpha 64]
  float* s_xch = s_const + 192;                                        // [2 row parities][2 groups][NCH][4 quarters][2][16]
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_xch + 2 * 2 * 2 * 4 * 2 * 16);   // [8 warps][32 rows][80 B]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 8); }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    const bool ok = i < a.bn;
    s_const[i] = ok ? __ldg(a.scale + i) : 0.f;
    s_const[64 + i] = ok ? __ldg(a.shift + i) : 0.f;
    s_const[128 + i] = ok ? __ldg(a.alpha + i) : 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // item -> (strip, class q, segment); identical in every role
  auto decode = [&](int item, int& w0, int& q, int& n0, int& Lc) {
    const int strip = item % a.nstrips;
    const int rest = item / a.nstrips;
    q = rest % DIL;
    const int seg = rest / DIL;
    w0 = strip * a.strip_w;
    const int Nq = (a.R - q + DIL - 1) / DIL;      // class rows
    n0 = seg * a.L;
    Lc = Nq - n0 < a.L ? Nq - n0 : a.L;            // output class rows of this item (may be <= 0: nothing to do)
  };

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapW)) : "memory");
      mbar_expect_tx(wfull, (uint32_t)w_bytes);
      for (int p = 0; p < a.npanels; ++p) tma_load_2d(&mapW, wfull, sW + p * panel_bytes, 0, p * 2 * a.bn);
      uint32_t g = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        int w0, q, n0, Lc;
        decode(item, w0, q, n0, Lc);
        if (Lc <= 0) continue;
        const int nload = Lc + POOL + 2;
        for (int m = 0; m < nload; ++m, ++g) {
          const uint32_t s = g % a.stages, use = g / a.stages;
          if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
          mbar_expect_tx(&full[s], (uint32_t)(NBOX * BOX_BYTES));
          uint8_t* sp = sRing + s * SLOT;
          const int r = q + DIL * (n0 + m);          // rows beyond the map are zero-filled by the TMA unit
          tma_load_3d(&mapA, &full[s], sp, 0, w0, r);
          if (NBOX == 2) tma_load_3d(&mapA, &full[s], sp + SW_SLOT_HALF, 64, w0, r);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.bn >> 2) << 17) | ((uint32_t)(128 >> 4) << 24);  // N = 2*bn
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // N = bn
    mbar_wait(wfull, 0);
    const uint32_t w_lo = desc_lo(smem_u32(sW));
    const uint32_t ring_lo = desc_lo(smem_u32(sRing));
    const uint32_t panel16 = (uint32_t)(panel_bytes >> 4);
    uint32_t g0 = 0, t = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      int w0, q, n0, Lc;
      decode(item, w0, q, n0, Lc);
      if (Lc <= 0) continue;
      const int nrow = Lc + POOL;
      for (int m = 0; m < nrow; ++m, ++t) {
        const uint32_t b = t & 1;
        mbar_wait(&tempty[b], ((t >> 1) & 1) ^ 1);
        uint32_t sl[3];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const uint32_t g = g0 + m + ky;
          const uint32_t s = g % a.stages;
          if (m == 0 || ky == 2) mbar_wait(&full[s], (g / a.stages) & 1);
          sl[ky] = ring_lo + s * (SLOT >> 4);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + b * 256;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
              const int gk = (ky * 3 + kx) * KSTEPS + ks;                    // compile-time after unrolling
              const uint32_t a_hi = sl[ky] + (uint32_t)(kx * DIL * 8 + ks * 2);
              const uint32_t a_lo = a_hi + (uint32_t)(LO_OFF >> 4);
              const uint32_t bd = w_lo + (uint32_t)(gk >> 2) * panel16 + (uint32_t)((gk & 3) * 2);
              sweep_mma(acc, a_hi, bd, idesc1, gk != 0, leader);
              sweep_mma(acc, a_lo, bd, idesc2, 1, leader);
            }
          }
        }
        if (leader) {
          umma_commit(&empty[(g0 + m) % a.stages]);
          if (m == nrow - 1) {
            umma_commit(&empty[(g0 + m + 1) % a.stages]);
            umma_commit(&empty[(g0 + m + 2) % a.stages]);
          }
          umma_commit(&tfull[b]);
        }
        __syncwarp();
      }
      g0 += nrow + 2;
    }
  } else {
    const int q4 = warp & 3;                 // TMEM lane quarter this warp may read
    const int ew = warp - 2;                 // 0..7
    const int grp = ew >> 2;                 // column-chunk group
    uint8_t* stg = s_stage + ew * 2560;
    uint8_t* mine = stg + lane * 80;
    const int px_bytes = a.out_fmt ? 128 : 256;
    const int lo_byte = a.out_fmt ? 64 : 128;
    uint32_t t = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      int w0, q, n0, Lc;
      decode(item, w0, q, n0, Lc);
      if (Lc <= 0) continue;
      const int nrow = Lc + POOL;
      float hprev[NCH][16];
#pragma unroll
      for (int j = 0; j < NCH; ++j)
#pragma unroll
        for (int k = 0; k < 16; ++k) hprev[j][k] = 0.f;
      for (int m = 0; m < nrow; ++m, ++t) {
        const uint32_t b = t & 1;
        mbar_wait(&tfull[b], (t >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float v[NCH][16];
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const int c0 = (grp + 2 * j) * 16;
          if (c0 < a.bn) {
            uint32_t r1[16], r2[16];
            const uint32_t taddr = tmem_base + b * 256 + ((uint32_t)(q4 * 32) << 16) + (uint32_t)c0;
            tmem_ld16(taddr, r1);
            tmem_ld16(taddr + (uint32_t)a.bn, r2);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float acc = __uint_as_float(r1[k]) + __uint_as_float(r2[k]);
              v[j][k] = prelu(fmaf(acc, s_const[c0 + k], s_const[64 + c0 + k]), s_const[128 + c0 + k]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[j][k] = 0.f;
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[b]);

        int r_out = q + DIL * (n0 + m);
        bool emit = true;
        if (POOL) {
          // horizontal neighbour (DIL pixels to the right): same warp by shuffle, next quarter through shared memory
          float* xw = s_xch + ((((t & 1) * 2 + grp) * 2) * 4 + q4) * 2 * 16;      // [row parity][grp][j][q4][pd][16], j = 0
          if (lane < DIL) {
#pragma unroll
            for (int j = 0; j < NCH; ++j)
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                *reinterpret_cast<float4*>(xw + j * 4 * 2 * 16 + lane * 16 + 4 * k4) = make_float4(v[j][4 * k4], v[j][4 * k4 + 1], v[j][4 * k4 + 2], v[j][4 * k4 + 3]);
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          const float* xr = s_xch + ((((t & 1) * 2 + grp) * 2) * 4 + ((q4 + 1) & 3)) * 2 * 16 + (lane >= 32 - DIL ? (lane - (32 - DIL)) * 16 : 0);
#pragma unroll
          for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              float nb = __shfl_down_sync(0xffffffffu, v[j][k], DIL);
              if (lane >= 32 - DIL) nb = xr[j * 4 * 2 * 16 + k];
              const float h = fmaxf(v[j][k], nb);
              v[j][k] = fmaxf(h, hprev[j][k]);     // vertical neighbour: previous class row
              hprev[j][k] = h;
            }
          emit = m > 0;
          r_out -= DIL;
        }
        if (!emit || r_out >= a.R) continue;       // warp-uniform
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const int c0 = (grp + 2 * j) * 16;
          if (c0 >= a.out_chunks * 16) continue;   // warp-uniform
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v[j][2 * k]), h1 = __float2bfloat16_rn(v[j][2 * k + 1]);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(v[j][2 * k] - __bfloat162float(h0));
            const __nv_bfloat16 l1 = __float2bfloat16_rn(v[j][2 * k + 1] - __bfloat162float(h1));
            hi[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lo[k] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          }
          *reinterpret_cast<uint4*>(mine) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(mine + 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          *reinterpret_cast<uint4*>(mine + 32) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<uint4*>(mine + 48) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          __syncwarp();
          unsigned char* orow = a.out + (long long)r_out * a.Pw * px_bytes + c0 * 2;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int half = i >> 1, row = (i & 1) * 16 + (lane >> 1), part = lane & 1;
            const uint4 d = *reinterpret_cast<const uint4*>(stg + row * 80 + half * 32 + part * 16);
            const int pl = q4 * 32 + row;          // pixel inside the strip
            const int w = w0 + pl;
            if (pl < a.strip_w && w < a.Pw)
              *reinterpret_cast<uint4*>(orow + (long long)w * px_bytes + half * lo_byte + part * 16) = d;
          }
          __syncwarp();
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
template <int KSTEPS, int DIL, int POOL, int NCH, bool LO64>
static int launch_sweep_t(sc_ctx* ctx, const CUtensorMap& mapA, const CUtensorMap& mapW, const SweepArgs& a, size_t smem, cudaStream_t st) {
  auto kern = conv_sweep_kernel<KSTEPS, DIL, POOL, NCH, LO64>;
  static bool configured = false;
  if (!configured) {
    SC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int grid = a.n_items < ctx->sm_count ? a.n_items : ctx->sm_count;
  kern<<<grid, 320, smem, st>>>(mapA, mapW, a);
  ctx->launches++;
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int launch_conv_sweep(sc_ctx* ctx, const SweepW& w, int layer, const float* in, int in_fmt, float* out, int out_fmt, int out_chunks,
                      int Pw, int R, int dil, int pool, int prof_cls, cudaStream_t st) {
  TcState* s = reinterpret_cast<TcState*>(ctx->tc_state);
  SC_CHECK(s != nullptr, SC_ERR_UNSUPPORTED, "tcgen05 back-end not initialised");
  if (Pw <= 0 || R <= 0) return SC_OK;
  SweepArgs a;
  a.Pw = Pw; a.R = R; a.bn = w.bn;
  a.strip_w = 128 - (pool ? dil : 0);
  a.nstrips = (Pw + a.strip_w - 1) / a.strip_w;
  const int nq = (R + dil - 1) / dil;                                   // class rows of the largest class
  // segments: enough items to balance the SMs, rows per item long enough to amortise the 2 (+1) halo rows
  int nseg = (4 * ctx->sm_count + a.nstrips * dil - 1) / (a.nstrips * dil);
  int L = (nq + nseg - 1) / nseg;
  if (L < 12) L = 12;
  if (L > 96) L = 96;
  nseg = (nq + L - 1) / L;
  a.L = L; a.nseg = nseg;
  a.n_items = a.nstrips * dil * nseg;
  a.in_boxes = in_fmt ? 1 : 2;
  a.lo_off = in_fmt ? 64 : SW_SLOT_HALF;
  a.out_fmt = out_fmt; a.out_chunks = out_chunks;
  a.npanels = w.npanels;
  a.out = reinterpret_cast<unsigned char*>(out);
  a.scale = w.scale; a.shift = w.shift; a.alpha = w.alpha;
  SC_CHECK(out_chunks * 16 <= (out_fmt ? 32 : 64) && w.bn <= 64 && w.bn % 16 == 0, SC_ERR_ARG, "conv_sweep: bad channel geometry");
  const int slot = a.in_boxes * SW_SLOT_HALF;
  const int w_bytes = w.npanels * 2 * w.bn * 128;
  const int fixed = 1024 + w_bytes + 256 /*barriers, tmem slot*/ + 192 * 4 + 2 * 2 * 2 * 4 * 2 * 16 * 4 + 8 * 2560;
  a.stages = (227 * 1024 - fixed) / slot;
  if (a.stages > 8) a.stages = 8;
  SC_CHECK(a.stages >= 3, SC_ERR_ARG, "conv_sweep: ring does not fit (layer %d)", layer);
  const size_t smem = (size_t)fixed + (size_t)a.stages * slot;

  CUtensorMap mapA, mapW;
  {
    const int pxe = in_fmt ? 64 : 128;                                    // bf16 elements per pixel
    cuuint64_t dims[3] = {(cuuint64_t)pxe, (cuuint64_t)Pw, (cuuint64_t)R};
    cuuint64_t strides[2] = {(cuuint64_t)pxe * 2, (cuuint64_t)Pw * pxe * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)(128 + 2 * dil), 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = s->encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<float*>(in), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SC_CHECK(r == CUDA_SUCCESS, SC_ERR_CUDA, "conv_sweep: cuTensorMapEncodeTiled(A) failed with %d (Pw %d R %d)", (int)r, Pw, R);
  }
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)w.npanels * 2 * w.bn};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)(2 * w.bn)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = s->encode(&mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w.panels, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SC_CHECK(r == CUDA_SUCCESS, SC_ERR_CUDA, "conv_sweep: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
  }
  ProfScope prof(ctx, prof_cls, st);
  const int nch = (out_chunks + 1) / 2;
  if (w.ksteps == 2 && dil == 1 && pool && nch == 1 && in_fmt) return launch_sweep_t<2, 1, 1, 1, true>(ctx, mapA, mapW, a, smem, st);   // conv2 + pool1
  if (w.ksteps == 2 && dil == 2 && !pool && nch == 2 && in_fmt) return launch_sweep_t<2, 2, 0, 2, true>(ctx, mapA, mapW, a, smem, st);  // conv3
  if (w.ksteps == 3 && dil == 2 && pool && nch == 2 && !in_fmt) return launch_sweep_t<3, 2, 1, 2, false>(ctx, mapA, mapW, a, smem, st); // conv4 + pool2
  if (w.ksteps == 3 && dil == 4 && !pool && nch == 2 && !in_fmt) return launch_sweep_t<3, 4, 0, 2, false>(ctx, mapA, mapW, a, smem, st); // conv5
  // patchwise maps (dilation 1 everywhere, stride-2 pools stay separate passes)
  if (w.ksteps == 2 && dil == 1 && !pool && nch == 1 && in_fmt) return launch_sweep_t<2, 1, 0, 1, true>(ctx, mapA, mapW, a, smem, st);
  if (w.ksteps == 2 && dil == 1 && !pool && nch == 2 && in_fmt) return launch_sweep_t<2, 1, 0, 2, true>(ctx, mapA, mapW, a, smem, st);
  if (w.ksteps == 3 && dil == 1 && !pool && nch == 2 && !in_fmt) return launch_sweep_t<3, 1, 0, 2, false>(ctx, mapA, mapW, a, smem, st);
  set_error("conv_sweep: no kernel instance for ksteps=%d dil=%d pool=%d chunks=%d in_fmt=%d", w.ksteps, dil, pool, out_chunks, in_fmt);
  return SC_ERR_ARG;
}

}  // namespace sc
