// Device-side scan preparation: what load_patch_batch does on the host before the first patch is cut
// (cnn_cort/base.py:357-372) -- array-order import of a NIfTI volume, intensity normalisation over the non-zero voxels,
// candidate mask and its bounding box -- so that a scan goes host -> device once and the label volume comes back once.
//
// Normalisation is a floating-point reduction; the contract is numpy's own result, bit for bit, because the patches cut
// from the normalised volume are declared bit-exact.  numpy sums a contiguous 1-D array with its pairwise scheme
// (blocks of <= 128 elements summed with 8 strided accumulators, halves split at multiples of 8).  That scheme is
// reproduced here exactly: the non-zero values are compacted in C order (np.nonzero order), every leaf block is summed
// on the device in numpy's order (8 lanes = the 8 accumulators), and the host adds the ~n/100 leaf sums up the same tree.
//   float32 T1:  mean, variance and (x - mean) / std in float32 (numpy keeps float32 scalars with float32 arrays)
//   float64 T1:  everything in float64, result cast to float32 like the patches are (base.py:383)
//   integer T1:  numpy promotes to float64; the sum of integers is exact in any order
// The quotients sum / n are taken in float64 and rounded to the working type, which is what numpy 1.12's
// true_divide(arr, rcount) does for n > 65535 and is identical to the direct division otherwise.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace sc {

// ---------------------------------------------------------------------------------------------------------------
// F-order (NIfTI: x fastest, channel slowest) -> C-order [X][Y][Z][C] (z fastest, channels innermost), bit-preserving
// ---------------------------------------------------------------------------------------------------------------
// in: [C][Zi][Yi][Xi] (x fastest), the whole volume or the compact copy of a box; out: [X][Y][Z][C] at origin (ox, oy, oz)
template <typename E, int ZT>
__global__ void __launch_bounds__(256) f2c_kernel(const E* __restrict__ in, E* __restrict__ out, int Xi, int Yi, int Zi, int C,
                                                  int Y, int Z, int ox, int oy, int oz, int xlo, int xhi) {
  extern __shared__ __align__(16) unsigned char f2c_smem[];
  E* tile = reinterpret_cast<E*>(f2c_smem);                       // [C][ZT][33]
  const int x0 = (xlo & ~31) + blockIdx.x * 32, z0 = blockIdx.y * ZT, y = blockIdx.z;     // only the x tiles that overlap [xlo, xhi)
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  // read: 32 consecutive x per (c, z) line
  for (int line = wrp; line < C * ZT; line += 8) {
    const int c = line / ZT, zz = line - c * ZT;
    const int x = x0 + lane, z = z0 + zz;
    E v = E(0);
    if (x >= xlo && x < xhi && z < Zi) v = in[(((int64_t)c * Zi + z) * Yi + y) * Xi + x];
    tile[(c * ZT + zz) * 33 + lane] = v;
  }
  __syncthreads();
  // write: for one x the ZT x C block is contiguous in the output
  const int run = ZT * C;
  for (int e = threadIdx.x; e < 32 * run; e += 256) {
    const int xx = e / run, r = e - xx * run;
    const int zz = r / C, c = r - zz * C;
    const int x = x0 + xx, z = z0 + zz;
    if (x >= xlo && x < xhi && z < Zi) out[(((int64_t)(ox + x) * Y + (oy + y)) * Z + (oz + z)) * C + c] = tile[(c * ZT + zz) * 33 + xx];
  }
}

// idims: extents of the input array; odims / origin: the C-ordered output volume and where the input's corner goes in it
template <typename E>
static int launch_f2c(sc_ctx* ctx, const void* src, const int32_t* idims, int C, void* dst, const int32_t* odims, const int32_t* origin,
                      cudaStream_t st, int xlo, int xhi) {
  constexpr int ZT = 8;
  const size_t smem = (size_t)C * ZT * 33 * sizeof(E);
  SC_CHECK(smem <= 48 * 1024, SC_ERR_ARG, "sc_import_volume: too many channels (%d)", C);
  dim3 grid((xhi - (xlo & ~31) + 31) / 32, (idims[2] + ZT - 1) / ZT, idims[1]);
  SC_CHECK(grid.y <= 65535 && grid.z <= 65535, SC_ERR_ARG, "sc_import_volume: volume too large");
  f2c_kernel<E, ZT><<<grid, 256, smem, st>>>(reinterpret_cast<const E*>(src), reinterpret_cast<E*>(dst), idims[0], idims[1], idims[2], C,
                                             odims[1], odims[2], origin[0], origin[1], origin[2], xlo, xhi);
  ctx->launches++;
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

// [xlo, xhi): the x range of the input that is copied (the whole extent by default)
static int import_box(sc_ctx* ctx, const void* src, int elem_bytes, const int32_t* idims, int channels, void* dst, const int32_t* odims,
                      const int32_t* origin, cudaStream_t st, int xlo = 0, int xhi = -1) {
  if (xhi < 0) xhi = idims[0];
  switch (elem_bytes) {
    case 1: return launch_f2c<uint8_t>(ctx, src, idims, channels, dst, odims, origin, st, xlo, xhi);
    case 2: return launch_f2c<uint16_t>(ctx, src, idims, channels, dst, odims, origin, st, xlo, xhi);
    case 4: return launch_f2c<uint32_t>(ctx, src, idims, channels, dst, odims, origin, st, xlo, xhi);
    case 8: return launch_f2c<unsigned long long>(ctx, src, idims, channels, dst, odims, origin, st, xlo, xhi);
  }
  set_error("sc_import_volume: elem_bytes must be 1, 2, 4 or 8");
  return SC_ERR_ARG;
}

int import_volume(sc_ctx* ctx, const void* src, int elem_bytes, const int32_t* dims, int channels, void* dst, cudaStream_t st) {
  const int32_t origin[3] = {0, 0, 0};
  return import_box(ctx, src, elem_bytes, dims, channels, dst, dims, origin, st);
}

// Only the box {x0,x1,y0,y1,z0,z1} of a host volume [X,Y,Z(,C)] goes to the device: strided DMA copies (cudaMemcpy3DAsync), one per
// channel for a Fortran-ordered array (runs of x), one for a C-ordered array (runs of z * C), land in / are reordered into the
// box region of the C-ordered device volume `dst`; everything outside the box is left untouched.
int upload_volume_box(sc_ctx* ctx, const void* src_host, int elem_bytes, const int32_t* dims, int channels, int fortran_order,
                      const int32_t* box, void* staging_dev, void* dst, cudaStream_t st) {
  const int X = dims[0], Y = dims[1], Z = dims[2];
  const int bx = box[1] - box[0], by = box[3] - box[2], bz = box[5] - box[4];
  const size_t eb = (size_t)elem_bytes;
  if (!fortran_order) {
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    const size_t row = (size_t)Z * channels * eb;
    p.srcPtr = make_cudaPitchedPtr(const_cast<void*>(src_host), row, row, (size_t)Y);
    p.dstPtr = make_cudaPitchedPtr(dst, row, row, (size_t)Y);
    p.srcPos = make_cudaPos((size_t)box[4] * channels * eb, (size_t)box[2], (size_t)box[0]);
    p.dstPos = p.srcPos;
    p.extent = make_cudaExtent((size_t)bz * channels * eb, (size_t)by, (size_t)bx);
    p.kind = cudaMemcpyHostToDevice;
    SC_CUDA(cudaMemcpy3DAsync(&p, st));
    return SC_OK;
  }
  SC_CHECK(staging_dev != nullptr, SC_ERR_ARG, "sc_upload_volume_box: a Fortran-ordered source needs the staging buffer");
  // In memory the y range of one (channel, z) plane is ONE contiguous run of by * X elements: the copy moves those runs (full x rows;
  // DMA rows of a few hundred bytes -- the x range alone -- reach a third of the PCIe rate) and the reorder drops the x outside the box.
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof(p));
  const size_t run = (size_t)by * X * eb;
  p.srcPtr = make_cudaPitchedPtr(const_cast<void*>(src_host), (size_t)Y * X * eb, (size_t)Y * X * eb, (size_t)Z);
  p.dstPtr = make_cudaPitchedPtr(staging_dev, run, run, (size_t)bz);
  p.srcPos = make_cudaPos((size_t)box[2] * X * eb, (size_t)box[4], 0);
  p.dstPos = make_cudaPos(0, 0, 0);
  p.extent = make_cudaExtent(run, (size_t)bz, (size_t)channels);
  p.kind = cudaMemcpyHostToDevice;
  SC_CUDA(cudaMemcpy3DAsync(&p, st));
  const int32_t idims[3] = {X, by, bz}, origin[3] = {0, box[2], box[4]};
  return import_box(ctx, staging_dev, elem_bytes, idims, channels, dst, dims, origin, st, box[0], box[1]);
}

// ---------------------------------------------------------------------------------------------------------------
// typed access to a raw volume
// ---------------------------------------------------------------------------------------------------------------
template <typename T> struct Acc { typedef double type; };
template <> struct Acc<float> { typedef float type; };

template <typename T>
__device__ __forceinline__ bool is_nz(T v) { return v != T(0); }          // -0.0 == 0 -> False, NaN != 0 -> True (numpy)

__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

constexpr int kPrepPerThread = 16;
constexpr int kPrepBlock = 256 * kPrepPerThread;

template <typename T>
__global__ void __launch_bounds__(256) prep_count_kernel(const T* __restrict__ vol, int64_t total, int32_t* __restrict__ counts) {
  const int64_t start = (int64_t)blockIdx.x * kPrepBlock + (int64_t)threadIdx.x * kPrepPerThread;
  int c = 0;
#pragma unroll
  for (int k = 0; k < kPrepPerThread; ++k)
    if (start + k < total && is_nz(vol[start + k])) ++c;
  __shared__ int wsum[8];
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += wsum[w];
    counts[blockIdx.x] = s;
  }
}

// exclusive scan of the block counts (single CTA; <= a few thousand blocks)
__global__ void __launch_bounds__(1024) prep_scan_kernel(const int32_t* __restrict__ counts, int nblocks, int64_t* __restrict__ offsets,
                                                         int64_t* __restrict__ total) {
  __shared__ int64_t wsum[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nblocks; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    int64_t v = i < nblocks ? counts[i] : 0, inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      int64_t w = wsum[threadIdx.x], winc = w;
      for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (threadIdx.x >= o) winc += t;
      }
      wsum[threadIdx.x] = winc - w;
    }
    __syncthreads();
    const int64_t excl = carry + wsum[threadIdx.x >> 5] + inc - v;
    if (i < nblocks) offsets[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// ordered emit of the non-zero VALUES (np.nonzero order), converted to the accumulation type
template <typename T>
__global__ void __launch_bounds__(256) prep_emit_kernel(const T* __restrict__ vol, int64_t total, const int64_t* __restrict__ offsets,
                                                        typename Acc<T>::type* __restrict__ nzv) {
  typedef typename Acc<T>::type A;
  const int64_t start = (int64_t)blockIdx.x * kPrepBlock + (int64_t)threadIdx.x * kPrepPerThread;
  uint32_t bits = 0;
#pragma unroll
  for (int k = 0; k < kPrepPerThread; ++k)
    if (start + k < total && is_nz(vol[start + k])) bits |= 1u << k;
  const int c = __popc(bits);
  int inc = c;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if ((threadIdx.x & 31) >= o) inc += t;
  }
  __shared__ int wsum[8];
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
  __syncthreads();
  int wbase = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += wsum[w];
  int64_t o = offsets[blockIdx.x] + wbase + inc - c;
  while (bits) {
    const int k = __ffs(bits) - 1;
    bits &= bits - 1;
    nzv[o++] = (A)vol[start + k];
  }
}

// numpy's leaf sum (n <= 128): 8 strided accumulators, ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail; n < 8: plain
// loop from 0.  8 lanes per leaf (lane j = accumulator j); sq: sum (x - mean)^2 with each operation rounded on its own
// (numpy materialises x - mean and x * x as arrays first: no fused multiply-add).
template <typename A>
__global__ void __launch_bounds__(256) prep_leaf_kernel(const A* __restrict__ nzv, const int64_t* __restrict__ leaf_off, int nleaves,
                                                        int sq, A mean, A* __restrict__ leaf_sum) {
  const int leaf = (int)(((int64_t)blockIdx.x * 256 + threadIdx.x) >> 3);
  const int j = threadIdx.x & 7;
  const bool live = leaf < nleaves;
  const int64_t off = live ? leaf_off[leaf] : 0;
  const int n = live ? (int)(leaf_off[leaf + 1] - off) : 0;
  const A* a = nzv + off;
  auto val = [&](int i) -> A {
    A v = a[i];
    if (sq) { const A d = sub_rn(v, mean); v = mul_rn(d, d); }
    return v;
  };
  const bool big = n >= 8;
  const int nb = n - (n & 7);
  A r = A(0);
  if (big) {
    r = val(j);
    for (int i = 8; i < nb; i += 8) r = add_rn(r, val(i + j));
  }
  // every lane of the warp takes part in the shuffles (the four 8-lane groups of a warp may hold leaves of either kind)
  r = add_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
  r = add_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
  r = add_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
  if (live && j == 0) {
    A res = big ? r : A(0);
    for (int i = big ? nb : 0; i < n; ++i) res = add_rn(res, val(i));
    leaf_sum[leaf] = res;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) prep_normalise_kernel(const T* __restrict__ vol, int64_t total, typename Acc<T>::type mean,
                                                             typename Acc<T>::type sd, float* __restrict__ out) {
  typedef typename Acc<T>::type A;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256)
    out[i] = (float)div_rn(sub_rn((A)vol[i], mean), sd);
}

template <typename T>
__global__ void __launch_bounds__(256) prep_mask_kernel(const T* __restrict__ vol, int64_t total, uint8_t* __restrict__ mask) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256)
    mask[i] = is_nz(vol[i]) ? 1 : 0;
}

// host side of numpy's pairwise tree: leaf boundaries, then the same recursion over the leaf sums
static void plan_leaves(int64_t off, int64_t n, std::vector<int64_t>& offs) {
  if (n <= 128) { offs.push_back(off); return; }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  plan_leaves(off, n2, offs);
  plan_leaves(off + n2, n - n2, offs);
}
template <typename A>
static A combine_leaves(int64_t n, const A* sums, size_t& next) {
  if (n <= 128) return sums[next++];
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  const volatile A l = combine_leaves<A>(n2, sums, next);
  const volatile A r = combine_leaves<A>(n - n2, sums, next);
  const volatile A s = l + r;
  return s;
}

template <typename T>
static int normalise_t(sc_ctx* ctx, const void* vol_v, const int32_t* dims, float* out, double* mean_std_host, cudaStream_t st) {
  typedef typename Acc<T>::type A;
  const T* vol = reinterpret_cast<const T*>(vol_v);
  const int64_t total = (int64_t)dims[0] * dims[1] * dims[2];
  const int nblocks = (int)((total + kPrepBlock - 1) / kPrepBlock);
  // scratch: counts | offsets | nzv (worst case: every voxel) | leaf offsets | leaf sums
  const size_t max_leaves = (size_t)(total / 64 + 2);
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_counts = 0, o_offsets = al((size_t)nblocks * 4), o_nzv = o_offsets + al((size_t)nblocks * 8);
  const size_t o_loff = o_nzv + al((size_t)total * sizeof(A)), o_lsum = o_loff + al((max_leaves + 1) * 8);
  SC_TRY(ensure_ws(ctx->ws_train, o_lsum + al(max_leaves * sizeof(A))));
  char* base = reinterpret_cast<char*>(ctx->ws_train.ptr);
  int32_t* counts = reinterpret_cast<int32_t*>(base + o_counts);
  int64_t* offsets = reinterpret_cast<int64_t*>(base + o_offsets);
  A* nzv = reinterpret_cast<A*>(base + o_nzv);
  int64_t* d_loff = reinterpret_cast<int64_t*>(base + o_loff);
  A* d_lsum = reinterpret_cast<A*>(base + o_lsum);
  ProfScope prof(ctx, PC_NONZERO, st);
  prep_count_kernel<T><<<nblocks, 256, 0, st>>>(vol, total, counts);
  prep_scan_kernel<<<1, 1024, 0, st>>>(counts, nblocks, offsets, ctx->d_count);
  prep_emit_kernel<T><<<nblocks, 256, 0, st>>>(vol, total, offsets, nzv);
  ctx->launches += 3;
  SC_CUDA(cudaMemcpyAsync(ctx->h_count, ctx->d_count, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  SC_CUDA(cudaStreamSynchronize(st));
  const int64_t n = *ctx->h_count;
  A mean, sd;
  if (n == 0) {                                   // numpy: mean of an empty selection is nan (with a warning)
    mean = sd = (A)NAN;
  } else {
    std::vector<int64_t> loff;
    loff.reserve((size_t)(n / 64 + 2));
    plan_leaves(0, n, loff);
    loff.push_back(n);
    const int nleaves = (int)loff.size() - 1;
    std::vector<A> lsum((size_t)nleaves);
    SC_CUDA(cudaMemcpyAsync(d_loff, loff.data(), loff.size() * 8, cudaMemcpyHostToDevice, st));
    const unsigned grid = (unsigned)(((int64_t)nleaves * 8 + 255) / 256);
    for (int pass = 0; pass < 2; ++pass) {
      prep_leaf_kernel<A><<<grid, 256, 0, st>>>(nzv, d_loff, nleaves, pass, pass ? mean : A(0), d_lsum);
      ctx->launches++;
      SC_CUDA(cudaMemcpyAsync(lsum.data(), d_lsum, (size_t)nleaves * sizeof(A), cudaMemcpyDeviceToHost, st));
      SC_CUDA(cudaStreamSynchronize(st));
      size_t next = 0;
      const A s = combine_leaves<A>(n, lsum.data(), next);
      const A q = (A)((double)s / (double)n);     // rounded once more to the working type (see the header comment)
      if (pass == 0) mean = q;
      else sd = (A)sqrt((double)q);               // sqrt of a float32 through float64 rounds to the correctly rounded sqrtf
    }
  }
  if (mean_std_host) { mean_std_host[0] = (double)mean; mean_std_host[1] = (double)sd; }
  if (out) {
    const int64_t blocks = (total + 255) / 256;
    const unsigned g = (unsigned)(blocks < (int64_t)ctx->sm_count * 16 ? blocks : (int64_t)ctx->sm_count * 16);
    prep_normalise_kernel<T><<<g, 256, 0, st>>>(vol, total, mean, sd, out);
    ctx->launches++;
  }
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int normalise_volume(sc_ctx* ctx, const void* vol, int dtype, const int32_t* dims, float* out, double* mean_std_host, cudaStream_t st) {
  switch (dtype) {
    case SC_DT_U8: return normalise_t<uint8_t>(ctx, vol, dims, out, mean_std_host, st);
    case SC_DT_I8: return normalise_t<int8_t>(ctx, vol, dims, out, mean_std_host, st);
    case SC_DT_U16: return normalise_t<uint16_t>(ctx, vol, dims, out, mean_std_host, st);
    case SC_DT_I16: return normalise_t<int16_t>(ctx, vol, dims, out, mean_std_host, st);
    case SC_DT_U32: return normalise_t<uint32_t>(ctx, vol, dims, out, mean_std_host, st);
    case SC_DT_I32: return normalise_t<int32_t>(ctx, vol, dims, out, mean_std_host, st);
    case SC_DT_F32: return normalise_t<float>(ctx, vol, dims, out, mean_std_host, st);
    case SC_DT_F64: return normalise_t<double>(ctx, vol, dims, out, mean_std_host, st);
  }
  set_error("sc_normalise_volume: unknown dtype code %d", dtype);
  return SC_ERR_ARG;
}

template <typename T>
static int mask_t(sc_ctx* ctx, const void* vol, int64_t total, uint8_t* mask, cudaStream_t st) {
  const int64_t blocks = (total + 255) / 256;
  const unsigned g = (unsigned)(blocks < (int64_t)ctx->sm_count * 16 ? blocks : (int64_t)ctx->sm_count * 16);
  prep_mask_kernel<T><<<g, 256, 0, st>>>(reinterpret_cast<const T*>(vol), total, mask);
  ctx->launches++;
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

int candidate_mask(sc_ctx* ctx, const void* vol, int dtype, const int32_t* dims, uint8_t* mask, cudaStream_t st) {
  const int64_t total = (int64_t)dims[0] * dims[1] * dims[2];
  switch (dtype) {
    case SC_DT_U8: return mask_t<uint8_t>(ctx, vol, total, mask, st);
    case SC_DT_I8: return mask_t<int8_t>(ctx, vol, total, mask, st);
    case SC_DT_U16: return mask_t<uint16_t>(ctx, vol, total, mask, st);
    case SC_DT_I16: return mask_t<int16_t>(ctx, vol, total, mask, st);
    case SC_DT_U32: return mask_t<uint32_t>(ctx, vol, total, mask, st);
    case SC_DT_I32: return mask_t<int32_t>(ctx, vol, total, mask, st);
    case SC_DT_F32: return mask_t<float>(ctx, vol, total, mask, st);
    case SC_DT_F64: return mask_t<double>(ctx, vol, total, mask, st);
  }
  set_error("sc_candidate_mask: unknown dtype code %d", dtype);
  return SC_ERR_ARG;
}

// bounding box (half-open) and count of the non-zero voxels of a uint8 mask
__global__ void __launch_bounds__(256) bbox_kernel(const uint8_t* __restrict__ mask, int X, int Y, int Z, int* __restrict__ box,
                                                   unsigned long long* __restrict__ count) {
  // one warp per (x, y) line of Z voxels
  const int64_t line = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (line >= (int64_t)X * Y) return;
  const int x = (int)(line / Y), y = (int)(line - (int64_t)x * Y);
  const uint8_t* p = mask + line * Z;
  int zmin = Z, zmax = -1, c = 0;
  for (int z = lane; z < Z; z += 32)
    if (p[z]) { zmin = min(zmin, z); zmax = max(zmax, z); ++c; }
  for (int o = 16; o; o >>= 1) {
    zmin = min(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
    zmax = max(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if (lane == 0 && c > 0) {
    atomicMin(&box[0], x); atomicMax(&box[1], x + 1);
    atomicMin(&box[2], y); atomicMax(&box[3], y + 1);
    atomicMin(&box[4], zmin); atomicMax(&box[5], zmax + 1);
    atomicAdd(count, (unsigned long long)c);
  }
}

int mask_bbox(sc_ctx* ctx, const uint8_t* mask, const int32_t* dims, int32_t* box_host, int64_t* count_host, cudaStream_t st) {
  SC_TRY(ensure_ws(ctx->ws_train, 256));
  int* d_box = reinterpret_cast<int*>(ctx->ws_train.ptr);
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(d_box + 8);
  int init[10] = {dims[0], 0, dims[1], 0, dims[2], 0, 0, 0, 0, 0};
  SC_CUDA(cudaMemcpyAsync(d_box, init, sizeof(init), cudaMemcpyHostToDevice, st));
  const int64_t lines = (int64_t)dims[0] * dims[1];
  bbox_kernel<<<(unsigned)((lines * 32 + 255) / 256), 256, 0, st>>>(mask, dims[0], dims[1], dims[2], d_box, d_cnt);
  ctx->launches++;
  int res[10];
  SC_CUDA(cudaMemcpyAsync(res, d_box, sizeof(res), cudaMemcpyDeviceToHost, st));
  SC_CUDA(cudaStreamSynchronize(st));
  unsigned long long cnt;
  memcpy(&cnt, res + 8, sizeof(cnt));
  if (count_host) *count_host = (int64_t)cnt;
  for (int i = 0; i < 6; ++i) box_host[i] = cnt ? res[i] : 0;
  return SC_OK;
}

}  // namespace sc
