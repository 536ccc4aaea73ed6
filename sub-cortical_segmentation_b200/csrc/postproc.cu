// Connected-component post-processing on the device (reference: post_process_segmentation,
// cnn_cort/base.py:460-480).  Per class l = 1..14 the reference labels the 6-connected components of (seg == l)
// (scipy.ndimage.label: ids in raster order of first appearance), counts per component the voxels inside the registered
// sub-cortical mask, takes np.argmax of [0 (background), c_1, c_2, ...] (first maximum wins) and paints `labels == argmax`
// with l, classes in ascending order.  Quirks kept bit for bit: when no component of a class overlaps the mask
// (or the class is absent) argmax is 0 and EVERY voxel that is not of class l is painted with l; later classes overwrite.
//
// One union-find labelling pass serves all 14 classes (components of different classes are disjoint): the root of a
// component is its smallest linear index = the voxel scipy meets first, so "first maximum" = maximal count, then minimal root.
#include "common.cuh"

namespace sc {

__device__ __forceinline__ int cc_find(const int* parent, int v) {
  int p = parent[v];
  while (p != v) { v = p; p = parent[v]; }
  return v;
}
__device__ __forceinline__ void cc_union(int* parent, int a, int b) {
  while (true) {
    a = cc_find(parent, a);
    b = cc_find(parent, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&parent[b], a);      // roots only ever decrease: the final root is the minimal index
    if (old == b) return;
    b = old;
  }
}

__global__ void cc_init_kernel(const uint8_t* __restrict__ seg, int64_t total, int* __restrict__ parent, int* __restrict__ count) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t l = seg[v];
    parent[v] = (l >= 1 && l <= 14) ? (int)v : -1;
    count[v] = 0;
  }
}
__global__ void cc_merge_kernel(const uint8_t* __restrict__ seg, int X, int Y, int Z, int* __restrict__ parent) {
  const int64_t total = (int64_t)X * Y * Z;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t l = seg[v];
    if (l < 1 || l > 14) continue;
    const int z = (int)(v % Z), y = (int)((v / Z) % Y), x = (int)(v / ((int64_t)Y * Z));
    if (z + 1 < Z && seg[v + 1] == l) cc_union(parent, (int)v, (int)(v + 1));
    if (y + 1 < Y && seg[v + Z] == l) cc_union(parent, (int)v, (int)(v + Z));
    if (x + 1 < X && seg[v + (int64_t)Y * Z] == l) cc_union(parent, (int)v, (int)(v + (int64_t)Y * Z));
  }
}
__global__ void cc_flatten_count_kernel(const uint8_t* __restrict__ seg, const uint8_t* __restrict__ mask, int64_t total, int* __restrict__ parent,
                                        int* __restrict__ count) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
    if (parent[v] < 0) continue;
    const int r = cc_find(parent, (int)v);
    parent[v] = r;
    if (mask[v]) atomicAdd(&count[r], 1);
  }
}
// best[l] = max over the components of class l of (count << 32 | ~root): maximal count, then minimal root
__global__ void cc_best_kernel(const uint8_t* __restrict__ seg, int64_t total, const int* __restrict__ parent, const int* __restrict__ count,
                               unsigned long long* __restrict__ best) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
    if (parent[v] != (int)v) continue;
    const unsigned long long key = ((unsigned long long)(unsigned)count[v] << 32) | (unsigned long long)(0xffffffffu - (unsigned)v);
    atomicMax(&best[seg[v]], key);
  }
}
__global__ void cc_paint_kernel(const uint8_t* __restrict__ seg, int64_t total, const int* __restrict__ parent,
                                const unsigned long long* __restrict__ best, int any_other_voxel, uint8_t* __restrict__ out) {
  __shared__ unsigned long long sb[16];
  if (threadIdx.x < 16) sb[threadIdx.x] = best[threadIdx.x];
  __syncthreads();
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t s = seg[v];
    uint8_t o = 0;
    for (int l = 14; l >= 1; --l) {           // classes are painted in ascending order: the highest class that selects v wins
      const unsigned long long key = sb[l];
      bool sel;
      if ((key >> 32) == 0) sel = any_other_voxel && s != l;                                   // argmax == 0: `labels == 0`, everything that is not class l
      else sel = s == l && parent[v] == (int)(0xffffffffu - (unsigned)(key & 0xffffffffu));    // the winning component
      if (sel) { o = (uint8_t)l; break; }
    }
    out[v] = o;
  }
}

int post_process(sc_ctx* ctx, const uint8_t* seg, const uint8_t* mask, const int32_t* dims, uint8_t* out, cudaStream_t st) {
  const int64_t total = (int64_t)dims[0] * dims[1] * dims[2];
  SC_TRY(ensure_ws(ctx->ws_train, (size_t)total * 8 + 256));
  int* parent = reinterpret_cast<int*>(ctx->ws_train.ptr);
  int* count = parent + total;
  unsigned long long* best = reinterpret_cast<unsigned long long*>(count + total);
  const int64_t blocks = (total + 255) / 256;
  const unsigned g = (unsigned)(blocks < (int64_t)ctx->sm_count * 16 ? blocks : (int64_t)ctx->sm_count * 16);
  SC_CUDA(cudaMemsetAsync(best, 0, 16 * sizeof(unsigned long long), st));
  cc_init_kernel<<<g, 256, 0, st>>>(seg, total, parent, count);
  cc_merge_kernel<<<g, 256, 0, st>>>(seg, dims[0], dims[1], dims[2], parent);
  cc_flatten_count_kernel<<<g, 256, 0, st>>>(seg, mask, total, parent, count);
  cc_best_kernel<<<g, 256, 0, st>>>(seg, total, parent, count, best);
  // `labels == 0` is empty only if EVERY voxel of the volume belongs to the class; then np.unique(labels) has no 0 and index 0 is
  // component 1 -- a degenerate case (a volume filled with one structure) treated as "nothing else to paint"
  cc_paint_kernel<<<g, 256, 0, st>>>(seg, total, parent, best, 1, out);
  ctx->launches += 5;
  SC_CUDA(cudaGetLastError());
  return SC_OK;
}

}  // namespace sc
