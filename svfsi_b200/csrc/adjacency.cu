// adjacency.cu -- setup-time construction (on the device) of the lists the gather assembly
// walks: for every Val block the (element, a, b) triples that contribute to it, for every node
// the (element, a) pairs, both in ascending element order (the reference's accumulation order,
// S/LHSA.f:275-295 inside the element loop S/FLUID.f:74-182), plus a block processing order that
// groups blocks of similar list length into the same warp (diagonal blocks collect ~24
// elements, off-diagonal ones 4-6).
#include <cuda_runtime.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>

#include "ctx.h"
#include "kernels.h"

namespace svfsi {

__global__ void count_kernel(size_t n, const int *__restrict__ key, int *__restrict__ cnt) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) atomicAdd(cnt + key[t], 1);
}
// payload = (t >> shift) << shift_keep | (t & mask): for blocks t = e*16 + blk -> e<<4|blk = t
__global__ void fill_kernel(size_t n, const int *__restrict__ key, const int *__restrict__ ptr,
                            int *__restrict__ cursor, int *__restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int k = key[t];
  const int pos = atomicAdd(cursor + k, 1);
  out[ptr[k] + pos] = (int)t;
}
// ascending sort of every (short) segment: one thread per segment, insertion sort
__global__ void segsort_kernel(int nseg, const int *__restrict__ ptr, int *__restrict__ v) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const int a = ptr[s], b = ptr[s + 1];
  for (int i = a + 1; i < b; i++) {
    const int x = v[i];
    int j = i - 1;
    while (j >= a && v[j] > x) { v[j + 1] = v[j]; j--; }
    v[j + 1] = x;
  }
}
// within chunks of CH consecutive blocks, order by descending list length (stable)
template <int CH>
__global__ void __launch_bounds__(CH) order_kernel(int n, const int *__restrict__ ptr,
                                                   int *__restrict__ order) {
  __shared__ int len[CH];
  const int base = blockIdx.x * CH, t = threadIdx.x, me = base + t;
  len[t] = (me < n) ? ptr[me + 1] - ptr[me] : -1;
  __syncthreads();
  if (me >= n) return;
  int rank = 0;
  const int mine = len[t];
  for (int u = 0; u < CH; u++) {
    const int l = len[u];
    rank += (l > mine) || (l == mine && u < t);
  }
  order[base + rank] = me;
}

static int build_lists(cudaStream_t st, size_t nItems, const int *key, int nseg, int **ptrOut,
                       int **listOut) {
  int *cnt = nullptr, *ptr = nullptr, *list = nullptr;
  CUDA_TRY(cudaMalloc(&cnt, sizeof(int) * ((size_t)nseg + 1)));
  CUDA_TRY(cudaMalloc(&ptr, sizeof(int) * ((size_t)nseg + 1)));
  CUDA_TRY(cudaMalloc(&list, sizeof(int) * (nItems ? nItems : 1)));
  CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)nseg + 1), st));
  const unsigned blocks = (unsigned)((nItems + 255) / 256);
  count_kernel<<<blocks, 256, 0, st>>>(nItems, key, cnt);
  thrust::exclusive_scan(thrust::cuda::par.on(st), cnt, cnt + nseg + 1, ptr);
  CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)nseg + 1), st));
  fill_kernel<<<blocks, 256, 0, st>>>(nItems, key, ptr, cnt, list);
  segsort_kernel<<<(nseg + 255) / 256, 256, 0, st>>>(nseg, ptr, list);
  count_launch(4);
  CUDA_TRY(cudaStreamSynchronize(st));
  cudaFree(cnt);
  *ptrOut = ptr;
  *listOut = list;
  return 0;
}

// Pair lists for the pair-owner gather: position g of the processing order holds block p = (r,c);
// it is kept if c >= r (or if the pattern has no (c,r) entry, which an element-built pattern never
// does), together with the position of (c,r).
__global__ void pair_flag_kernel(int nnz, const int *__restrict__ blkOrder,
                                 const int *__restrict__ rowOf, const int *__restrict__ col,
                                 const int *__restrict__ rowPtr, int *__restrict__ flag,
                                 int *__restrict__ tposOut) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nnz) return;
  const int p = blkOrder[g];
  const int r = rowOf[p], c = col[p];
  int l = -1;
  if (r == c) l = p;
  else
    for (int j = rowPtr[c]; j < rowPtr[c + 1]; j++)
      if (col[j] == r) { l = j; break; }
  tposOut[g] = l;
  flag[g] = (c >= r || l < 0) ? 1 : 0;
}
__global__ void pair_fill_kernel(int nnz, const int *__restrict__ blkOrder,
                                 const int *__restrict__ flag, const int *__restrict__ pos,
                                 const int *__restrict__ tpos, int *__restrict__ pairList,
                                 int *__restrict__ pairT) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nnz || !flag[g]) return;
  pairList[pos[g]] = blkOrder[g];
  pairT[pos[g]] = tpos[g];
}
int build_pair_lists(cudaStream_t st, int nnz, const int *blkOrder, const int *rowOf, const int *col,
                     const int *rowPtr, int **pairList, int **pairT, int *nPair) {
  *pairList = *pairT = nullptr;
  *nPair = 0;
  if (nnz <= 0) return 0;
  int *flag = nullptr, *pos = nullptr, *tp = nullptr;
  CUDA_TRY(cudaMalloc(&flag, sizeof(int) * (size_t)nnz));
  CUDA_TRY(cudaMalloc(&pos, sizeof(int) * (size_t)nnz));
  CUDA_TRY(cudaMalloc(&tp, sizeof(int) * (size_t)nnz));
  const unsigned blocks = (unsigned)((nnz + 255) / 256);
  pair_flag_kernel<<<blocks, 256, 0, st>>>(nnz, blkOrder, rowOf, col, rowPtr, flag, tp);
  thrust::exclusive_scan(thrust::cuda::par.on(st), flag, flag + nnz, pos);
  int lastPos = 0, lastFlag = 0;
  CUDA_TRY(cudaMemcpyAsync(&lastPos, pos + nnz - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&lastFlag, flag + nnz - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const int n = lastPos + lastFlag;
  CUDA_TRY(cudaMalloc(pairList, sizeof(int) * (size_t)(n ? n : 1)));
  CUDA_TRY(cudaMalloc(pairT, sizeof(int) * (size_t)(n ? n : 1)));
  pair_fill_kernel<<<blocks, 256, 0, st>>>(nnz, blkOrder, flag, pos, tp, *pairList, *pairT);
  count_launch(3);
  CUDA_TRY(cudaStreamSynchronize(st));
  cudaFree(flag); cudaFree(pos); cudaFree(tp);
  *nPair = n;
  return 0;
}

// .w = row + 1 for a diagonal block (its group also sums the residual of that row), else 0
__global__ void block_desc_kernel(int nnz, const int *__restrict__ blkOrder,
                                  const int *__restrict__ adjPtr, const int *__restrict__ rowOf,
                                  const int *__restrict__ col, int4 *__restrict__ desc) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nnz) return;
  const int p = blkOrder[g];
  const int r = rowOf ? rowOf[p] : -1;
  desc[g] = make_int4(p, adjPtr[p], adjPtr[p + 1], (rowOf && col[p] == r) ? r + 1 : 0);
}
int build_block_desc(cudaStream_t st, int nnz, const int *blkOrder, const int *adjPtr, int4 **desc,
                     const int *rowOf, const int *col) {
  *desc = nullptr;
  if (nnz <= 0) return 0;
  CUDA_TRY(cudaMalloc(desc, sizeof(int4) * (size_t)nnz));
  block_desc_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(nnz, blkOrder, adjPtr, rowOf, col, *desc);
  count_launch();
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// paired descriptors: unit g of the pair list (processing order: chunks of 512 blocks of neighbouring rows,
// longest lists first) is an off-diagonal pair -> two ADJACENT entries (r,c), (c,r), or a diagonal block / a
// block without a partner -> one entry.  Units keep the pair-list order, so the diagonal blocks of a row chunk
// are processed next to its pairs and find the element records in L2 (putting all diagonal blocks after all
// pairs doubled the DRAM reads of the gather kernel, profiles/r02_asm_gather5.md).
__global__ void pdesc_size_kernel(int nPair, const int *__restrict__ pairList, const int *__restrict__ pairT,
                                  int *__restrict__ size) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nPair) return;
  const int p = pairList[g], pt = pairT[g];
  size[g] = (pt >= 0 && pt != p) ? 2 : 1;
}
__global__ void pdesc_fill_kernel(int nPair, const int *__restrict__ pairList, const int *__restrict__ pairT,
                                  const int *__restrict__ pos, const int *__restrict__ adjPtr,
                                  const int *__restrict__ rowOf, int4 *__restrict__ desc) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nPair) return;
  const int p = pairList[g], pt = pairT[g];
  const int at = pos[g];
  if (pt >= 0 && pt != p) {
    desc[at] = make_int4(p, adjPtr[p], adjPtr[p + 1], 0);
    desc[at + 1] = make_int4(pt, adjPtr[pt], adjPtr[pt + 1], 0);
  } else {
    desc[at] = make_int4(p, adjPtr[p], adjPtr[p + 1], pt == p ? rowOf[p] + 1 : 0);
  }
}
int build_paired_desc(cudaStream_t st, int nPair, const int *pairList, const int *pairT, const int *adjPtr,
                      const int *rowOf, int nnz, int4 **desc, int *nOffEntries) {
  *desc = nullptr;
  *nOffEntries = 0;
  if (nPair <= 0) return 0;
  int *size = nullptr, *pos = nullptr;
  CUDA_TRY(cudaMalloc(&size, sizeof(int) * (size_t)nPair));
  CUDA_TRY(cudaMalloc(&pos, sizeof(int) * (size_t)nPair));
  const unsigned blocks = (unsigned)((nPair + 255) / 256);
  pdesc_size_kernel<<<blocks, 256, 0, st>>>(nPair, pairList, pairT, size);
  thrust::exclusive_scan(thrust::cuda::par.on(st), size, size + nPair, pos);
  int lastPos = 0, lastSize = 0;
  CUDA_TRY(cudaMemcpyAsync(&lastPos, pos + nPair - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&lastSize, size + nPair - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const int total = lastPos + lastSize;
  if (total != nnz) {   // a block (c,r), c < r, whose partner is missing was dropped from the pair list
    cudaFree(size); cudaFree(pos);
    return 0;           // caller falls back to the unpaired kernels
  }
  CUDA_TRY(cudaMalloc(desc, sizeof(int4) * (size_t)nnz));
  pdesc_fill_kernel<<<blocks, 256, 0, st>>>(nPair, pairList, pairT, pos, adjPtr, rowOf, *desc);
  count_launch(3);
  CUDA_TRY(cudaStreamSynchronize(st));
  cudaFree(size); cudaFree(pos);
  *nOffEntries = total - (2 * nPair - total);   // entries that belong to pairs: 2 * (total - nPair)
  return 0;
}

int build_gather_adjacency(cudaStream_t st, int nEl, int nNo, int nnz, const int *ien,
                           const int *edest, int **blkAdjPtr, int **blkAdj, int **nodeAdjPtr,
                           int **nodeAdj, int **blkOrder) {
  // item t = e*16 + (a*4+b) has key edest[t]; the stored payload t equals e<<4 | a<<2 | b
  if (int rc = build_lists(st, (size_t)nEl * 16, edest, nnz, blkAdjPtr, blkAdj)) return rc;
  // item t = e*4 + a has key ien[t]; payload t = e<<2 | a
  if (int rc = build_lists(st, (size_t)nEl * 4, ien, nNo, nodeAdjPtr, nodeAdj)) return rc;
  int *order = nullptr;
  constexpr int CH = 512;
  const int padded = ((nnz + CH - 1) / CH) * CH;
  CUDA_TRY(cudaMalloc(&order, sizeof(int) * (size_t)padded));
  order_kernel<CH><<<padded / CH, CH, 0, st>>>(nnz, *blkAdjPtr, order);
  count_launch();
  CUDA_TRY(cudaStreamSynchronize(st));
  *blkOrder = order;
  return 0;
}

}  // namespace svfsi
