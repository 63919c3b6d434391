// fq28_scan.cu -- device-wide exclusive prefix sums used by the record splitter
// (newline ranks, symbol offsets), the N-position packer and the bit packer.
// Three launches: per-tile scan + tile totals, single-CTA scan of the totals,
// add-back.  HBM-bound; 2 reads + 2 writes per element.
#include "fq28_internal.cuh"

namespace fq28 {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_IPT = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;

template <typename T>
__device__ __forceinline__ T warp_inclusive(T v, unsigned lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= (unsigned)d) v += o;
  }
  return v;
}

// exclusive scan of one value per thread across a 1024-thread block;
// returns the exclusive prefix, *total receives the block sum
template <typename T>
__device__ __forceinline__ T block_exclusive(T v, T *total, T *smem /*[32]*/) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nwarps = (blockDim.x + 31) >> 5;
  T inc = warp_inclusive(v, lane);
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    T w = lane < nwarps ? smem[lane] : T(0);
    T winc = warp_inclusive(w, lane);
    smem[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  T res = smem[warp] + inc - v;
  *total = smem[32];
  __syncthreads();
  return res;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tiles(const TIn *__restrict__ in, TOut *__restrict__ out, TOut *__restrict__ sums, size_t n) {
  __shared__ TOut sm[33];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_IPT;
  TOut v[SCAN_IPT];
  TOut tsum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_IPT; i++) {
    v[i] = (base + i < n) ? (TOut)in[base + i] : TOut(0);
    tsum += v[i];
  }
  TOut total;
  TOut ex = block_exclusive<TOut>(tsum, &total, sm);
#pragma unroll
  for (int i = 0; i < SCAN_IPT; i++) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_sums(TOut *__restrict__ sums, size_t n_tiles, TOut *__restrict__ grand_total) {
  __shared__ TOut sm[33];
  TOut carry = 0;
  for (size_t base = 0; base < n_tiles; base += SCAN_THREADS) {
    const size_t i = base + threadIdx.x;
    TOut v = i < n_tiles ? sums[i] : TOut(0);
    TOut total;
    TOut ex = block_exclusive<TOut>(v, &total, sm);
    if (i < n_tiles) sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *grand_total = carry;
}

template <typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_add(TOut *__restrict__ out, const TOut *__restrict__ sums, size_t n) {
  const TOut add = sums[blockIdx.x];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_IPT;
#pragma unroll
  for (int i = 0; i < SCAN_IPT; i++)
    if (base + i < n) out[base + i] += add;
}

template <typename TIn, typename TOut>
static int scan_impl(fq28_handle *h, const TIn *in, TOut *out, size_t n, bool side) {
  cudaStream_t strm = side ? h->side : h->stream;
  DevBuf &tmpb = side ? h->scan_tmp_side : h->scan_tmp;
  if (n == 0) {
    FQ28_CUDA(h, cudaMemsetAsync(out, 0, sizeof(TOut), strm));
    return FQ28_OK;
  }
  const size_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  FQ28_TRY(ensure(h, tmpb, (n_tiles + 1) * sizeof(TOut)));
  TOut *sums = tmpb.as<TOut>();
  k_scan_tiles<TIn, TOut><<<(unsigned)n_tiles, SCAN_THREADS, 0, strm>>>(in, out, sums, n);
  FQ28_LAUNCH_CHECK(h);
  k_scan_sums<TOut><<<1, SCAN_THREADS, 0, strm>>>(sums, n_tiles, out + n);
  FQ28_LAUNCH_CHECK(h);
  if (n_tiles > 1) {
    k_scan_add<TOut><<<(unsigned)n_tiles, SCAN_THREADS, 0, strm>>>(out, sums, n);
    FQ28_LAUNCH_CHECK(h);
  }
  return FQ28_OK;
}

int scan_exclusive_u16_to_u32(fq28_handle *h, const uint16_t *in, uint32_t *out, size_t n, bool side) {
  return scan_impl<uint16_t, uint32_t>(h, in, out, n, side);
}
int scan_exclusive_u32(fq28_handle *h, const uint32_t *in, uint32_t *out, size_t n, bool side) {
  return scan_impl<uint32_t, uint32_t>(h, in, out, n, side);
}
int scan_exclusive_u32_to_u64(fq28_handle *h, const uint32_t *in, uint64_t *out, size_t n, bool side) {
  return scan_impl<uint32_t, unsigned long long>(h, in, reinterpret_cast<unsigned long long *>(out), n, side);
}

}  // namespace fq28
