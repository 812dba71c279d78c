// sgc_peer_allreduce: one-shot all-reduce over NVLink PEER MEMORY for the view-sharded cross-view statistics
// (partial sums / counts, score maxima, partial-softmax sums -- the log-sum-exp merge -- and in the
// backward the softmax-normaliser dot and the query gradient) and for the weight gradients of scene-batch data parallelism.
// Every rank exposes its partial in a SYMMETRIC buffer: one cudaMalloc per rank (sgc_peer_alloc) whose CUDA IPC handle the
// host side exchanges through torch.distributed and maps into every peer's address space (sgc_peer_open); the first
// kPeerSigBytes of the allocation are the signal pad, the rest is data.  This kernel
//   1. meets the peers at a barrier built from flags in the symmetric signal pads (system-scope CAS, one flag per
//      (CTA, peer): stateless, so the kernel can be replayed from a CUDA graph -- which an NCCL call in this stack cannot),
//   2. streams its slice of EVERY rank's buffer through 16-byte loads (peer loads bypass L1) and reduces in registers,
//   3. writes the result to a local tensor and meets the peers again, so that the symmetric buffer may be overwritten.
// One launch, no host involvement, no staging copies on the peers' side: the transfer IS the reduction's operand fetch.
#include "common.cuh"

#include <cstdlib>
#include <cstring>

namespace sgc {

constexpr int kPeerThreads = 512;
constexpr int kPeerMaxWorld = 8;
constexpr int kPeerSigOffset = 0;      // uint32 slot offset inside the signal pad the caller passes (one pad per channel)
constexpr int kPeerMaxBlocks = 128;
constexpr int kPeerStatusSlot = kPeerMaxBlocks * kPeerMaxWorld;     // uint32 after the flags: 1 = a barrier timed out
constexpr int kPeerSigBytes = kPeerMaxBlocks * kPeerMaxWorld * 4 + 256;   // flags (CTA, peer) -> uint32, then the status word
constexpr long long kPeerSpinCycles = 8000000000ll;                  // ~4 s at 1.9 GHz

struct PeerPtrs {
  const float* buf[kPeerMaxWorld];
  uint32_t* sig[kPeerMaxWorld];
};

__device__ __forceinline__ void peer_barrier(const PeerPtrs& pp, int rank, int world) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int peer = threadIdx.x;
    uint32_t* send = pp.sig[peer] + kPeerSigOffset + blockIdx.x * kPeerMaxWorld + rank;
    uint32_t* recv = pp.sig[rank] + kPeerSigOffset + blockIdx.x * kPeerMaxWorld + peer;
    __threadfence_system();
    // a peer that never arrives (died, different launch order) must not hang the GPU: give up after ~4 s and record it in
    // the status word of the OWN pad (PeerMemory.check() reads it); the result of this launch is then garbage
    const long long t0 = clock64();
    bool dead = false;
    while (atomicCAS_system(send, 0u, 1u) != 0u) {
      if (clock64() - t0 > kPeerSpinCycles) { dead = true; break; }
    }
    while (!dead && atomicCAS_system(recv, 1u, 0u) != 1u) {
      if (clock64() - t0 > kPeerSpinCycles) { dead = true; break; }
    }
    if (dead) atomicExch_system(pp.sig[rank] + kPeerStatusSlot, 1u);
    __threadfence_system();
  }
  __syncthreads();
}

// U float4 elements per thread and pass (all their loads -- U x world, (world - 1) of them over the links -- are in flight
// together): a collective confined to a few CTAs beside the step's large kernels is bound by link latency x passes.
// W = upper bound of `world` for this instantiation (U x W float4 registers hold one pass).
template <bool MAXOP, int U, int W>
__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(const __grid_constant__ PeerPtrs pp, int rank, int world,
                                                                      long long n, float scale, float* __restrict__ out) {
  peer_barrier(pp, rank, world);      // every rank's partial is complete and visible
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
    // fetched starting with the next rank (spreads the load over the links), REDUCED in rank order on every rank: all ranks
    // get bit-identical results, which the replicated voxel chain (and its deterministic top-k) relies on
    float4 v[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
#pragma unroll
      for (int r = 0; r < W; ++r) {
        if (r < world && i < n4) {
          const int p = rank + r < world ? rank + r : rank + r - world;
          v[u][r] = __ldcv(reinterpret_cast<const float4*>(pp.buf[p]) + i);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= n4) break;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < W; ++p) {
        if (p < world) {
          const int r = p >= rank ? p - rank : p - rank + world;     // slot that holds rank p's value
          float4 b = v[u][0];
#pragma unroll
          for (int q = 1; q < W; ++q) if (q == r) b = v[u][q];
          if (p == 0) a = b;
          else if (MAXOP) { a.x = fmaxf(a.x, b.x); a.y = fmaxf(a.y, b.y); a.z = fmaxf(a.z, b.z); a.w = fmaxf(a.w, b.w); }
          else { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
        }
      }
      if (!MAXOP) { a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale; }
      reinterpret_cast<float4*>(out)[i] = a;
    }
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      float a = __ldcv(pp.buf[0] + i);
      for (int p = 1; p < world; ++p) {
        const float b = __ldcv(pp.buf[p] + i);
        a = MAXOP ? fmaxf(a, b) : a + b;
      }
      out[i] = MAXOP ? a : a * scale;
    }
  }
  peer_barrier(pp, rank, world);      // nobody still reads this rank's buffer when the caller overwrites it
}

// Two-shot variant for more than two ranks and large payloads: every rank reduces ONE slice of the buffer (reading that
// slice from all ranks, writing the result into its own symmetric buffer in place), the ranks meet again, and every rank
// collects the world reduced slices: 2 (W-1)/W n words cross the links per rank instead of (W-1) n.  CTA b of every rank
// works on the same element subset of every slice in both phases, so the per-CTA flags are all the synchronisation needed.
struct PeerPtrsRW {
  float* buf[kPeerMaxWorld];
  uint32_t* sig[kPeerMaxWorld];
};

template <bool MAXOP>
__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_2shot_kernel(const __grid_constant__ PeerPtrsRW pw, int rank,
                                                                            int world, long long n, float scale,
                                                                            float* __restrict__ out) {
  PeerPtrs pp;
#pragma unroll
  for (int r = 0; r < kPeerMaxWorld; ++r) { pp.buf[r] = pw.buf[r]; pp.sig[r] = pw.sig[r]; }
  peer_barrier(pp, rank, world);
  const long long n4 = n >> 2;
  const long long per = (n4 + world - 1) / world;          // float4 elements per slice
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // phase 1: my slice, reduced in rank order, written back into MY buffer
  for (long long j = j0; j < per; j += stride) {
    const long long i = (long long)rank * per + j;
    if (i >= n4) break;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kPeerMaxWorld; ++p) {
      if (p < world) {
        const float4 b = __ldcv(reinterpret_cast<const float4*>(pw.buf[p]) + i);
        if (p == 0) a = b;
        else if (MAXOP) { a.x = fmaxf(a.x, b.x); a.y = fmaxf(a.y, b.y); a.z = fmaxf(a.z, b.z); a.w = fmaxf(a.w, b.w); }
        else { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
      }
    }
    if (!MAXOP) { a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale; }
    __stcg(reinterpret_cast<float4*>(pw.buf[rank]) + i, a);
  }
  peer_barrier(pp, rank, world);      // every rank's slice is reduced (the elements this CTA is about to read)
  // phase 2: collect the reduced slices (slice s from rank s), starting with the next rank; the loads of all slices are
  // issued before the first store (one link round trip per pass instead of one per slice)
  for (long long j = j0; j < per; j += stride) {
    float4 v[kPeerMaxWorld];
#pragma unroll
    for (int r = 0; r < kPeerMaxWorld; ++r) {
      if (r < world) {
        const int sidx = rank + r < world ? rank + r : rank + r - world;
        const long long i = (long long)sidx * per + j;
        if (i < n4) v[r] = __ldcv(reinterpret_cast<const float4*>(pw.buf[sidx]) + i);
      }
    }
#pragma unroll
    for (int r = 0; r < kPeerMaxWorld; ++r) {
      if (r < world) {
        const int sidx = rank + r < world ? rank + r : rank + r - world;
        const long long i = (long long)sidx * per + j;
        if (i < n4) reinterpret_cast<float4*>(out)[i] = v[r];
      }
    }
  }
  if (blockIdx.x == 0) {             // the (n mod 4) tail: nobody rewrote it, reduce it directly
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      float a = __ldcv(pw.buf[0] + i);
      for (int p = 1; p < world; ++p) {
        const float b = __ldcv(pw.buf[p] + i);
        a = MAXOP ? fmaxf(a, b) : a + b;
      }
      out[i] = MAXOP ? a : a * scale;
    }
  }
  peer_barrier(pp, rank, world);      // nobody still reads this rank's buffer when the caller overwrites it
}

// Gather / scatter between a list of tensors and a flat buffer as ONE small-footprint launch (128-thread CTAs fit next to the
// persistent tcgen05 kernels of the step's tail, which library copy kernels and 512-thread CTAs did not: the "overlapped"
// gradient average then ran only after them).
constexpr int kSegMax = 64;
struct CopySegs {
  const float* src[kSegMax];
  float* dst[kSegMax];
  long long n[kSegMax];
  int count;
};

__global__ void __launch_bounds__(128) copy_segs_kernel(const __grid_constant__ CopySegs cs) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (int k = 0; k < cs.count; ++k) {
    const float* __restrict__ src = cs.src[k];
    float* __restrict__ dst = cs.dst[k];
    const long long n = cs.n[k];
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
      const long long n4 = n >> 2;
      for (long long i = t0; i < n4; i += stride) reinterpret_cast<float4*>(dst)[i] = __ldcs(reinterpret_cast<const float4*>(src) + i);
      for (long long i = (n4 << 2) + t0; i < n; i += stride) dst[i] = src[i];
    } else {
      for (long long i = t0; i < n; i += stride) dst[i] = src[i];
    }
  }
}

// In-place all-reduce of a LIST of tensors as one launch: gather into the symmetric buffer, meet, reduce, scatter.  For the tail
// of the gradient average (the few tensors that are final only after the step's last kernel): three launches -- gather,
// all-reduce, scatter -- cost ~38 us there, this one ~15.  CTA b gathers exactly the flat elements CTA b of every peer reads
// after the barrier, so the per-CTA flags are all the synchronisation needed.  Every tensor starts at a float4 boundary of the
// flat space; the pad behind its last element is zero.
struct SegTable {
  float* ptr[kSegMax];
  long long off4[kSegMax + 1];   // float4 offset of segment k in the flat space; off4[count] = total
  long long n[kSegMax];
  int count;
};

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_segs_kernel(const __grid_constant__ PeerPtrsRW pw,
                                                                          const __grid_constant__ SegTable st, int rank, int world,
                                                                          float scale) {
  PeerPtrs pp;
#pragma unroll
  for (int r = 0; r < kPeerMaxWorld; ++r) { pp.buf[r] = pw.buf[r]; pp.sig[r] = pw.sig[r]; }
  const long long n4 = st.off4[st.count];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // gather: my elements of the flat space from the tensors into MY symmetric buffer (kept in registers for the first pass)
  for (long long i = i0; i < n4; i += stride) {
    int k = 0;
    while (k + 1 < st.count && st.off4[k + 1] <= i) ++k;
    const long long e = (i - st.off4[k]) * 4;          // element offset inside segment k
    const float* src = st.ptr[k] + e;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long left = st.n[k] - e;
    if (left >= 4 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) v = *reinterpret_cast<const float4*>(src);
    else {
      if (left > 0) v.x = src[0];
      if (left > 1) v.y = src[1];
      if (left > 2) v.z = src[2];
      if (left > 3) v.w = src[3];
    }
    __stcg(reinterpret_cast<float4*>(pw.buf[rank]) + i, v);
  }
  peer_barrier(pp, rank, world);
  for (long long i = i0; i < n4; i += stride) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kPeerMaxWorld; ++p) {
      if (p < world) {
        const float4 b = __ldcv(reinterpret_cast<const float4*>(pw.buf[p]) + i);
        if (p == 0) a = b;
        else { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
      }
    }
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
    int k = 0;
    while (k + 1 < st.count && st.off4[k + 1] <= i) ++k;
    const long long e = (i - st.off4[k]) * 4;
    float* dst = st.ptr[k] + e;
    const long long left = st.n[k] - e;
    if (left >= 4 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) *reinterpret_cast<float4*>(dst) = a;
    else {
      if (left > 0) dst[0] = a.x;
      if (left > 1) dst[1] = a.y;
      if (left > 2) dst[2] = a.z;
      if (left > 3) dst[3] = a.w;
    }
  }
  peer_barrier(pp, rank, world);      // nobody still reads this rank's buffer when the caller overwrites it
}

}  // namespace sgc

// bufs / sigs: HOST arrays of `world` device pointers (rank r's symmetric buffer / signal pad as mapped in THIS process).
// The reduction runs in rank order on every rank: all ranks obtain bit-identical results.  n floats, 16-byte aligned buffers.
// op 0: out = scale * sum over ranks (scale = 1 / world averages gradients); op 1: out = max over ranks (scale ignored).
extern "C" int sgc_peer_copy_segments(const void* const* src, void* const* dst, const long long* n, int count, int max_blocks,
                                      void* stream) {
  using namespace sgc;
  if (count < 0 || (count && (!src || !dst || !n))) return (int)cudaErrorInvalidValue;
  for (int k0 = 0; k0 < count; k0 += kSegMax) {
    CopySegs cs;
    cs.count = count - k0 < kSegMax ? count - k0 : kSegMax;
    long long total = 0;
    for (int k = 0; k < cs.count; ++k) {
      cs.src[k] = reinterpret_cast<const float*>(src[k0 + k]);
      cs.dst[k] = reinterpret_cast<float*>(dst[k0 + k]);
      cs.n[k] = n[k0 + k];
      if (!cs.src[k] || !cs.dst[k] || cs.n[k] < 0) return (int)cudaErrorInvalidValue;
      total += cs.n[k];
    }
    long long blocks = (total / 4 + 127) / 128;
    const long long cap = max_blocks > 0 ? max_blocks : 148 * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    copy_segs_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(cs);
    SGC_CUDA_CHECK_LAST();
  }
  return 0;
}

// block_threads: 0 = 512; 128 for a collective that has to fit next to the step's large kernels.
extern "C" int sgc_peer_allreduce(const void* const* bufs, void* const* sigs, int rank, int world, long long n, int op, float scale,
                                  float* out, int max_blocks, int block_threads, void* stream) {
  using namespace sgc;
  if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world || n < 0 || (op != 0 && op != 1) || !out) return (int)cudaErrorInvalidValue;
  if (n == 0) return 0;
  const int threads = block_threads > 0 ? block_threads : kPeerThreads;
  if (threads < 32 || threads > kPeerThreads || (threads & 31)) return (int)cudaErrorInvalidValue;
  PeerPtrs pp;
  for (int r = 0; r < kPeerMaxWorld; ++r) {
    pp.buf[r] = r < world ? reinterpret_cast<const float*>(bufs[r]) : nullptr;
    pp.sig[r] = r < world ? reinterpret_cast<uint32_t*>(sigs[r]) : nullptr;
    if (r < world && (!pp.buf[r] || !pp.sig[r] || (reinterpret_cast<uintptr_t>(pp.buf[r]) & 15))) return (int)cudaErrorInvalidValue;
  }
  if (reinterpret_cast<uintptr_t>(out) & 15) return (int)cudaErrorInvalidValue;
  // more than two ranks and a payload worth a second meeting: two-shot (SGC_PEER_TWO_SHOT_MIN_BYTES, default 256 KB)
  static const long long two_shot_min = getenv("SGC_PEER_TWO_SHOT_MIN_BYTES") ? atoll(getenv("SGC_PEER_TWO_SHOT_MIN_BYTES")) : (256ll << 10);
  const bool two_shot = world > 2 && n * 4 >= two_shot_min;
  const long long work4 = two_shot ? ((n >> 2) + world - 1) / world : (n >> 2);
  long long blocks = (work4 + threads - 1) / threads;
  if (blocks < 1) blocks = 1;
  // max_blocks (0 = 128): a collective that runs BESIDE the step's large kernels asks for a few CTAs only -- a spinning CTA of
  // 512 threads keeps half an SM's registers from them
  const long long cap = max_blocks > 0 && max_blocks < kPeerMaxBlocks ? max_blocks : kPeerMaxBlocks;
  if (blocks > cap) blocks = cap;
  // every rank must launch the SAME grid (the barrier pairs CTA b with CTA b of the peers): it depends on n and world only
  if (two_shot) {
    PeerPtrsRW pw;
    for (int r = 0; r < kPeerMaxWorld; ++r) { pw.buf[r] = const_cast<float*>(pp.buf[r]); pw.sig[r] = pp.sig[r]; }
    if (op == 1) peer_allreduce_2shot_kernel<true><<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(pw, rank, world, n, scale, out);
    else peer_allreduce_2shot_kernel<false><<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(pw, rank, world, n, scale, out);
  } else {
    // passes per thread decide: with few CTAs (or few ranks, i.e. few registers per element) fetch several elements at once
    const long long passes = (work4 + blocks * threads - 1) / (blocks * threads);
    const int U = passes >= 4 && world <= 2 ? 4 : (passes >= 2 && world <= 4 ? 2 : 1);
    cudaStream_t st = (cudaStream_t)stream;
#define SGC_PEER_ONE_SHOT(MAXOP, UU) peer_allreduce_kernel<MAXOP, UU, 8 / UU><<<(int)blocks, threads, 0, st>>>(pp, rank, world, n, scale, out)
    if (op == 1) { if (U == 4) SGC_PEER_ONE_SHOT(true, 4); else if (U == 2) SGC_PEER_ONE_SHOT(true, 2); else SGC_PEER_ONE_SHOT(true, 1); }
    else { if (U == 4) SGC_PEER_ONE_SHOT(false, 4); else if (U == 2) SGC_PEER_ONE_SHOT(false, 2); else SGC_PEER_ONE_SHOT(false, 1); }
#undef SGC_PEER_ONE_SHOT
  }
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Symmetric allocations.  sgc_peer_alloc: cudaMalloc + zero fill (the flags must start at 0) + the 64-byte CUDA IPC handle
// the peers open.  The memory is NOT torch's: the owner frees it with sgc_peer_free after every peer has closed its mapping.
extern "C" int sgc_peer_sig_bytes() { return sgc::kPeerSigBytes; }
// byte offset of the status word inside the signal pad (0 = fine, 1 = a barrier gave up waiting for a peer)
extern "C" int sgc_peer_status_offset() { return sgc::kPeerStatusSlot * 4; }

extern "C" int sgc_peer_alloc(long long bytes, void** ptr, void* handle64) {
  if (bytes <= 0 || !ptr || !handle64) return (int)cudaErrorInvalidValue;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return (int)e; }
  memcpy(handle64, &h, sizeof(h));
  *ptr = p;
  return 0;
}

extern "C" int sgc_peer_open(const void* handle64, void** ptr) {
  if (!handle64 || !ptr) return (int)cudaErrorInvalidValue;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" int sgc_peer_close(void* ptr) { return ptr ? (int)cudaIpcCloseMemHandle(ptr) : 0; }
extern "C" int sgc_peer_free(void* ptr) { return ptr ? (int)cudaFree(ptr) : 0; }

// tensors[k] (n[k] floats each, k < count <= 64; HOST arrays) <- scale * sum over the ranks, in place, one launch.  The flat
// space (every tensor padded to a multiple of 4 floats) has to fit the symmetric data buffers; every rank passes the same
// sizes in the same order.
extern "C" int sgc_peer_allreduce_tensors(const void* const* bufs, void* const* sigs, int rank, int world, void* const* tensors,
                                          const long long* n, int count, float scale, void* stream) {
  using namespace sgc;
  if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world || count < 0 || count > kSegMax) return (int)cudaErrorInvalidValue;
  if (count == 0) return 0;
  PeerPtrsRW pw;
  for (int r = 0; r < kPeerMaxWorld; ++r) {
    pw.buf[r] = r < world ? reinterpret_cast<float*>(const_cast<void*>(bufs[r])) : nullptr;
    pw.sig[r] = r < world ? reinterpret_cast<uint32_t*>(sigs[r]) : nullptr;
    if (r < world && (!pw.buf[r] || !pw.sig[r] || (reinterpret_cast<uintptr_t>(pw.buf[r]) & 15))) return (int)cudaErrorInvalidValue;
  }
  SegTable st;
  st.count = count;
  long long off = 0;
  for (int k = 0; k < count; ++k) {
    if (!tensors[k] || n[k] <= 0 || (reinterpret_cast<uintptr_t>(tensors[k]) & 3)) return (int)cudaErrorInvalidValue;
    st.ptr[k] = reinterpret_cast<float*>(tensors[k]);
    st.n[k] = n[k];
    st.off4[k] = off;
    off += (n[k] + 3) / 4;
  }
  st.off4[count] = off;
  long long blocks = (off + kPeerThreads - 1) / kPeerThreads;
  if (blocks > kPeerMaxBlocks) blocks = kPeerMaxBlocks;
  peer_allreduce_segs_kernel<<<(int)blocks, kPeerThreads, 0, (cudaStream_t)stream>>>(pw, st, rank, world, scale);
  SGC_CUDA_CHECK_LAST();
  return 0;
}
