// MBPO model rollouts kept on the device (SURVEY.md section 8f "next" #2): the bookkeeping around ModelEnv.step in
//   mbrl/algorithms/mbpo.py:31-63  rollout_model_and_populate_sac_buffer
// i.e. the `accum_dones` mask and the `x[~accum_dones]` selections handed to sac_buffer.add_batch, which the reference
// does in numpy after a device -> host copy of every step's outputs.  Here the k steps write straight into [k][B][..]
// staging buffers (b200pets_step's output pointers), the mask is one tiny kernel per step, and ONE ordered compaction
// at the end packs the alive transitions of all steps in (step, row) order -- exactly the rows, in exactly the order,
// of the reference's k add_batch calls -- so the host receives one dense D2H copy.
//
// HBM-bound byte work: the compaction reads every staged row once and writes the alive ones once; copies are
// row-cooperative (consecutive threads move consecutive floats of a row) so that both sides coalesce.
#include "common.cuh"

namespace {

constexpr int kChunk = 1024;  // rows per block

// alive[r] = !accum[r]; accum[r] |= done[r]      (mbpo.py:51-62: the mask is applied BEFORE it absorbs this step's dones)
__global__ void mbpo_mask_kernel(long long B, const uint8_t* __restrict__ done, uint8_t* __restrict__ accum,
                                 uint8_t* __restrict__ alive) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  const uint8_t acc = accum[r];
  alive[r] = acc ? 0 : 1;
  accum[r] = (acc | done[r]) ? 1 : 0;
}

// block (step i, chunk c): number of alive rows among rows [c * kChunk, (c + 1) * kChunk) of step i
__global__ void __launch_bounds__(kChunk) mbpo_count_kernel(long long B, int bps, const uint8_t* __restrict__ alive,
                                                            int* __restrict__ block_counts) {
  const int i = blockIdx.x / bps, c = blockIdx.x % bps;
  const long long r = (long long)c * kChunk + threadIdx.x;
  const int flag = (r < B && alive[(long long)i * B + r]) ? 1 : 0;
  const int n = __syncthreads_count(flag);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = n;
}

// exclusive scan of the block counts (one block), per-step totals and the grand total
__global__ void __launch_bounds__(kChunk) mbpo_scan_kernel(int nblocks, int bps, int steps, const int* __restrict__ block_counts,
                                                           long long* __restrict__ block_offsets, long long* __restrict__ counts) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += kChunk) {
    const int b = base + tid;
    const long long v = b < nblocks ? block_counts[b] : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      long long w = warp_tot[lane], winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long n = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += n;
      }
      warp_tot[lane] = winc - w;
    }
    __syncthreads();
    const long long excl = carry + warp_tot[warp] + inc - v;
    if (b < nblocks) block_offsets[b] = excl;
    __syncthreads();
    if (tid == kChunk - 1) carry = excl + v;
    __syncthreads();
  }
  // per-step totals: counts[i] = offset of the first block of step i + 1 minus that of step i
  for (int i = tid; i < steps; i += kChunk) {
    const long long lo = block_offsets[i * bps];
    const long long hi = (i + 1 < steps) ? block_offsets[(i + 1) * bps] : carry;
    counts[i] = hi - lo;
  }
  if (tid == 0) counts[steps] = carry;
}

struct CompactArgs {
  long long B;
  int bps, D, A;
  const float* obs0;      // [B][D] observations before step 0
  const float* act;       // [k][B][A]
  const float* next_obs;  // [k][B][D]
  const float* reward;    // [k][B]
  const uint8_t* done;    // [k][B]
  const uint8_t* alive;   // [k][B]
  const long long* block_offsets;
  float *obs_out, *act_out, *next_out, *rew_out;
  uint8_t* done_out;
};

__global__ void __launch_bounds__(kChunk) mbpo_scatter_kernel(const CompactArgs a) {
  __shared__ int dst[kChunk];  // destination row of each local row relative to this block's offset, -1 = dropped
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = blockIdx.x / a.bps, c = blockIdx.x % a.bps;
  const long long r0 = (long long)c * kChunk;
  const long long r = r0 + tid;
  const long long src_row = (long long)i * a.B + r;
  const int flag = (r < a.B && a.alive[src_row]) ? 1 : 0;
  int inc = flag;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    warp_tot[lane] = winc - w;
  }
  __syncthreads();
  const int rank = warp_tot[warp] + inc - flag;
  dst[tid] = flag ? rank : -1;
  const long long base = a.block_offsets[blockIdx.x];
  if (flag) {
    a.rew_out[base + rank] = a.reward[src_row];
    a.done_out[base + rank] = a.done[src_row];
  }
  __syncthreads();
  const int rows = (int)min((long long)kChunk, a.B - r0);
  // observation before step i: the initial batch for i == 0, else the previous step's prediction (mbpo.py:61)
  const float* obs_src = i == 0 ? a.obs0 + r0 * a.D : a.next_obs + ((long long)(i - 1) * a.B + r0) * a.D;
  const float* nxt_src = a.next_obs + ((long long)i * a.B + r0) * a.D;
  const float* act_src = a.act + ((long long)i * a.B + r0) * a.A;
  for (int e = tid; e < rows * a.D; e += kChunk) {
    const int lr = e / a.D, col = e - lr * a.D;
    const int d = dst[lr];
    if (d >= 0) {
      a.obs_out[(base + d) * a.D + col] = obs_src[e];
      a.next_out[(base + d) * a.D + col] = nxt_src[e];
    }
  }
  for (int e = tid; e < rows * a.A; e += kChunk) {
    const int lr = e / a.A, col = e - lr * a.A;
    const int d = dst[lr];
    if (d >= 0) a.act_out[(base + d) * a.A + col] = act_src[e];
  }
}

}  // namespace

static size_t al256m(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" {

int b200pets_mbpo_mask(int64_t batch, const uint8_t* done, uint8_t* accum_dones, uint8_t* alive, void* stream) {
  if (batch <= 0 || !done || !accum_dones || !alive) return b200pets_set_error(B200PETS_EINVAL, "mbpo_mask: bad argument");
  mbpo_mask_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, (cudaStream_t)stream>>>(batch, done, accum_dones, alive);
  CUDA_TRY(cudaGetLastError());
  return B200PETS_OK;
}

size_t b200pets_mbpo_compact_workspace_bytes(int32_t steps, int64_t batch) {
  if (steps <= 0 || batch <= 0) return 0;
  const size_t nblocks = (size_t)steps * (size_t)((batch + kChunk - 1) / kChunk);
  return al256m(nblocks * sizeof(int)) + al256m(nblocks * sizeof(long long));
}

int b200pets_mbpo_compact(int32_t steps, int64_t batch, int32_t obs_dim, int32_t act_dim, const float* obs0,
                          const float* act, const float* next_obs, const float* reward, const uint8_t* done,
                          const uint8_t* alive, float* obs_out, float* act_out, float* next_obs_out, float* reward_out,
                          uint8_t* done_out, int64_t* counts, void* workspace, size_t workspace_bytes, void* stream_) {
  if (steps <= 0 || batch <= 0 || obs_dim <= 0 || act_dim <= 0 || !obs0 || !act || !next_obs || !reward || !done || !alive ||
      !obs_out || !act_out || !next_obs_out || !reward_out || !done_out || !counts || !workspace)
    return b200pets_set_error(B200PETS_EINVAL, "mbpo_compact: bad argument");
  if (workspace_bytes < b200pets_mbpo_compact_workspace_bytes(steps, batch))
    return b200pets_set_error(B200PETS_EINVAL, "mbpo_compact: workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int bps = (int)((batch + kChunk - 1) / kChunk);
  const long long nblocks_ll = (long long)steps * bps;
  if (nblocks_ll > 0x7fffffffLL) return b200pets_set_error(B200PETS_EUNSUPPORTED, "mbpo_compact: too many rows");
  const int nblocks = (int)nblocks_ll;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  int* block_counts = reinterpret_cast<int*>(ws);
  long long* block_offsets = reinterpret_cast<long long*>(ws + al256m((size_t)nblocks * sizeof(int)));
  mbpo_count_kernel<<<nblocks, kChunk, 0, stream>>>(batch, bps, alive, block_counts);
  mbpo_scan_kernel<<<1, kChunk, 0, stream>>>(nblocks, bps, steps, block_counts, block_offsets,
                                              reinterpret_cast<long long*>(counts));
  CompactArgs a{};
  a.B = batch; a.bps = bps; a.D = obs_dim; a.A = act_dim;
  a.obs0 = obs0; a.act = act; a.next_obs = next_obs; a.reward = reward; a.done = done; a.alive = alive;
  a.block_offsets = block_offsets;
  a.obs_out = obs_out; a.act_out = act_out; a.next_out = next_obs_out; a.rew_out = reward_out; a.done_out = done_out;
  mbpo_scatter_kernel<<<nblocks, kChunk, 0, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return B200PETS_OK;
}

}  // extern "C"
