// Tensor-core rollout of the ensemble MLP on sm_100a: bf16 operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM walks 128-row tiles through steps [t0, t1) of the horizon without leaving the
// chip: the row state (observation, return, dead flag) stays in shared memory / registers, each layer is a
// chain of tcgen05.mma (M=128, N=padded layer width, K=16) whose A operand (activations) is written by the
// epilogue warps straight into the UMMA canonical shared-memory layout and whose B operand (weights) is
// streamed from the L2-resident packed image through a ring of 1-D TMA bulk copies.
//
// Warp roles (64 + 128 * CS threads):  warp 0 = weight producer (cp.async.bulk + mbarrier),
//                            warp 1 = TMEM allocator + MMA issuer (single thread),
//                            warps 2.. = 4 * CS epilogue warps: warp (q, cs) <-> TMEM lanes 32q..32q+31 (tile rows)
//                            and every CS-th 16-column chunk; thread (row, cs == 0) owns the row's scalar state.
//
// Bias is folded into the GEMM: every A tile carries two constant-one columns after the real inputs and the
// packed weight image carries bf16(b) and bf16(b - bf16(b)) in the matching K rows (api.cu pack kernel).
//
// Reference semantics restated: see rollout_f32.cu header (same per-row maths, same file:line anchors).
#include "common.cuh"
#include "sm100.cuh"

using namespace sm100;

struct TcPlan {
  int nlayers;
  int nblk[B200PETS_MAX_LAYERS];  // K-blocks (64 K-elements) per layer
  uint32_t stage_bytes;
  int nstages;
  uint32_t off_A, off_ring, off_obs, off_act, off_const, off_bar;
  uint32_t tmem_cols;
  int obs_ld, act_ld;
  uint32_t smem_bytes;
};

namespace {

constexpr int kEpiSplit = 4;  // column splits of the epilogue (16 epilogue warps)
constexpr int kTileM = 128;
constexpr int kMaxStages = 8;

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int ACT>
__device__ __forceinline__ float act_tc(float x, float slope) {
  if (ACT == B200PETS_ACT_SILU) {  // x * sigmoid(x) = h + h * tanh(h), h = x / 2  (one MUFU op)
    float h = 0.5f * x;
    return fmaf(h, tanh_approx(h), h);
  }
  if (ACT == B200PETS_ACT_RELU) return fmaxf(x, 0.f);
  return x > 0.f ? x : x * slope;
}

// byte offset of the 16-byte chunk (row i, k-chunk kc) in the A tile: [kc][i / 8][i % 8][8 x bf16]
__device__ __forceinline__ uint32_t a_chunk_off(int i, int kc) { return (uint32_t)((kc * 16 + (i >> 3)) * 128 + (i & 7) * 16); }

// CS = column splits of the epilogue: 4 * CS epilogue warps; warp (q, cs) owns TMEM lane quadrant q (rows
// 32q..32q+31) and every CS-th 16-column chunk.  Thread (row i, cs == 0) also owns the row's scalar state.
template <int ACT, int CS>
__global__ void __launch_bounds__(64 + 128 * CS, 1)
rollout_tc_kernel(const ModelDev m, const RolloutArgs a, const TcPlan p, const long long num_tiles) {
  constexpr int kEpiThreads = 128 * CS;
  constexpr int kThreadsAll = 64 + kEpiThreads;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* A_s = smem + p.off_A;
  uint8_t* ring = smem + p.off_ring;
  float* obs_s = reinterpret_cast<float*>(smem + p.off_obs);
  float* act_s = reinterpret_cast<float*>(smem + p.off_act);
  float* c_mean = reinterpret_cast<float*>(smem + p.off_const);
  float* c_istd = c_mean + m.in;
  float* c_minlv = c_istd + m.in;
  float* c_maxlv = c_minlv + m.out;
  float* c_nodelta = c_maxlv + m.out;  // [D] 1.0 = keep raw prediction
  float* rew_s = c_nodelta + m.D;      // [128] learned-reward column of the current step
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* bar_empty = bar_full + kMaxStages;
  uint64_t* bar_a_ready = bar_empty + kMaxStages;
  uint64_t* bar_acc = bar_a_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.nstages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(bar_a_ready, kEpiThreads);
    mbar_init(bar_acc, 1);
    mbar_fence_init();
  }
  for (int j = threadIdx.x; j < m.in; j += kThreadsAll) {
    c_mean[j] = m.norm_mode ? m.norm_mean_f[j] : 0.f;
    c_istd[j] = m.norm_mode ? m.norm_istd_f[j] : 1.f;
  }
  for (int j = threadIdx.x; j < m.out; j += kThreadsAll) {
    c_minlv[j] = m.deterministic ? 0.f : m.min_lv[j];
    c_maxlv[j] = m.deterministic ? 0.f : m.max_lv[j];
  }
  for (int j = threadIdx.x; j < m.D; j += kThreadsAll) c_nodelta[j] = (!m.target_is_delta || m.no_delta[j]) ? 1.f : 0.f;
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool shuffle = a.slot_mode >= 1;
  const long long Bm = shuffle ? 0 : a.B / m.M;
  const int tpm = shuffle ? 1 : (int)((Bm + kTileM - 1) / kTileM);
  const int nlayers = p.nlayers;

  if (warp == 0) {
    // =========================== weight producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int member = shuffle ? 0 : (int)(tile / tpm);
        for (int t = a.t0; t < a.t1; ++t) {
          const int mem = shuffle ? shuffle_member(a, (int)tile, t, m.M) : member;
          const uint8_t* base = m.img + (size_t)(blockIdx.x % m.img_replicas) * m.img_replica_stride + (size_t)mem * m.img_member_stride;
          for (int l = 0; l < nlayers; ++l) {
            const uint8_t* lsrc = base + m.img_layer_off[l];
            const uint32_t col_bytes = (uint32_t)m.Np[l] * 16u;  // one K core column (8 K-rows x Np)
            for (int j = 0; j < p.nblk[l]; ++j) {
              const int kb = min(64, m.Kp[l] - 64 * j);
              const int ncol = kb >> 3;
              const uint32_t bytes = (uint32_t)ncol * col_bytes;
              mbar_wait(&bar_empty[stage], phase ^ 1u);
              mbar_arrive_expect_tx(&bar_full[stage], bytes);
              uint8_t* dst = ring + (size_t)stage * p.stage_bytes;
              const uint8_t* src = lsrc + (size_t)64 * j * m.Np[l] * 2;
              // the columns of a stage may land in any order: start at a CTA-dependent column so that CTAs streaming
              // the same member do not walk the same L2 lines in lockstep
              int c = (int)((blockIdx.x >> 3) % ncol);
              for (int q = 0; q < ncol; ++q) {
                bulk_g2s(dst + (size_t)c * col_bytes, src + (size_t)c * col_bytes, col_bytes, &bar_full[stage]);
                if (++c == ncol) c = 0;
              }
              if (++stage == S) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, a_par = 0;
      const uint32_t A_addr = smem_u32(A_s);
      const uint32_t ring_addr = smem_u32(ring);
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int t = a.t0; t < a.t1; ++t) {
          for (int l = 0; l < nlayers; ++l) {
            const uint32_t np = (uint32_t)m.Np[l];
            const uint32_t idesc = umma_idesc_bf16_m128(np);
            const uint32_t b_lbo = np * 16u;
            const bool stamp = a.timeline && blockIdx.x == 0 && t == a.t0 + 5 && tile == blockIdx.x;
            if (stamp) a.timeline[64 + l * 4 + 0] = clock64();
            mbar_wait(bar_a_ready, a_par);
            a_par ^= 1u;
            tc_fence_after();
            if (stamp) a.timeline[64 + l * 4 + 1] = clock64();
            for (int j = 0; j < p.nblk[l]; ++j) {
              const int kb = min(64, m.Kp[l] - 64 * j);
              mbar_wait(&bar_full[stage], phase);
              tc_fence_after();
              if (stamp && j == p.nblk[l] - 1) a.timeline[64 + l * 4 + 2] = clock64();
              const uint32_t b_addr = ring_addr + (uint32_t)stage * p.stage_bytes;
              for (int kk = 0; kk < kb / 16; ++kk) {
                const uint32_t k0 = (uint32_t)(64 * j + 16 * kk);
                const uint64_t adesc = umma_smem_desc(A_addr + (k0 >> 3) * 2048u, 2048u, 128u);
                const uint64_t bdesc = umma_smem_desc(b_addr + (uint32_t)(2 * kk) * b_lbo, b_lbo, 128u);
                umma_bf16_ss(tmem_base, adesc, bdesc, idesc, (j | kk) != 0 ? 1u : 0u);
              }
              umma_commit(&bar_empty[stage]);
              if (++stage == S) { stage = 0; phase ^= 1u; }
            }
            umma_commit(bar_acc);
            if (stamp) a.timeline[64 + l * 4 + 3] = clock64();
          }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue ===========================
    const int q = warp & 3;            // TMEM lane quadrant this warp may access (hardware: warp id % 4)
    const int cs = (warp - 2) >> 2;    // column split
    const int i = q * 32 + lane;       // tile row
    const bool owner = cs == 0;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* my_obs = obs_s + i * p.obs_ld;
    float* my_act = act_s + i * p.act_ld;
    uint32_t acc_par = 0;
    const int Kp0 = m.Kp[0];
    const int ngroups = (m.out + 3) >> 2;
    auto epi_bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); };

    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      long long slot0;
      int nv;
      if (shuffle) {
        slot0 = tile * kTileM;
        nv = (int)min((long long)kTileM, a.B - slot0);
      } else {
        const int member = (int)(tile / tpm);
        const int c = (int)(tile % tpm);
        slot0 = (long long)member * Bm + (long long)c * kTileM;
        nv = (int)min((long long)kTileM, Bm - (long long)c * kTileM);
      }
      const bool valid = i < nv;
      const long long rid = valid ? slot_to_rid(a, slot0 + i) : 0;
      float tot = 0.f;
      int dead = 0;
      epi_bar();  // previous tile fully consumed before its row state is overwritten
      if (owner) {
        for (int d = 0; d < m.D; ++d) {
          float v = 0.f;
          if (valid) v = a.init_from_obs0 ? a.obs0[d] : a.obs_in[rid * m.D + d];
          my_obs[d] = v;
        }
        if (a.load_state && valid) {
          tot = a.total_state[rid];
          dead = a.dead_state[rid];
        }
      }
      const float* act_row = a.act + (rid / a.act_div) * a.act_row_stride;
      const bool act_regs = m.A <= 8;  // next-step actions prefetched into registers (hidden behind the layers)
      float an[8];
      if (owner) {
        const float* ap = act_row + (long long)a.t0 * a.act_t_stride;
#pragma unroll
        for (int j = 0; j < 8; ++j) an[j] = (valid && j < m.A) ? ap[j] : 0.f;
      }

      for (int t = a.t0; t < a.t1; ++t) {
        const bool stamp = a.timeline && blockIdx.x == 0 && warp == 2 && lane == 0 && t == a.t0 + 5 && tile == blockIdx.x;
        int sp = 0;
        if (stamp) a.timeline[sp++] = clock64();  // 0: step start
        if (owner) {
          if (act_regs) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < m.A) my_act[j] = an[j];
          } else {
            const float* ap = act_row + (long long)t * a.act_t_stride;
            for (int j = 0; j < m.A; ++j) my_act[j] = valid ? ap[j] : 0.f;
          }
        }
        epi_bar();
        if (stamp) a.timeline[sp++] = clock64();  // 1: actions loaded + barrier
        // ---- layer-0 operand: normalise(cat(proc(obs), act)), two constant-one bias columns, zero pad ----
        for (int kc = cs; kc < Kp0 / 8; kc += CS) {
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = kc * 8 + e;
            float v;
            if (j < m.Dp) {
              v = proc_obs_elem(my_obs, j, m.obs_process);
              v = (v - c_mean[j]) * c_istd[j];
            } else if (j < m.in) {
              v = (my_act[j - m.Dp] - c_mean[j]) * c_istd[j];
            } else {
              v = (j < m.in + 2) ? 1.f : 0.f;
            }
            x[e] = v;
          }
          uint4 pk = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
          *reinterpret_cast<uint4*>(A_s + a_chunk_off(i, kc)) = pk;
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_a_ready);
        if (stamp) a.timeline[sp++] = clock64();  // 2: input built
        // work hidden behind the first MMAs: next step's actions (global loads) and this step's model noise
        if (owner && act_regs && t + 1 < a.t1) {
          const float* ap = act_row + (long long)(t + 1) * a.act_t_stride;
#pragma unroll
          for (int j = 0; j < 8; ++j) an[j] = (valid && j < m.A) ? ap[j] : 0.f;
        }
        const bool draw = !m.deterministic && a.sample;
        float zpre[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int g = cs + u * CS;
          if (draw && !a.eps && g < ngroups)
            philox_normal4((uint32_t)rid, (uint32_t)t, RNG_STREAM_EPS | (uint32_t)g, (uint32_t)a.offset, a.seed, zpre[u]);
        }

        // ---- hidden layers: TMEM accumulator -> activation -> bf16 -> next A operand ----
        for (int l = 0; l < nlayers - 1; ++l) {
          const int np = m.Np[l];
          const int n_true = m.N[l];
          const int kp_next = m.Kp[l + 1];
          mbar_wait(bar_acc, acc_par);
          acc_par ^= 1u;
          tc_fence_after();
          if (stamp) a.timeline[sp++] = clock64();  // 3 + 2l: accumulator ready
          for (int c = cs; c < kp_next / 16; c += CS) {
            float v[16];
            if (16 * c < np) {
              uint32_t r[16];
              tmem_ld16(t_lane + (uint32_t)(16 * c), r);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = act_tc<ACT>(__uint_as_float(r[e]), m.leaky);
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = 0.f;
            }
            if (16 * c + 15 >= n_true && 16 * c <= n_true + 1) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int col = 16 * c + e;
                if (col == n_true || col == n_true + 1) v[e] = 1.f;
              }
            }
            uint4 p0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            uint4 p1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
            *reinterpret_cast<uint4*>(A_s + a_chunk_off(i, 2 * c)) = p0;
            *reinterpret_cast<uint4*>(A_s + a_chunk_off(i, 2 * c + 1)) = p1;
          }
          tc_fence_before();
          fence_proxy_async_smem();
          mbar_arrive(bar_a_ready);
          if (stamp) a.timeline[sp++] = clock64();  // 4 + 2l: activations written
        }

        // ---- output layer: groups of 4 outputs; Gaussian sample, delta add-back ----
        mbar_wait(bar_acc, acc_par);
        acc_par ^= 1u;
        tc_fence_after();
        if (stamp) a.timeline[sp++] = clock64();  // output accumulator ready
        for (int g = cs; g < ngroups; g += CS) {
          uint32_t rm[4], rl[4] = {0u, 0u, 0u, 0u};
          tmem_ld4(t_lane + (uint32_t)(4 * g), rm);
          if (!m.deterministic) tmem_ld4(t_lane + (uint32_t)(m.outp + 4 * g), rl);
          tmem_ld_wait();
          float z[4] = {0.f, 0.f, 0.f, 0.f};
          if (draw && !a.eps) {
            const int u = (g - cs) / CS;
            if (u < 2) {
#pragma unroll
              for (int e = 0; e < 4; ++e) z[e] = u == 0 ? zpre[0][e] : zpre[1][e];
            } else {
              philox_normal4((uint32_t)rid, (uint32_t)t, RNG_STREAM_EPS | (uint32_t)g, (uint32_t)a.offset, a.seed, z);
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int o = 4 * g + e;
            if (o < m.out) {
              const float mean = __uint_as_float(rm[e]);
              float pred = mean;
              if (draw) {
                float lv = __uint_as_float(rl[e]);
                const float mx = c_maxlv[o], mn = c_minlv[o];
                float d1 = mx - lv;
                lv = mx - (d1 > 20.f ? d1 : __logf(1.f + __expf(d1)));
                float d2 = lv - mn;
                lv = mn + (d2 > 20.f ? d2 : __logf(1.f + __expf(d2)));
                const float sd = __expf(0.5f * lv);
                const float ev = a.eps ? (valid ? a.eps[((size_t)(t - a.t0) * a.B + rid) * m.out + o] : 0.f) : z[e];
                pred = fmaf(sd, ev, mean);
              }
              if (m.learned_rewards && o == m.out - 1) {
                rew_s[i] = pred;
              } else {
                my_obs[o] = c_nodelta[o] != 0.f ? pred : pred + my_obs[o];
              }
            }
          }
        }
        tc_fence_before();
        if (stamp) a.timeline[sp++] = clock64();  // outputs sampled
        epi_bar();
        if (stamp) a.timeline[sp++] = clock64();  // barrier
        // ---- reward, termination, accumulate: the row's owner thread ----
        if (owner) {
          float rew = m.learned_rewards ? rew_s[i] : reward_eval(m.reward_fn, my_act, m.A, 1, my_obs, m.D, 1);
          const bool done = term_eval(m.term_fn, my_obs, m.D, 1);
          if (valid) {
            if (a.reward_out) a.reward_out[rid] = rew;
            if (a.done_out) a.done_out[rid] = done ? 1 : 0;
          }
          if (dead) rew = 0.f;
          dead |= done ? 1 : 0;
          tot += rew;
        }
        if (stamp) a.timeline[sp++] = clock64();  // reward done
      }
      // ---- store row state ----
      if (owner && a.store_state && valid) {
        if (a.obs_out)
          for (int d = 0; d < m.D; ++d) a.obs_out[rid * m.D + d] = my_obs[d];
        if (a.total_state) a.total_state[rid] = tot;
        if (a.dead_state) a.dead_state[rid] = (uint8_t)dead;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------
// self test: one 128 x n x k GEMM through the same operand layouts / descriptors
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(int k, int n, const float* __restrict__ A,
                                                                const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* A_s = smem;                       // 128 * k * 2
  uint8_t* B_s = smem + 128 * k * 2;         // n * k * 2
  uint64_t* bar = reinterpret_cast<uint64_t*>(B_s + n * k * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  // A: [128][k] row-major fp32 -> canonical bf16
  for (int idx = tid; idx < 128 * (k / 8); idx += 128) {
    int i = idx % 128, kc = idx / 128;
    float x[8];
    for (int e = 0; e < 8; ++e) x[e] = A[i * k + kc * 8 + e];
    *reinterpret_cast<uint4*>(A_s + a_chunk_off(i, kc)) =
        make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
  }
  // B: [n][k] row-major fp32 -> [kc][n / 8][n % 8][8]
  for (int idx = tid; idx < n * (k / 8); idx += 128) {
    int nn = idx % n, kc = idx / n;
    float x[8];
    for (int e = 0; e < 8; ++e) x[e] = B[nn * k + kc * 8 + e];
    *reinterpret_cast<uint4*>(B_s + (size_t)(kc * (n / 8) + (nn >> 3)) * 128 + (nn & 7) * 16) =
        make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
  }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16_m128((uint32_t)n);
    const uint32_t b_lbo = (uint32_t)n * 16u;
    for (int kk = 0; kk < k / 16; ++kk) {
      const uint64_t adesc = umma_smem_desc(smem_u32(A_s) + (uint32_t)(2 * kk) * 2048u, 2048u, 128u);
      const uint64_t bdesc = umma_smem_desc(smem_u32(B_s) + (uint32_t)(2 * kk) * b_lbo, b_lbo, 128u);
      umma_bf16_ss(tmem_base, adesc, bdesc, idesc, kk != 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  __syncwarp();
  mbar_wait(bar, 0);
  tc_fence_after();
  const int i = warp * 32 + lane;
  for (int c = 0; c < n / 16; ++c) {
    uint32_t r[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(16 * c), r);
    tmem_ld_wait();
    for (int e = 0; e < 16; ++e) D[i * n + 16 * c + e] = __uint_as_float(r[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}


// self test of the A-from-TMEM form: the four warps write their rows of A (bf16 pairs) with tcgen05.st
__global__ void __launch_bounds__(128, 1) umma_selftest_ts_kernel(int k, int n, const float* __restrict__ A,
                                                                   const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* B_s = smem;
  uint64_t* bar = reinterpret_cast<uint64_t*>(B_s + n * k * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  for (int idx = tid; idx < n * (k / 8); idx += 128) {
    int nn = idx % n, kc = idx / n;
    float x[8];
    for (int e = 0; e < 8; ++e) x[e] = B[nn * k + kc * 8 + e];
    *reinterpret_cast<uint4*>(B_s + (size_t)(kc * (n / 8) + (nn >> 3)) * 128 + (nn & 7) * 16) =
        make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
  }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t a_col = 256;
  const int i = warp * 32 + lane;
  const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int kk = 0; kk < k / 16; ++kk) {
    uint32_t r[8];
    for (int e = 0; e < 8; ++e) r[e] = pack_bf16(A[i * k + kk * 16 + 2 * e], A[i * k + kk * 16 + 2 * e + 1]);
    tmem_st8(t_lane + a_col + (uint32_t)(8 * kk), r);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16_m128((uint32_t)n);
    const uint32_t b_lbo = (uint32_t)n * 16u;
    for (int kk = 0; kk < k / 16; ++kk) {
      const uint64_t bdesc = umma_smem_desc(smem_u32(B_s) + (uint32_t)(2 * kk) * b_lbo, b_lbo, 128u);
      umma_bf16_ts(tmem_base, tmem_base + a_col + (uint32_t)(8 * kk), bdesc, idesc, kk != 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  __syncwarp();
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c = 0; c < n / 16; ++c) {
    uint32_t r[16];
    tmem_ld16(t_lane + (uint32_t)(16 * c), r);
    tmem_ld_wait();
    for (int e = 0; e < 16; ++e) D[i * n + 16 * c + e] = __uint_as_float(r[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------
// micro-benchmark: issue rate of back-to-back tcgen05.mma for different shared-memory operand layouts
//   mode 0: no swizzle, K core columns strided (LBO = rows*16, SBO = 128)      <- layout of the rollout kernel
//   mode 1: SWIZZLE_128B K-major (SBO = 1024, 32 B start-address advance per K step inside a 64-element block)
//   mode 2: no swizzle, the two K halves of a row group adjacent (LBO = 128, SBO = (K/8)*128)
// Data is whatever is in shared memory: timing only.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;                      // LBO (unused for swizzled K-major)
  d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;  // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(128, 1) umma_bench_kernel(int mode, int k, int n, int reps, long long* out) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* A_s = smem;
  uint8_t* B_s = smem + 128 * 256 * 2;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 256 * 2 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16_m128((uint32_t)n);
    const uint32_t a0 = smem_u32(A_s), b0 = smem_u32(B_s);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int kk = 0; kk < (mode == 3 ? 0 : k / 16); ++kk) {
        uint64_t ad, bd;
        if (mode == 0) {
          ad = umma_smem_desc(a0 + (uint32_t)(2 * kk) * 2048u, 2048u, 128u);
          bd = umma_smem_desc(b0 + (uint32_t)(2 * kk) * (uint32_t)n * 16u, (uint32_t)n * 16u, 128u);
        } else if (mode == 1) {
          ad = umma_smem_desc_sw128(a0 + (uint32_t)(kk >> 2) * 128u * 128u + (uint32_t)(kk & 3) * 32u);
          bd = umma_smem_desc_sw128(b0 + (uint32_t)(kk >> 2) * (uint32_t)n * 128u + (uint32_t)(kk & 3) * 32u);
        } else {
          ad = umma_smem_desc(a0 + (uint32_t)kk * 256u, 128u, (uint32_t)(k / 8) * 128u);
          bd = umma_smem_desc(b0 + (uint32_t)kk * 256u, 128u, (uint32_t)(k / 8) * 128u);
        }
        if (mode != 3) umma_bf16_ss(tmem_base, ad, bd, idesc, 1u);
      }
      if (mode == 3) {  // tight issue: precomputed descriptors, start-address field advanced by a constant
        uint64_t ad = umma_smem_desc(a0, 2048u, 128u), bd = umma_smem_desc(b0, (uint32_t)n * 16u, 128u);
        const uint64_t a_inc = (2u * 2048u) >> 4, b_inc = (2u * (uint32_t)n * 16u) >> 4;
        const int nk = k / 16;
#pragma unroll 4
        for (int kk = 0; kk < nk; ++kk) {
          umma_bf16_ss(tmem_base, ad, bd, idesc, 1u);
          ad += a_inc;
          bd += b_inc;
        }
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static int g_sm_count = 0, g_max_smem = 0;

static int tc_device_limits() {
  if (g_sm_count) return B200PETS_OK;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
  CUDA_TRY(cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return B200PETS_OK;
}

// smem plan for a model; returns false when the tensor-core path does not cover the dimensions
bool tc_make_plan(const ModelDev& m, int max_smem, TcPlan* out) {
  TcPlan p{};
  p.nlayers = m.L + 1;
  if (m.in + 2 > 256 || m.hid + 2 > 256 || m.nout > 256 || m.D > 256) return false;
  int kp_max = 0, tm = 32;
  uint32_t stage = 0;
  for (int l = 0; l < p.nlayers; ++l) {
    if (m.Np[l] > 256 || m.Kp[l] > 256) return false;
    p.nblk[l] = (m.Kp[l] + 63) / 64;
    kp_max = max(kp_max, m.Kp[l]);
    stage = max(stage, (uint32_t)min(64, m.Kp[l]) * m.Np[l] * 2u);
    while (tm < m.Np[l]) tm *= 2;
  }
  p.stage_bytes = (stage + 127u) & ~127u;
  p.tmem_cols = (uint32_t)tm;
  p.obs_ld = m.D | 1;
  p.act_ld = m.A | 1;
  uint32_t off = 0;
  p.off_A = off; off += (uint32_t)kTileM * kp_max * 2;
  p.off_obs = off; off += (uint32_t)kTileM * p.obs_ld * 4;
  p.off_act = off; off += (uint32_t)kTileM * p.act_ld * 4;
  p.off_const = off; off += (uint32_t)(2 * m.in + 2 * m.out + m.D + kTileM) * 4;
  off = (off + 15u) & ~15u;
  p.off_bar = off; off += (2 * kMaxStages + 2) * 8 + 16;
  off = (off + 127u) & ~127u;
  p.off_ring = off;
  int S = ((int)max_smem - (int)off) / (int)p.stage_bytes;
  if (S < 2) return false;
  p.nstages = min(S, kMaxStages);
  p.smem_bytes = off + (uint32_t)p.nstages * p.stage_bytes;
  *out = p;
  return true;
}

bool tc_supported(const ModelDev& m) {
  if (tc_device_limits() != B200PETS_OK) return false;
  TcPlan p;
  return tc_make_plan(m, g_max_smem, &p);
}

int launch_rollout_tc(const ModelDev& m, const RolloutArgs& a, cudaStream_t stream) {
  int rc = tc_device_limits();
  if (rc) return rc;
  if (a.propagation == B200PETS_PROP_EXPECTATION)
    return b200pets_set_error(B200PETS_EUNSUPPORTED, "tensor-core path does not cover propagation='expectation'");
  TcPlan p;
  if (!tc_make_plan(m, g_max_smem, &p))
    return b200pets_set_error(B200PETS_EUNSUPPORTED, "model dimensions outside the tensor-core path (in %d hid %d out %d)",
                              m.in, m.hid, m.out);
  long long tiles;
  if (a.slot_mode >= 1) {
    tiles = (a.B + kTileM - 1) / kTileM;
  } else {
    long long Bm = a.B / m.M;
    tiles = (long long)m.M * ((Bm + kTileM - 1) / kTileM);
  }
  const unsigned grid = (unsigned)min((long long)g_sm_count, tiles);
  void (*kern)(const ModelDev, const RolloutArgs, const TcPlan, const long long) = nullptr;
  switch (m.act) {
    case B200PETS_ACT_SILU: kern = rollout_tc_kernel<B200PETS_ACT_SILU, kEpiSplit>; break;
    case B200PETS_ACT_RELU: kern = rollout_tc_kernel<B200PETS_ACT_RELU, kEpiSplit>; break;
    default: kern = rollout_tc_kernel<B200PETS_ACT_LEAKY_RELU, kEpiSplit>; break;
  }
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
  kern<<<grid, 64 + 128 * kEpiSplit, p.smem_bytes, stream>>>(m, a, p, tiles);
  CUDA_TRY(cudaGetLastError());
  return B200PETS_OK;
}

int launch_umma_selftest(int k, int n, const float* a, const float* b, float* d, cudaStream_t stream) {
  const bool ts = k < 0;  // negative k selects the A-from-TMEM form
  if (ts) k = -k;
  if (k % 16 || n % 16 || n > 256 || k > 256 || k < 16 || n < 16)
    return b200pets_set_error(B200PETS_EINVAL, "selftest needs k, n multiples of 16, <= 256");
  if (ts) {
    size_t smem_ts = (size_t)n * k * 2 + 64;
    CUDA_TRY(cudaFuncSetAttribute(umma_selftest_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ts));
    umma_selftest_ts_kernel<<<1, 128, smem_ts, stream>>>(k, n, a, b, d);
    CUDA_TRY(cudaGetLastError());
    return B200PETS_OK;
  }
  size_t smem = (size_t)128 * k * 2 + (size_t)n * k * 2 + 64;
  CUDA_TRY(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, stream>>>(k, n, a, b, d);
  CUDA_TRY(cudaGetLastError());
  return B200PETS_OK;
}

int launch_umma_bench(int mode, int k, int n, int reps, long long* out, cudaStream_t stream) {
  if (k % 16 || n % 16 || n > 256 || k > 256) return b200pets_set_error(B200PETS_EINVAL, "umma_bench: bad shape");
  size_t smem = (size_t)(128 + 256) * 256 * 2 + 1024;
  CUDA_TRY(cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_bench_kernel<<<1, 128, smem, stream>>>(mode, k, n, reps, out);
  CUDA_TRY(cudaGetLastError());
  return B200PETS_OK;
}
