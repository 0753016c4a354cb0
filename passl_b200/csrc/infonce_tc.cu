// Fused InfoNCE forward on tcgen05 — the headline kernel (BASELINE.json: "fused InfoNCE HBM GB/s vs roofline").
//
//   logits[i, j] = scale * <Q_i, K_j>  are produced tile by tile in TMEM and consumed in place by an online
//   softmax (running row max / sum) — the [N, K] logit matrix never exists in HBM.  HBM traffic is the
//   algorithmic minimum: the key matrix (MoCo queue) is streamed exactly once by TMA, Q is read once per CTA
//   (L2 resident).   SURVEY.md §8(d):  bytes = (2*N*D + D*K)*2 + 4*N,  C3: 16.91 MB, 4.295 GFLOP.
//
// Replaces: paddle.matmul(q, queue) + concat + /T + CrossEntropyLoss + topk
//           (passl_v110/modeling/architectures/moco.py:178-182, heads/contrastive_head.py:37-60),
//           einsum('nc,mc->nm')/T + CE (passl/models/mocov3.py:187-198), CLIP logits + CE (clip.py:331-335).
//
// Work decomposition: CTA = (row group of MB*128 queries) x (contiguous slice of 64-key tiles).
//   warp 0 lane 0 : TMA producer  (key tiles through a STAGES ring, SWIZZLE_128B)
//   warp 1 lane 0 : MMA issuer    (tcgen05.mma M=128 N=64 K=16, A = Q resident in TMEM (written once by the softmax warps
//                                  with tcgen05.st) so only the key tile is read from shared memory; S double-buffered)
//   warps 2..2+4*MB : softmax warps — tcgen05.ld 64 logits / thread / tile, online max+sum in the log2 domain,
//                     rank counter for top-1 / top-5 accuracy.
// Partials (m, l, cnt) per (row, slice) are merged by simce_finalize_kernel (shared with the fp32 variant).
#include "common.cuh"
#include "host_utils.h"
#include "simce_common.cuh"
#include "../../include/passl_b200.h"

#include <string.h>

namespace pb {

constexpr int NCE_BK = 64;          // keys per tile (= MMA N)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct InfoNceTcParams {
  CUtensorMap q_map;  // [N, D] bf16, box {64, 128}
  CUtensorMap k_map;  // [K, D] bf16, box {64, 64}
  const __nv_bfloat16* Q;
  const __nv_bfloat16* Kmat;
  const float* P;           // [N, D] fp32 positive keys (extra column) or null
  const long long* label;   // [N] or null
  const int* excl;          // [N] or null
  float scale;
  int N, K, D;
  int row_groups, slices, tiles;
  float* part_m; float* part_l; int* part_cnt;  // [N, slices]
  float* tgt;                                   // [N]
  const float* tgt_raw;                         // [N] raw target dot products from infonce_target_kernel (PDL producer)
  unsigned long long* dbg;                      // optional per-CTA timeline (globaltimer ns), 16 slots per CTA
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define NCE_STAMP(slot) do { if (p.dbg) p.dbg[blockIdx.x * 16 + (slot)] = gtime(); } while (0)

// Target logits <q_i, k+_i> (MoCo: P = positive keys, fp32) or <q_i, K[label_i]> (MoCo v3 / CLIP), one warp per row with whole-row
// coalesced loads.  Runs as the programmatic-dependent-launch PRODUCER of the main kernel: it releases its dependents at once,
// so the main kernel's setup (barriers, TMEM, Q staging, first key tiles) overlaps it; the main kernel waits
// (griddepcontrol.wait) just before it needs the 1 KB of results.  Previously every CTA recomputed all its rows' targets
// (148 x 128 KB of L2 reads, ~8 us of latency-bound prologue).
__global__ void __launch_bounds__(256) infonce_target_kernel(const __nv_bfloat16* __restrict__ Q, const __nv_bfloat16* __restrict__ Kmat,
                                                             const float* __restrict__ P, const long long* __restrict__ label,
                                                             float* __restrict__ tgt_raw, unsigned* __restrict__ ticket, int N, int K,
                                                             int D) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (blockIdx.x == 0 && threadIdx.x == 0) *ticket = 0u;   // finalize's last-block ticket (it runs two grids later): no memset node
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  float acc = 0.f;
  long long lab = 0;
  if (!P) {
    lab = label[row];
    lab = lab < 0 ? 0 : (lab >= K ? K - 1 : lab);
  }
  for (int d4 = lane * 4; d4 < D; d4 += 128) {
    const uint2 qu = *reinterpret_cast<const uint2*>(Q + (size_t)row * D + d4);
    const float2 q0 = unpack_bf16x2(qu.x), q1 = unpack_bf16x2(qu.y);
    if (P) {
      const float4 pa = *reinterpret_cast<const float4*>(P + (size_t)row * D + d4);
      acc += q0.x * pa.x + q0.y * pa.y + q1.x * pa.z + q1.y * pa.w;
    } else {
      const uint2 ku = *reinterpret_cast<const uint2*>(Kmat + (size_t)lab * D + d4);
      const float2 k0 = unpack_bf16x2(ku.x), k1 = unpack_bf16x2(ku.y);
      acc += q0.x * k0.x + q0.y * k0.y + q1.x * k1.x + q1.y * k1.y;
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) tgt_raw[row] = acc;
}

template <int MB>
__global__ void __launch_bounds__(64 + 128 * MB, 1) infonce_tc_fwd_kernel(const __grid_constant__ InfoNceTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int DC = p.D / 64;                       // 64-wide feature chunks
  const int q_bytes = MB * 128 * p.D * 2;
  const int stage_bytes = NCE_BK * p.D * 2;
  int STAGES = (200 * 1024 - q_bytes) / stage_bytes;   // same rule as nce_plan() on the host
  if (STAGES > 8) STAGES = 8;
  uint8_t* q_smem = smem;                              // staging only: rows go smem -> registers -> TMEM
  uint8_t* k_smem = smem + q_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(k_smem + STAGES * stage_bytes);
  uint64_t* q_ready = bars;           // 4*MB softmax-warp arrivals: Q rows are in TMEM
  uint64_t* k_full = bars + 1;        // STAGES
  uint64_t* k_empty = k_full + 8;     // STAGES
  uint64_t* s_full = k_empty + 8;     // 2
  uint64_t* s_empty = s_full + 2;     // 2
  uint64_t* q_full = s_empty + 2;     // TMA: Q block staged in shared memory
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(q_full + 1);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the finalize grid may be scheduled early (it waits for us)
  const uint32_t warp = warp_id(), lane = lane_id();
  const int group = blockIdx.x / p.slices;
  const int slice = blockIdx.x - group * p.slices;
  const int row_base = group * MB * 128;
  const int t_begin = (int)((long long)slice * p.tiles / p.slices);
  const int t_end = (int)((long long)(slice + 1) * p.tiles / p.slices);
  constexpr uint32_t TMEM_COLS = 512;
  const uint32_t q_cols = p.D / 2;                 // bf16x2 per 32-bit TMEM column

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.q_map);
    tma_prefetch_desc(&p.k_map);
    mbar_init(q_ready, 4 * MB);
    mbar_init(q_full, 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4 * MB);
    }
    fence_barrier_init();
  }
  if (threadIdx.x == 64) NCE_STAMP(0);
  if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 64) NCE_STAMP(1);
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tm_q = tmem_base;                       // Q (A operand): block b at columns [b*q_cols, (b+1)*q_cols)
  const uint32_t tm_s = tmem_base + MB * q_cols;         // S accumulators: (buf*MB + b) * 64

  // role loops are warp-uniform; only the TMA / tcgen05 issue is elect-predicated (keeps descriptors in uniform registers)
  if (warp == 0) {
    {
      // ---------------- TMA producer ----------------
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, (uint32_t)q_bytes);
        for (int b = 0; b < MB; ++b)
          for (int c = 0; c < DC; ++c)
            tma_load_2d(q_smem + (b * DC + c) * (128 * 128), &p.q_map, q_full, c * 64, row_base + b * 128);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&k_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[stage], (uint32_t)stage_bytes);
          for (int c = 0; c < DC; ++c)
            tma_load_2d(k_smem + stage * stage_bytes + c * (NCE_BK * 128), &p.k_map, &k_full[stage], c * 64, t * NCE_BK);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc = make_idesc_bf16(128, NCE_BK, false, false);
      const uint64_t db0 = make_smem_desc_sw128(smem_u32(k_smem), 16, 1024);   // stage 0, chunk 0, k-step 0
      mbar_wait(q_ready, 0);
      tc_fence_after();
      if (lane == 0) NCE_STAMP(14);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        mbar_wait(&s_empty[buf], bphase ^ 1);
        mbar_wait(&k_full[stage], phase);
        tc_fence_after();
        const uint64_t dbs = db0 + (uint64_t)((stage * stage_bytes) >> 4);
        if (elect_one()) {
        for (int b = 0; b < MB; ++b) {
          const uint32_t d_tmem = tm_s + (buf * MB + b) * NCE_BK;
          const uint32_t a_tmem = tm_q + b * q_cols;
          for (int c = 0; c < DC; ++c) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ts(d_tmem, a_tmem + c * 32 + k * 8, dbs + (uint64_t)((c * (NCE_BK * 128) + k * 32) >> 4), idesc,
                           (c > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&k_empty[stage]);
        umma_commit(&s_full[buf]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ---------------- softmax warps ----------------
    const int ew = warp - 2;            // 0 .. 4*MB-1
    const int b = ew >> 2;              // row block
    const uint32_t q4 = warp & 3;       // TMEM lane quarter
    const int row = row_base + b * 128 + q4 * 32 + lane;
    const bool row_ok = row < p.N;
    const float c2 = p.scale * kLog2e;  // logits in the log2 domain: y = dot * c2

    // Prologue: Q block arrives by TMA in shared memory; each thread copies its own row smem -> registers -> TMEM (A operand
    // of every MMA); the target logit of the row (positive pair or labelled column) is read from the producer kernel's output.
    float tgt2 = 0.f, tgt_raw = 0.f;
    long long lab = -1;
    int ex = -1;
    {
      const int rl = q4 * 32 + lane;                       // row inside the 128-row block
      mbar_wait(q_full, 0);
      if (threadIdx.x == 64) NCE_STAMP(2);
      const uint32_t tq = tm_q + ((q4 * 32u) << 16) + b * q_cols;
      for (int ch = 0; ch < DC; ++ch) {
        const uint8_t* base = q_smem + (b * DC + ch) * (128 * 128) + (rl >> 3) * 1024 + (rl & 7) * 128;
        uint32_t w[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint4 u = *reinterpret_cast<const uint4*>(base + (((g ^ (rl & 7)) & 7) << 4));
          w[g * 4 + 0] = u.x; w[g * 4 + 1] = u.y; w[g * 4 + 2] = u.z; w[g * 4 + 3] = u.w;
        }
        tmem_st_32x32(tq + ch * 32, w);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready);
      if (threadIdx.x == 64) NCE_STAMP(3);
      if (row_ok) {
        if (!p.P) lab = p.label[row];
        if (p.excl) ex = p.excl[row];
      }
      // targets come from infonce_target_kernel (PDL producer): block here, as late as possible, until that grid has completed
      asm volatile("griddepcontrol.wait;" ::: "memory");
      const float s_mine = row_ok ? __ldcg(p.tgt_raw + row) : 0.f;
      tgt2 = s_mine * c2;
      tgt_raw = s_mine;
      if (threadIdx.x == 64) NCE_STAMP(4);
    }

    float m = -INFINITY, l = 0.f;
    int cnt = 0;
    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      mbar_wait(&s_full[buf], bphase);
      tc_fence_after();
      if (threadIdx.x == 64 && it < 8) NCE_STAMP(5 + it);
      const uint32_t taddr = tm_s + ((q4 * 32u) << 16) + (buf * MB + b) * NCE_BK;
      uint32_t v[64];
      tmem_ld_32x32(taddr, v);
      tmem_ld_32x32(taddr + 32, v + 32);
      tmem_ld_wait();
      // TMEM buffer can be refilled as soon as the values are in registers
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[buf]);

      const int key0 = t * NCE_BK;
      const bool special = (key0 + NCE_BK > p.K) || (ex >= key0 && ex < key0 + NCE_BK) ||
                           (lab >= key0 && lab < key0 + NCE_BK);
      if (!special) {
        // fast path (instruction diet): work on the raw dot products — max is monotone in the positive scale, the scaling
        // and the max subtraction fold into one FFMA in front of ex2, the rank counter uses set.gt + FADD on 4 chains.
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
        float c0 = 0.f, c1 = 0.f, c2n = 0.f, c3 = 0.f;
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          const float a0 = __uint_as_float(v[j]), a1 = __uint_as_float(v[j + 1]), a2 = __uint_as_float(v[j + 2]),
                      a3 = __uint_as_float(v[j + 3]);
          mx0 = fmaxf(mx0, a0); mx1 = fmaxf(mx1, a1); mx2 = fmaxf(mx2, a2); mx3 = fmaxf(mx3, a3);
          float g0, g1, g2, g3;
          asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(g0) : "f"(a0), "f"(tgt_raw));
          asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(g1) : "f"(a1), "f"(tgt_raw));
          asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(g2) : "f"(a2), "f"(tgt_raw));
          asm("set.gt.f32.f32 %0, %1, %2;" : "=f"(g3) : "f"(a3), "f"(tgt_raw));
          c0 += g0; c1 += g1; c2n += g2; c3 += g3;
        }
        cnt += (int)((c0 + c1) + (c2n + c3));
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c2;
        const float mn = fmaxf(m, mx);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          s0 += exp2f(fmaf(__uint_as_float(v[j]), c2, -mn));
          s1 += exp2f(fmaf(__uint_as_float(v[j + 1]), c2, -mn));
          s2 += exp2f(fmaf(__uint_as_float(v[j + 2]), c2, -mn));
          s3 += exp2f(fmaf(__uint_as_float(v[j + 3]), c2, -mn));
        }
        l = l * exp2f(m - mn) + ((s0 + s1) + (s2 + s3));
        m = mn;
      } else {
        float y[64];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const int key = key0 + j;
          const bool ok = (key < p.K) && (key != ex);
          y[j] = ok ? __uint_as_float(v[j]) * c2 : -INFINITY;
          mx = fmaxf(mx, y[j]);
          cnt += (ok && key != lab && y[j] > tgt2) ? 1 : 0;
        }
        if (mx > -INFINITY) {
          const float mn = fmaxf(m, mx);
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 64; ++j) s += exp2f(y[j] - mn);
          l = l * exp2f(m - mn) + s;
          m = mn;
        }
      }
    }
    if (threadIdx.x == 64) NCE_STAMP(13);
    if (row_ok) {
      // back to the natural-log domain used by simce_finalize_kernel
      p.part_m[(size_t)row * p.slices + slice] = m * kLn2;
      p.part_l[(size_t)row * p.slices + slice] = l;
      p.part_cnt[(size_t)row * p.slices + slice] = cnt;
      if (slice == 0) p.tgt[row] = tgt2 * kLn2;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static void nce_plan(int N, int K, int D, int& MB, int& groups, int& slices, int& tiles, int& smem) {
  MB = (D <= 128 && N > 128) ? 2 : 1;
  if (D > 256) MB = 1;
  groups = (N + MB * 128 - 1) / (MB * 128);
  tiles = (K + NCE_BK - 1) / NCE_BK;
  slices = num_sms() / groups;
  if (slices < 1) slices = 1;
  if (slices > tiles) slices = tiles;
  int q_bytes = MB * 128 * D * 2, stage_bytes = NCE_BK * D * 2;
  int stages = (200 * 1024 - q_bytes) / stage_bytes;
  if (stages > 8) stages = 8;
  smem = q_bytes + stages * stage_bytes + 512 + 1024;
}

}  // namespace pb

using namespace pb;

static unsigned long long* g_nce_dbg = nullptr;
// developer hook: per-CTA timeline buffer (uint64 [grid * 16]) filled by the next launches; NULL disables
extern "C" int passl_b200_infonce_tc_set_debug(void* buf) { g_nce_dbg = reinterpret_cast<unsigned long long*>(buf); return 0; }

extern "C" long long passl_b200_infonce_tc_workspace_bytes(int N, int K, int D) {
  int MB, groups, slices, tiles, smem;
  nce_plan(N, K, D, MB, groups, slices, tiles, smem);
  return (long long)N * slices * 12 + simce_finalize_scratch_bytes(N) + (long long)N * 4 + 512;
}

// Forward.  Q [N,D] bf16 (normalised queries), Kmat [K,D] bf16 keys, P [N,D] fp32 optional positive keys
// (MoCo), label int64 [N] (when P == NULL), excl int32 [N] optional.  Outputs as passl_b200_simce_fwd_f32.
extern "C" int passl_b200_infonce_tc_fwd(const void* Q, const void* Kmat, const float* P, const long long* label,
                                         const int* excl, float scale, float loss_scale, int N, int K, int D, float* lse,
                                         float* tgt, float* loss_rows, float* out_scalars, void* workspace,
                                         long long workspace_bytes, void* stream) {
  if (N <= 0 || K <= 0 || D < 64 || D % 64 || D > 512) return PB_ERR_BAD_ARG;
  if (!P && !label) return PB_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(Kmat) | reinterpret_cast<uintptr_t>(P)) & 15) return PB_ERR_BAD_ARG;
  if (workspace_bytes < passl_b200_infonce_tc_workspace_bytes(N, K, D)) return PB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  InfoNceTcParams p;
  memset(&p, 0, sizeof(p));
  int MB, smem;
  nce_plan(N, K, D, MB, p.row_groups, p.slices, p.tiles, smem);
  p.Q = reinterpret_cast<const __nv_bfloat16*>(Q);
  p.Kmat = reinterpret_cast<const __nv_bfloat16*>(Kmat);
  p.P = P; p.label = label; p.excl = excl; p.scale = scale; p.N = N; p.K = K; p.D = D;
  char* ws = reinterpret_cast<char*>(workspace);
  p.part_m = reinterpret_cast<float*>(ws); ws += (size_t)N * p.slices * 4;
  p.part_l = reinterpret_cast<float*>(ws); ws += (size_t)N * p.slices * 4;
  p.part_cnt = reinterpret_cast<int*>(ws);
  p.tgt = tgt;
  p.dbg = g_nce_dbg;
  uint64_t qd[2] = {(uint64_t)D, (uint64_t)N}, qs[1] = {(uint64_t)D * 2};
  uint32_t qbx[2] = {64, 128};
  int rc = make_tmap_bf16(&p.q_map, Q, 2, qd, qs, qbx);
  if (rc) return rc;
  uint64_t kd[2] = {(uint64_t)D, (uint64_t)K};
  uint32_t kbx[2] = {64, NCE_BK};
  rc = make_tmap_bf16(&p.k_map, Kmat, 2, kd, qs, kbx);
  if (rc) return rc;
  int grid = p.row_groups * p.slices;
  static bool attr_done = false;
  if (!attr_done) {  // once per process: allow up to the full 227 KB dynamic smem carve-out
    PB_CUDA_CHECK(cudaFuncSetAttribute(infonce_tc_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PB_CUDA_CHECK(cudaFuncSetAttribute(infonce_tc_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  // three launches chained by programmatic dependent launch: target (also zeroes the ticket) -> main (PDL) -> finalize (PDL)
  float* scratch = reinterpret_cast<float*>(p.part_cnt + (size_t)N * p.slices);
  const int fin_blk = (N + 7) / 8;
  float* tgt_raw = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + ((simce_finalize_scratch_bytes(N) + 15) & ~15LL));
  p.tgt_raw = tgt_raw;
  infonce_target_kernel<<<fin_blk, 256, 0, st>>>(p.Q, p.Kmat, P, label, tgt_raw, reinterpret_cast<unsigned*>(scratch + (size_t)fin_blk * 3),
                                                 N, K, D);
  PB_LAUNCH_CHECK();
  {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + 128 * MB); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (MB == 2) PB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, infonce_tc_fwd_kernel<2>, p));
    else PB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, infonce_tc_fwd_kernel<1>, p));
    PB_LAUNCH_CHECK();
  }
  PB_CUDA_CHECK(launch_simce_finalize(p.part_m, p.part_l, p.part_cnt, tgt, N, p.slices, P ? 1 : 0, loss_scale, lse,
                                      loss_rows, out_scalars, scratch, st, /*pdl=*/true));
  passl_b200_launch_counter_add(1);
  return PB_OK;
}
