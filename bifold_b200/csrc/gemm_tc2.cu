// tcgen05 GEMM, CTA-pair variant (cta_group::2): a cluster of two CTAs computes one 256 x 256 output tile.
//
//   * CTA r of the pair owns rows [128 r, 128 r + 128) of the tile (its TMEM holds them, all 256 columns) and
//     stages only HALF of the B tile (128 of the 256 n-rows): 32 KB per 64-wide K block instead of 48 KB, i.e.
//     1.5x the arithmetic intensity per byte pulled from L2 and 6 pipeline stages instead of 4.
//   * both CTAs issue their own TMA loads (cta_group::2 form, completing on the LEADER's mbarrier);
//     only the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16);
//     tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both CTAs;
//     both CTAs' epilogue warps release the accumulator on the leader's barrier (remote arrive).
//   * same fused epilogue as gemm_tc.cu (tc_epilogue.cuh).
// Opt-in from gemm_tc.cu (BFK_GEMM_2CTA).
#include "tc_epilogue.cuh"

namespace bfk {

int gemm_dbg();   // gemm_tc.cu: BFK_GEMM_DBG measurement knobs (4 = no stores, 8 = no TMEM read, 16 = no epilogue body)
bool make_map(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long ld, int box_inner, int box_outer, int swizzle_bytes = 128,
              int elem_bytes = 2);

#ifndef BFK_GG_DOUBLE_STAGE
#define BFK_GG_DOUBLE_STAGE 1      // A/B knob: 0 = 6 operand stages and 2 KB of store staging per warp for the value + derivative modes too
#endif
// Operand ring depth / store staging per epilogue warp.  The dual-output (value + derivative) modes trade one of the six operand
// stages for 4 KB of staging per warp, so a half chunk is staged while the copy engine still reads the previous one.
__host__ __device__ constexpr bool t2_is_gg(int mode) { return BFK_GG_DOUBLE_STAGE && (mode == EPI_GELUG_TANH_BF16 || mode == EPI_GELUG_ERF_BF16); }
__host__ __device__ constexpr int t2_stages(int mode) { return t2_is_gg(mode) ? 5 : 6; }
__host__ __device__ constexpr int t2_stage_bytes_per_warp(int mode) { return t2_is_gg(mode) ? 4096 : 2048; }
constexpr int T2_A_BYTES = 128 * 64 * 2, T2_B_BYTES = 128 * 64 * 2;
constexpr int T2_STAGE_BYTES = T2_A_BYTES + T2_B_BYTES;
constexpr int T2_BN = 256;
constexpr int T2_TLBAR_OFF = 256 + 2 * T2_BN * 4;                 // barriers + TMEM slot | bias [2][256] | per-warp operand barriers
constexpr int T2_AUX_BYTES = 3072;                                // rounded up: the staging area behind it is 1024-byte aligned (SWIZZLE_128B boxes)
constexpr int T2_MAX_EPI_WARPS = 16;
// operand ring | barriers, bias | TMA-store staging (2 KB per epilogue warp); the dynamic shared memory is declared 1024-byte aligned
__host__ __device__ constexpr int t2_smem_bytes(int mode) {
  return t2_stages(mode) * T2_STAGE_BYTES + T2_AUX_BYTES + T2_MAX_EPI_WARPS * t2_stage_bytes_per_warp(mode);
}
static_assert(t2_smem_bytes(EPI_GENERIC) <= 232448 && t2_smem_bytes(EPI_GELUG_TANH_BF16) <= 232448, "CTA-pair GEMM: shared memory budget");

__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// the data lands in THIS CTA's smem, the transaction bytes complete on the barrier at `cluster_bar` (the leader's)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}

// Work item -> (row-pair block, n block, MMA N of the tile).  When the last n block holds at most 128 live columns
// (N = 784: 16) it becomes a THIN tile: the UMMA runs with N = live columns rounded up to 16 instead of 256, and all thin
// tiles are enumerated after the full ones so the static round-robin over clusters stays balanced.
struct T2Item { int mp, nb, ntile; };
__device__ __forceinline__ T2Item t2_decode(int item, int num_mp, int num_n, int N) {
  const int tail_w = N - (num_n - 1) * T2_BN;
  const bool thin = tail_w <= 128 && num_n > 1;
  const int nf = thin ? num_n - 1 : num_n;
  const int full_items = num_mp * nf;
  T2Item it;
  if (item < full_items) { it.mp = item / nf; it.nb = item % nf; it.ntile = T2_BN; }
  else { it.mp = item - full_items; it.nb = num_n - 1; it.ntile = (tail_w + 15) & ~15; }
  return it;
}

// MODE: compile-time EpiMode.  EPI_GENERIC keeps every fused variant behind runtime flags (8 epilogue warps, 168
// registers); a specialised MODE compiles only its own epilogue and runs 16 epilogue warps (launch bound 576 threads).
template <bool AMN, bool BMN, int MODE>
__global__ void __launch_bounds__(64 + 32 * epi_warps(MODE), 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_d,
                const __grid_constant__ CUtensorMap tma_x, int M, int N, int K, EpiParams ep) {
  constexpr int EW = epi_warps(MODE);
  constexpr int T2_STAGES = t2_stages(MODE);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t bar_base = smem_base + T2_STAGES * T2_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (T2_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * T2_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * T2_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * T2_STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* bias_smem = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));   // [2][256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int num_mp = (M + 255) / 256, num_n = (N + T2_BN - 1) / T2_BN;
  const int items = num_mp * num_n;
  const int num_kb = (K + 63) / 64;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
    if (epi_mode_can_tma_load(MODE) && ep.tma_load) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_x)) : "memory");
    if (epi_mode_can_tma_store(MODE) && ep.tma_store) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_d)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_x)) : "memory");
    }
    for (int s = 0; s < T2_STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * EW); }
    if (epi_mode_can_tma_load(MODE))
      for (int w = 0; w < EW; ++w) mbar_init(bar_base + T2_TLBAR_OFF + 8u * w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();     // a dependent kernel may be scheduled from here on (it blocks in its own pdl_wait until this grid is done)
  pdl_wait();        // PDL launch: everything above overlapped the predecessor's tail; its results may be read from here on

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer (both CTAs): own 128 rows of A, own 128 n-rows of B =====
      int stage = 0; uint32_t phase = 0;
      for (int item = cid; item < items; item += ncl) {
        const T2Item it = t2_decode(item, num_mp, num_n, N);
        const int m0 = it.mp * 256 + crank * 128, nb0 = it.nb * T2_BN + crank * (it.ntile >> 1);   // this CTA's half of the B rows
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t lead_full = map_to_cta(full_bar(stage), 0);
          mbar_expect_tx_cluster(lead_full, T2_STAGE_BYTES);
          const uint32_t sa = smem_base + stage * T2_STAGE_BYTES, sb = sa + T2_A_BYTES;
          const int k0 = kb * 64;
          if constexpr (!AMN) {
            tma_load_2d_2sm(sa, &tma_a, lead_full, k0, m0);
          } else {
            tma_load_2d_2sm(sa, &tma_a, lead_full, m0, k0);
            tma_load_2d_2sm(sa + 8192, &tma_a, lead_full, m0 + 64, k0);
          }
          if constexpr (!BMN) {
            tma_load_2d_2sm(sb, &tma_b, lead_full, k0, nb0);
          } else {
            tma_load_2d_2sm(sb, &tma_b, lead_full, nb0, k0);
            tma_load_2d_2sm(sb + 8192, &tma_b, lead_full, nb0 + 64, k0);
          }
          if (++stage == T2_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0 && elect_one()) {
      // ===== MMA issuer (leader CTA only): UMMA 256 x 256 x 16 over both CTAs' smem / TMEM =====
      constexpr uint32_t idesc_full = make_idesc(256, T2_BN, AMN, BMN);
      constexpr int KSTEP_A = (AMN ? 2048 : 32) >> 4, KSTEP_B = (BMN ? 2048 : 32) >> 4;   // descriptor step per UMMA_K slice
      const uint64_t da0 = operand_desc<AMN>(smem_base, 0), db0 = operand_desc<BMN>(smem_base + T2_A_BYTES, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int item = cid; item < items; item += ncl) {
        const T2Item it = t2_decode(item, num_mp, num_n, N);
        const uint32_t idesc = it.ntile == T2_BN ? idesc_full : make_idesc(256, it.ntile, AMN, BMN);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * T2_BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t soff = (uint64_t)(stage * (T2_STAGE_BYTES >> 4));     // descriptors are built once (da0, db0)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16_2sm(tmem_d, da0 + soff + (uint64_t)(k * KSTEP_A), db0 + soff + (uint64_t)(k * KSTEP_B), idesc,
                            (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit_2sm_mc(empty_bar(stage), (uint16_t)0x3);
          if (++stage == T2_STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit_2sm_mc(tfull_bar(acc), (uint16_t)0x3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===== epilogue warps (both CTAs): own 128 rows x 256 columns; quarter = TMEM lane quarter, sub = warp within it =====
    const int quarter = warp & 3, sub = (warp - 2) >> 2;
    const int half = sub;   // generic path: two warps per quarter
    const int et = threadIdx.x - 64;
    int acc = 0; uint32_t acc_phase = 0;
    TlState tl{0u, false};
    const uint32_t tl_bar = bar_base + T2_TLBAR_OFF + 8u * (uint32_t)(warp - 2);
    const uint32_t tl_stage = bar_base + T2_AUX_BYTES + (uint32_t)((warp - 2) * (MODE == EPI_RES_F32 ? 4096 : 2048));
    for (int item = cid; item < items; item += ncl) {
      const T2Item it = t2_decode(item, num_mp, num_n, N);
      const int m0 = it.mp * 256 + crank * 128, n0 = it.nb * T2_BN;
      float* bias_s = bias_smem + acc * T2_BN;
      for (int i = et; i < T2_BN; i += 32 * EW) bias_s[i] = (ep.bias && n0 + i < N) ? __ldg(ep.bias + n0 + i) : 0.f;
      named_bar_sync(1, 32 * EW);
      // (no L2 prefetch of the next tile's epilogue operands: one prefetch per thread and ROW costs 32 LSU wavefronts per warp
      //  instruction like the loads themselves; measured 2 us per launch slower on every variant that reads operands)
      const int m = m0 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * T2_BN);
      const bool skip = (ep.vec_ok & 16) != 0;   // measurement only
      if constexpr (MODE != EPI_GENERIC) {
        EpiOps<MODE> ops;
        const bool use_tl = epi_mode_can_tma_load(MODE) && ep.tma_load;
        if (use_tl) {
          // this tile's first box must be under way before the wait for the accumulator (normally requested one tile ago)
          if (n0 + sub * 32 < N && !tl.pending) {
            if (lane == 0) tl_request<MODE>(tl_stage, tl_bar, &tma_x, n0 + sub * 32, m0 + quarter * 32);
            tl.pending = true;
          }
        } else {
          epi_tile_preload<MODE>(ep, m, n0, sub, ops);         // first chunk's operands: in flight during the wait
        }
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        if (!skip) {
          if (use_tl) {
            const bool has_next = item + ncl < items;
            const T2Item nit = t2_decode(has_next ? item + ncl : item, num_mp, num_n, N);
            epi_tile_tl<MODE, (T2_BN / 32) / (EW / 4), EW / 4>(ep, m, m0 + quarter * 32, n0, sub, taddr, bias_s, tl_stage, tl_bar, tl, &tma_x, lane,
                                                              has_next, nit.mp * 256 + crank * 128 + quarter * 32, nit.nb * T2_BN);
          } else if (epi_mode_can_tma_store(MODE) && ep.tma_store)
            epi_tile_ts<MODE, (T2_BN / 32) / (EW / 4), EW / 4, t2_is_gg(MODE)>(ep, m, m0 + quarter * 32, n0, sub, taddr, bias_s, ops,
                                                              bar_base + T2_AUX_BYTES + (uint32_t)((warp - 2) * t2_stage_bytes_per_warp(MODE)), &tma_d, &tma_x, lane);
          else
            epi_tile<MODE, (T2_BN / 32) / (EW / 4), EW / 4>(ep, m, n0, sub, taddr, bias_s, ops);
        }
      } else {
        EpiOperands cur, nxt;
        if (n0 + half * 32 < N) epi_load(ep, m, n0 + half * 32, cur);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = half; c < T2_BN / 32; c += 2) {
          if (n0 + c * 32 >= N || skip) break;
          if (c + 2 < T2_BN / 32 && n0 + (c + 2) * 32 < N) epi_load(ep, m, n0 + (c + 2) * 32, nxt);
          epi_chunk32(ep, m, n0 + c * 32, taddr + (uint32_t)(c * 32), bias_s + c * 32, cur);
          cur = nxt;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_cta(tempty_bar(acc), 0));   // the leader's MMA thread waits for both CTAs
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (epi_mode_can_tma_store(MODE) && ep.tma_store && lane == 0) bulk_wait_all();   // the staged rows have left shared memory
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

template <bool AMN, bool BMN, int MODE>
static int launch_tc2_mode(const bfk_gemm_args* a, cudaStream_t s, const CUtensorMap& ma, const CUtensorMap& mb, const EpiParams& ep_in) {
  EpiParams ep = ep_in;
  // TMA-store epilogue (bf16 outputs): 32 x 16 boxes, SWIZZLE_32B; needs 16-byte aligned rows.  BFK_GEMM_TMA_STORE=0 keeps the
  // register stores (A/B knob).
  CUtensorMap md = ma, mx = ma;
  static const bool ts_on = [] { const char* e = getenv("BFK_GEMM_TMA_STORE"); return !(e && e[0] == '0'); }();
  if (epi_mode_can_tma_store(MODE) && ts_on && ep.vec_ok && (a->ldd % 8) == 0 && (reinterpret_cast<uintptr_t>(a->D) & 15u) == 0) {
    const bool needs_x = MODE == EPI_GELUG_TANH_BF16 || MODE == EPI_GELUG_ERF_BF16 ||
                         ((MODE == EPI_GELU_TANH_BF16 || MODE == EPI_GELU_ERF_BF16) && a->aux != nullptr);
    bool ok = make_map(&md, a->D, a->N, a->M, a->ldd, 16, 32, 32);
    if (ok && needs_x) ok = (a->ld_aux % 8) == 0 && (reinterpret_cast<uintptr_t>(a->aux) & 15u) == 0 && make_map(&mx, a->aux, a->N, a->M, a->ld_aux, 16, 32, 32);
    ep.tma_store = ok ? 1 : 0;
  }
  // TMA-load of the epilogue operand (fp32 residual / bf16 aux factor): 32 x 32 boxes.  BFK_GEMM_TMA_LOAD=0 keeps the register loads.
  static const bool tl_on = [] { const char* e = getenv("BFK_GEMM_TMA_LOAD"); return !(e && e[0] == '0'); }();
  if (epi_mode_can_tma_load(MODE) && tl_on && ep.vec_ok && !(ep.vec_ok & 16)) {
    bool ok;
    if (MODE == EPI_RES_F32)
      ok = a->residual && a->res_row_mod == 0 && (a->ldr % 4) == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15u) == 0 &&
           make_map(&mx, a->residual, a->N, a->M, a->ldr, 32, 32, 128, 4);
    else
      ok = a->aux && (a->ld_aux % 8) == 0 && (reinterpret_cast<uintptr_t>(a->aux) & 15u) == 0 && make_map(&mx, a->aux, a->N, a->M, a->ld_aux, 32, 32, 64, 2);
    ep.tma_load = ok ? 1 : 0;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<AMN, BMN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, t2_smem_bytes(MODE));
    if (e != cudaSuccess) { set_error("gemm_tc2: cannot set smem %d: %s", t2_smem_bytes(MODE), cudaGetErrorString(e)); return BFK_E_CUDA; }
    attr_set = true;
  }
  const int items = ((a->M + 255) / 256) * ((a->N + T2_BN - 1) / T2_BN);
  const int grid = 2 * std::min(items, sm_count() / 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + 32 * epi_warps(MODE)); cfg.dynamicSmemBytes = t2_smem_bytes(MODE); cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<AMN, BMN, MODE>, ma, mb, md, mx, a->M, a->N, a->K, ep);
  if (le != cudaSuccess) { set_error("gemm_tc2: launch failed: %s", cudaGetErrorString(le)); return BFK_E_CUDA; }
  BFK_CHECK_LAUNCH("gemm_tc2");
  return BFK_OK;
}

// Specialised epilogues are instantiated only for the operand layouts that use them:
//   forward (A, B K-major): store bf16, bias + residual fp32, bias + GELU (+ derivative);   dgrad (B MN-major): store bf16,
//   dGELU, multiply by aux;
//   wgrad (A, B MN-major): store fp32.
template <bool AMN, bool BMN>
static int launch_tc2(const bfk_gemm_args* a, cudaStream_t s) {
  CUtensorMap ma, mb;
  bool ok;
  if (!AMN) ok = make_map(&ma, a->A, a->K, a->M, a->lda, 64, 128);
  else      ok = make_map(&ma, a->A, a->M, a->K, a->lda, 64, 64);
  if (!ok) return BFK_E_CUDA;
  if (!BMN) ok = make_map(&mb, a->B, a->K, a->N, a->ldb, 64, 128);
  else      ok = make_map(&mb, a->B, a->N, a->K, a->ldb, 64, 64);
  if (!ok) return BFK_E_CUDA;
  EpiParams ep;
  epi_from_args(a, ep);
  if (gemm_dbg() & 32) ep.mode = EPI_GENERIC;     // measurement only: force the generic epilogue
  if (ep.vec_ok) ep.vec_ok |= gemm_dbg() & 92;     // bit 6 (64): no L2 prefetch of the next tile's epilogue operands
#define T2_CASE(MODE_) case MODE_: return launch_tc2_mode<AMN, BMN, MODE_>(a, s, ma, mb, ep)
  if constexpr (!AMN && !BMN) {
    switch (ep.mode) { T2_CASE(EPI_STORE_BF16); T2_CASE(EPI_RES_F32); T2_CASE(EPI_GELU_TANH_BF16); T2_CASE(EPI_GELU_ERF_BF16);
                       T2_CASE(EPI_GELUG_TANH_BF16); T2_CASE(EPI_GELUG_ERF_BF16); default: break; }
  } else if constexpr (!AMN && BMN) {
    switch (ep.mode) { T2_CASE(EPI_STORE_BF16); T2_CASE(EPI_DGELU_TANH_BF16); T2_CASE(EPI_DGELU_ERF_BF16); T2_CASE(EPI_MULAUX_BF16); default: break; }
  } else if constexpr (AMN && BMN) {
    switch (ep.mode) { T2_CASE(EPI_STORE_F32); default: break; }
  }
#undef T2_CASE
  ep.mode = EPI_GENERIC;
  return launch_tc2_mode<AMN, BMN, EPI_GENERIC>(a, s, ma, mb, ep);
}

int gemm_tc2(const bfk_gemm_args* a, cudaStream_t s) {
  const bool ta = a->trans_a != 0, tb = a->trans_b != 0;
  if (!ta && !tb) return launch_tc2<false, false>(a, s);
  if (!ta && tb) return launch_tc2<false, true>(a, s);
  if (ta && tb) return launch_tc2<true, true>(a, s);
  return launch_tc2<true, false>(a, s);
}

}  // namespace bfk
