// tcgen05 flash attention for sm_100a on the packed QKV matrix (bf16 operands, fp32 softmax statistics).
//
// Three kernels, all with the same skeleton (one control warp issuing TMA + tcgen05.mma from a single
// thread, four "row" warps that own one TMEM lane = one matrix row each, mbarrier hand-offs):
//   attn_fwd_tc   : CTA = (128 query rows, head, sample).  Loop over 128-key blocks:
//                   S = Q K^T (TMEM) -> online softmax in registers -> P (bf16, swizzled smem) ->
//                   O += P V (TMEM), O rescaled in TMEM when the running max moves.
//   attn_bwd_dq_tc: same ownership; S and dP = dO V^T recomputed per key block, dS -> smem,
//                   dQ += dS K (K tile re-read MN-major).
//   attn_bwd_dkv_tc: CTA = (128 key rows, head, sample).  Loop over 128-query blocks with the
//                   transposed products S^T = K Q^T, dP^T = V dO^T, then dV += P^T dO, dK += dS^T Q.
// Q/K/V/dO tiles are fetched by 4-D TMA maps {head_dim, slot, token, sample} straight from the packed
// [rows, 3*H*hd] layout; the 64-wide (128-byte swizzle) box zero-fills beyond head_dim and beyond
// the sample's last token, so head dims 48 / 72 and ragged sequence lengths need no padding copies.
// Query-row mask (bifold/models/transformer.py:33-35): a masked row has p == 1/N for every key and dS == 0.
#include <mutex>
#include <unordered_map>

#include "attention_common.cuh"

namespace bfk {

#ifndef BFK_AT_SINGLE_READ
#define BFK_AT_SINGLE_READ 0                    // A/B knob: 1 = keep the S chunk in registers between the maximum and the exponential pass (measured: the 32 live registers spill at the 72-register cap of three CTAs per SM: vision fwd 56 vs 46 us)
#endif
#ifndef BFK_AT_LAZY
#define BFK_AT_LAZY 8.f                         // A/B knob: 0 = move the reference with every block maximum (round-1 behaviour)
#endif
constexpr float AT_LAZY = BFK_AT_LAZY;          // forward: log2 slack of the lazily updated softmax reference
constexpr int AT_ROW_WARPS = 8;                 // two warps per TMEM lane quarter, alternating 32-column chunks
constexpr int AT_THREADS = 32 * AT_ROW_WARPS + 32;  // + control warp (TMA, tcgen05.mma issue, TMEM alloc)

struct AttnTcParams {
  int B, N, H, hd;
  int ksteps;  // ceil(hd/16): UMMA_K steps over the head dim
  int n_o;     // ksteps*16: MMA N of the head-dim-wide products
  float scale;
  const float* row_mask;
  float* lse;
  const float* delta;
  __nv_bfloat16* out; long long ld_out;
  __nv_bfloat16* dqkv; long long ld_dqkv;
};

struct AttnSmem {
  uint32_t bar_q, bar_kv[2], bar_s, bar_p, bar_o, tmem_slot;
};

// =====================================================================================================
// forward
// =====================================================================================================
// BW = key-block width.  BW = 64: 128 TMEM columns and 64 KB of shared memory per CTA (three CTAs per SM), and each
// row warp owns exactly one 32-column chunk of S per step.
template <int NA, int BW>
__global__ void __launch_bounds__(AT_THREADS, (NA == 1 && BW == 64) ? 3 : 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_kv, AttnTcParams p) {
  pdl_trigger();   // the GEMM behind this kernel may set itself up during its tail (common.cuh)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t TILE = NA * AT_ATOM;                    // Q: 128 rows
  constexpr uint32_t KATOM = BW * 128, KTILE = NA * KATOM;   // K, V blocks: BW rows
  constexpr uint32_t P_BYTES = (BW / 64) * AT_ATOM;          // P: 128 rows x BW keys (bf16)
  constexpr int TCOLS = (BW + NA * 64) <= 128 ? 128 : 256;
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t sQ = base, sK = sQ + TILE, sV = sK + 2 * KTILE, sP = sV + 2 * KTILE, bars = sP + P_BYTES;
  const uint32_t bar_q = bars, bar_k0 = bars + 8, bar_v0 = bars + 24, bar_s = bars + 40, bar_p = bars + 48, bar_o = bars + 56,
                 tmem_slot = bars + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nblk = (p.N + BW - 1) / BW;

  if (threadIdx.x == 0) {
    if (base & 1023u) __trap();
    mbar_init(bar_q, 1);
    mbar_init(bar_k0, 1); mbar_init(bar_k0 + 8, 1); mbar_init(bar_v0, 1); mbar_init(bar_v0 + 8, 1);
    mbar_init(bar_s, 1); mbar_init(bar_p, 32 * AT_ROW_WARPS); mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == AT_ROW_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - base));
  const uint32_t tS = tmem, tO = tmem + BW;

  if (warp == AT_ROW_WARPS) {
    if (elect_one()) {
      // Issue order: ... P V(j) | S(j+1) -> commit(bar_s) ...; tcgen05 ops of one thread complete in order, so
      // bar_s(j+1) also tells the row warps that P V(j) is done (P tile free, O safe to rescale).
      // Descriptors are built once; inside the loops an operand advances by a constant (this thread's instruction
      // stream is the critical path between the row warps' hand-over and the next S tile).
      const uint32_t idesc_o = make_idesc(128, p.n_o, false, true);
      const uint64_t dq0 = kmajor_desc(sQ, 0), dk0 = kmajor_desc(sK, 0, KATOM), dp0 = kmajor_desc(sP, 0), dv0 = mnmajor_desc(sV, 0, KATOM);
      auto issue_k = [&](int j) {
        mbar_expect_tx(bar_k0 + 8 * (j & 1), KTILE);
        load_tile<NA>(sK + (j & 1) * KTILE, &map_kv, bar_k0 + 8 * (j & 1), p.H + h, j * BW, b, KATOM);
      };
      auto issue_v = [&](int j) {
        mbar_expect_tx(bar_v0 + 8 * (j & 1), KTILE);
        load_tile<NA>(sV + (j & 1) * KTILE, &map_kv, bar_v0 + 8 * (j & 1), 2 * p.H + h, j * BW, b, KATOM);
      };
      auto issue_s = [&](int j) {
        mbar_wait(bar_k0 + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        const uint32_t idesc_sj = make_idesc(128, kv_cols(p.N, j, BW), false, false);
        uint64_t da = dq0, db = dk0 + (uint64_t)((j & 1) * (KTILE >> 4));
        for (int k = 0; k < p.ksteps; ++k) {
          tc_mma_bf16(tS, da, db, idesc_sj, k > 0);
          da += 2; db += 2;                                         // next 16 elements of K: +32 bytes
          if ((k & 3) == 3) { da += (AT_ATOM >> 4) - 8; db += (KATOM >> 4) - 8; }   // next 64-wide atom
        }
        tc_commit(bar_s);
      };
      mbar_expect_tx(bar_q, TILE);
      load_tile<NA>(sQ, &map_qkv, bar_q, h, q0, b);
      issue_k(0); issue_v(0);
      if (nblk > 1) { issue_k(1); issue_v(1); }
      mbar_wait(bar_q, 0);
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(bar_p, j & 1);     // P(j) in smem, S(j) read; S(j) done => K buffer j&1 free, P V(j-1) done => V buffer (j-1)&1 free
        tc_fence_after();
        if (j + 2 < nblk) issue_k(j + 2);
        if (j >= 1 && j + 1 < nblk) issue_v(j + 1);
        mbar_wait(bar_v0 + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        const int nk = kv_cols(p.N, j, BW) >> 4;
        uint64_t da = dp0, db = dv0 + (uint64_t)((j & 1) * (KTILE >> 4));
        for (int k = 0; k < nk; ++k) {
          tc_mma_bf16(tO, da, db, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
          da += 2; db += 2048 >> 4;                                 // P: next 16 keys (+32 bytes); V: next 16 rows (+2048 bytes)
          if ((k & 3) == 3) da += (AT_ATOM >> 4) - 8;
        }
        if (j + 1 < nblk) issue_s(j + 1); else tc_commit(bar_o);
      }
    }
  } else {
    // ---- row threads: quarter = warp % 4 owns TMEM lanes / query rows [32q, 32q+32); the two warps of a quarter
    //      (half 0 / 1) take alternating 32-column chunks and exchange row max / row sum through `xch` ----
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane, q = q0 + r;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    float* xch = reinterpret_cast<float*>(smem_raw + (bars + 128 - base));   // [128]
    __nv_bfloat16* xch16 = reinterpret_cast<__nv_bfloat16*>(xch);            // [2][128] view of the same 512 bytes
    const bool quarter_live = q0 + quarter * 32 < p.N;      // warp-uniform: any real query row in this quarter?
    const bool masked = p.row_mask && q < p.N && p.row_mask[(long long)b * p.N + q] == 0.f;
    // masked query row: every score equal => p == 1 for every key (scale 0, max 0)
    const float sl2 = masked ? 0.f : p.scale * LOG2E;
    float m_run = -INFINITY, l_run = 0.f;   // running max (scaled, log2 domain) and this warp's partial row sum
    for (int j = 0; j < nblk; ++j) {
      const int nc = kv_cols(p.N, j, BW);          // columns computed for this block (multiple of 16)
      const int nvalid = min(BW, p.N - j * BW);    // real keys in this block
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      if (!quarter_live) {                         // rows beyond the sequence: stay in lock-step, do nothing
        mbar_arrive(bar_p);
        continue;
      }
      float mx = -INFINITY, rowsum = 0.f, m_new, alpha;
      auto exchange_max = [&]() {
        // The two halves of a row must subtract the SAME reference; it need not be the exact maximum (the result is
        // invariant to it), so each half publishes its max rounded UP to bf16 and both take the larger one: one
        // 64-thread barrier, 512 bytes of exchange space.
        __nv_bfloat16 mb = __float2bfloat16_ru(mx);
        xch16[half * 128 + r] = mb;
        named_bar_sync(1 + quarter, 64);
        mx = fmaxf(__bfloat162float(mb), __bfloat162float(xch16[(half ^ 1) * 128 + r]));
        // LAZY reference: the running reference only moves when the block maximum exceeds it by more than 2^AT_LAZY (softmax is
        // invariant to the reference as long as P, the row sum and O use the same one; exp2 of at most AT_LAZY stays far inside
        // bf16 / fp32 range).  After the first block almost no step changes it, so the TMEM round trip that rescales O goes away.
        const float cand = mx * sl2;
        m_new = masked ? 0.f : ((j == 0 || cand > m_run + AT_LAZY) ? fmaxf(m_run, cand) : m_run);
        alpha = ex2f(m_run - m_new);
      };
      auto pass2 = [&](const uint32_t (&sreg)[32], int c) {
        const int ng = min(4, (nc - c) >> 3);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g < ng) {
            float v[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = ex2f(fmaf(__uint_as_float(sreg[g * 8 + t]), sl2, -m_new));
            if (c + g * 8 + 8 > nvalid) {   // tail: keys beyond the sequence get p = 0
#pragma unroll
              for (int t = 0; t < 8; ++t) if (c + g * 8 + t >= nvalid) v[t] = 0.f;
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) rowsum += v[t];
            store_p_chunk(sP, r, (c >> 3) + g, v);
          }
        }
      };
      if constexpr (BW == 64 && BFK_AT_SINGLE_READ) {
        // one 32-column chunk per warp and block: S is read from TMEM ONCE, the maximum and the exponentials both come from
        // the same registers
        const int c = half * 32;
        uint32_t sreg[32];
        if (c < nc) {
          tc_ld32(tS + lane_off + c, sreg);
          tc_wait_ld();
          const int lim = nvalid - c;
          if (lim >= 32) {
#pragma unroll
            for (int t = 0; t < 32; ++t) mx = fmaxf(mx, __uint_as_float(sreg[t]));
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t)
              if (t < lim) mx = fmaxf(mx, __uint_as_float(sreg[t]));
          }
        }
        exchange_max();
        if (c < nc) pass2(sreg, c);
      } else {
        // pass 1: raw row max over the real keys of this warp's chunks
#pragma unroll 1
        for (int c = half * 32; c < nc; c += 64) {
          uint32_t sreg[32];
          tc_ld32(tS + lane_off + c, sreg);
          tc_wait_ld();
          const int lim = nvalid - c;
          if (lim >= 32) {                       // full chunk (all but the sequence tail): no per-element predicate
#pragma unroll
            for (int t = 0; t < 32; ++t) mx = fmaxf(mx, __uint_as_float(sreg[t]));
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t)
              if (t < lim) mx = fmaxf(mx, __uint_as_float(sreg[t]));
          }
        }
        exchange_max();
#pragma unroll 1
        for (int c = half * 32; c < nc; c += 64) {
          uint32_t sreg[32];
          tc_ld32(tS + lane_off + c, sreg);
          tc_wait_ld();
          pass2(sreg, c);
        }
      }
      l_run = l_run * alpha + rowsum;
      m_run = m_new;
      if (j > 0 && !__all_sync(0xffffffffu, alpha == 1.f)) {
        for (int c = half; c < p.n_o / 16; c += 2) {   // the two halves rescale alternating 16-column chunks of O
          uint32_t o[16];
          tc_ld16(tO + lane_off + c * 16, o);
          tc_wait_ld();
#pragma unroll
          for (int t = 0; t < 16; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
          tc_st16(tO + lane_off + c * 16, o);
        }
        tc_wait_st();
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    mbar_wait(bar_o, 0);
    tc_fence_after();
    if (quarter_live) {
      // total row sum = sum of the two halves' partial sums (same rescaling history)
      if (half == 1) xch[r] = l_run;
      named_bar_sync(1 + quarter, 64);
      if (half == 0) { l_run += xch[r]; xch[r] = l_run; }
      named_bar_sync(1 + quarter, 64);
      if (half == 1) l_run = xch[r];
      const float inv = 1.f / l_run;
      __nv_bfloat16* orow = p.out + ((long long)b * p.N + q) * p.ld_out + h * p.hd;
      for (int c = half; c < p.n_o / 16; c += 2) {
        uint32_t o[16];
        tc_ld16(tO + lane_off + c * 16, o);
        tc_wait_ld();
        if (q < p.N) store_row16(orow, c * 16, p.hd, o, inv);
      }
      if (half == 0 && q < p.N && p.lse) p.lse[((long long)b * p.H + h) * p.N + q] = m_run * LN2 + logf(l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AT_ROW_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TCOLS) : "memory");
  }
}

// =====================================================================================================
// backward: dQ
// =====================================================================================================
// BW = key-block width.  BW = 64 keeps the TMEM footprint at 256 columns (S, dP: 64 each, dQ: up to 128), so two CTAs
// share an SM and one's MMA / TMA / barrier latencies hide behind the other's softmax math.
// Issue order of the control thread after the row warps hand over dS(j):  S(j+1), dP(j+1) -> commit(bar_s) | dQ(j) ->
// commit(bar_dq).  The products the row warps wait for go first; dS is double-buffered so dQ(j) may still read buffer
// j&1 while the rows fill buffer (j+1)&1 (tcgen05 ops of one thread complete in order: bar_s(j+2) implies dQ(j) done).
template <int NA, int BW, int NBUF>
__global__ void __launch_bounds__(AT_THREADS, NA == 1 && BW == 64 ? 2 : 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                      const __grid_constant__ CUtensorMap map_do, AttnTcParams p) {
  pdl_trigger();   // the GEMM behind this kernel may set itself up during its tail (common.cuh)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t TILE = NA * AT_ATOM;             // Q, dO: 128 rows
  constexpr uint32_t KATOM = BW * 128, KTILE = NA * KATOM;   // K, V blocks: BW rows
  constexpr uint32_t DS_BYTES = (BW / 64) * AT_ATOM;  // dS: 128 rows x BW keys (bf16)
  constexpr int TCOLS = (2 * BW + NA * 64) <= 256 ? 256 : 512;
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t sQ = base, sdO = sQ + TILE, sK = sdO + TILE, sV = sK + NBUF * KTILE, sdS = sV + NBUF * KTILE, bars = sdS + 2 * DS_BYTES;
  const uint32_t bar_q = bars, bar_kv0 = bars + 8, bar_s = bars + 40, bar_p = bars + 48, bar_o = bars + 56, bar_dq = bars + 72, tmem_slot = bars + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nblk = (p.N + BW - 1) / BW;

  if (threadIdx.x == 0) {
    if (base & 1023u) __trap();
    mbar_init(bar_q, 1);
    for (int i = 0; i < NBUF; ++i) mbar_init(bar_kv0 + 8 * i, 1);
    mbar_init(bar_s, 1); mbar_init(bar_p, 32 * AT_ROW_WARPS); mbar_init(bar_o, 1); mbar_init(bar_dq, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == AT_ROW_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - base));
  const uint32_t tS = tmem, tdP = tmem + BW, tdQ = tmem + 2 * BW;

  if (warp == AT_ROW_WARPS) {
    if (elect_one()) {
      const uint32_t idesc_q = make_idesc(128, p.n_o, false, true);
      // descriptors built once; an operand advances by a constant inside the loops (see the forward kernel)
      const uint64_t dq0 = kmajor_desc(sQ, 0), ddo0 = kmajor_desc(sdO, 0), dk0 = kmajor_desc(sK, 0, KATOM), dv0 = kmajor_desc(sV, 0, KATOM),
                     dds0 = kmajor_desc(sdS, 0), dkm0 = mnmajor_desc(sK, 0, KATOM);
      auto issue_kv = [&](int j) {
        const int nb = j % NBUF;
        mbar_expect_tx(bar_kv0 + 8 * nb, 2 * KTILE);
        load_tile<NA>(sK + nb * KTILE, &map_kv, bar_kv0 + 8 * nb, p.H + h, j * BW, b, KATOM);
        load_tile<NA>(sV + nb * KTILE, &map_kv, bar_kv0 + 8 * nb, 2 * p.H + h, j * BW, b, KATOM);
      };
      auto issue_s = [&](int j) {
        const int nb = j % NBUF;
        mbar_wait(bar_kv0 + 8 * nb, (j / NBUF) & 1);
        tc_fence_after();
        const uint32_t idesc_sj = make_idesc(128, kv_cols(p.N, j, BW), false, false);
        const uint64_t boff = (uint64_t)(nb * (KTILE >> 4));
        uint64_t da = dq0, db = dk0 + boff;
        for (int k = 0; k < p.ksteps; ++k) {
          tc_mma_bf16(tS, da, db, idesc_sj, k > 0);
          da += 2; db += 2;
          if ((k & 3) == 3) { da += (AT_ATOM >> 4) - 8; db += (KATOM >> 4) - 8; }
        }
        da = ddo0; db = dv0 + boff;
        for (int k = 0; k < p.ksteps; ++k) {
          tc_mma_bf16(tdP, da, db, idesc_sj, k > 0);
          da += 2; db += 2;
          if ((k & 3) == 3) { da += (AT_ATOM >> 4) - 8; db += (KATOM >> 4) - 8; }
        }
        tc_commit(bar_s);
      };
      mbar_expect_tx(bar_q, 2 * TILE);
      load_tile<NA>(sQ, &map_q, bar_q, h, q0, b, AT_ATOM);
      load_tile<NA>(sdO, &map_do, bar_q, h, q0, b, AT_ATOM);
      for (int j = 0; j < NBUF && j < nblk; ++j) issue_kv(j);
      mbar_wait(bar_q, 0);
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(bar_p, j & 1);      // dS(j) is in smem buffer j&1, S(j)/dP(j) have been read
        tc_fence_after();
        if (j + 1 < nblk) issue_s(j + 1);
        if (j >= 1) {                 // every bar_dq phase is consumed, in order, BEFORE the next one is committed
          mbar_wait(bar_dq, (j - 1) & 1);
          if (j + NBUF - 1 < nblk) issue_kv(j + NBUF - 1);   // dQ(j-1) done: refill the buffer block j-1 used
        }
        const int nk = kv_cols(p.N, j, BW) >> 4;
        uint64_t da = dds0 + (uint64_t)((j & 1) * (DS_BYTES >> 4)), db = dkm0 + (uint64_t)((j % NBUF) * (KTILE >> 4));
        for (int k = 0; k < nk; ++k) {
          tc_mma_bf16(tdQ, da, db, idesc_q, (j > 0 || k > 0) ? 1u : 0u);
          da += 2; db += 2048 >> 4;
          if ((k & 3) == 3) da += (AT_ATOM >> 4) - 8;
        }
        if (j + 1 < nblk) tc_commit(bar_dq); else tc_commit(bar_o);
      }
    }
  } else {
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane, q = q0 + r;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const bool quarter_live = q0 + quarter * 32 < p.N;
    const bool live = q < p.N && !(p.row_mask && p.row_mask[(long long)b * p.N + q] == 0.f);
    const float sl2 = p.scale * LOG2E;
    const long long si = ((long long)b * p.H + h) * p.N + q;
    // non-live rows (masked query / beyond the sequence): lse = +inf => p == 0 => dS == 0
    const float lse2 = live ? p.lse[si] * LOG2E : INFINITY, delta = live ? p.delta[si] : 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int nc = kv_cols(p.N, j, BW);
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      // keys in [N, nc) have zero K rows: their dS values multiply zero rows in dQ += dS K, no masking needed.
      // Quarters without any real query row skip the math (their dS rows only feed dQ rows that are never stored).
      if (quarter_live) {
#pragma unroll 1
        for (int c = half * 32; c < nc; c += 64) {
          uint32_t sreg[32], d[32];
          tc_ld32(tS + lane_off + c, sreg);
          tc_ld32(tdP + lane_off + c, d);
          tc_wait_ld();
          const int ng = min(4, (nc - c) >> 3);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < ng) {
              float v[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const float pr = ex2f(fmaf(__uint_as_float(sreg[g * 8 + t]), sl2, -lse2));
                v[t] = pr * (__uint_as_float(d[g * 8 + t]) - delta);
              }
              store_p_chunk(sdS + (j & 1) * DS_BYTES, r, (c >> 3) + g, v);
            }
          }
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    mbar_wait(bar_o, 0);
    tc_fence_after();
    __nv_bfloat16* drow = p.dqkv + ((long long)b * p.N + q) * p.ld_dqkv + h * p.hd;
    for (int c = half; c < p.n_o / 16; c += 2) {
      uint32_t o[16];
      tc_ld16(tdQ + lane_off + c * 16, o);
      tc_wait_ld();
      if (q < p.N) store_row16(drow, c * 16, p.hd, o, p.scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AT_ROW_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TCOLS) : "memory");
  }
}

// =====================================================================================================
// backward: dK, dV
// =====================================================================================================
// CTA = 128 key rows; loop over BW-wide query blocks (BW = 64: 256 TMEM columns, two CTAs per SM, same issue order as dQ).
template <int NA, int BW, int NBUF>
__global__ void __launch_bounds__(AT_THREADS, NA == 1 && BW == 64 ? 2 : 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_q,
                       const __grid_constant__ CUtensorMap map_do, AttnTcParams p) {
  pdl_trigger();   // the GEMM behind this kernel may set itself up during its tail (common.cuh)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t TILE = NA * AT_ATOM;             // K, V: 128 rows
  constexpr uint32_t QATOM = BW * 128, QTILE = NA * QATOM;   // Q, dO blocks: BW rows
  constexpr uint32_t PT_BYTES = (BW / 64) * AT_ATOM;  // P^T, dS^T: 128 key rows x BW queries (bf16)
  constexpr int TCOLS = (2 * BW + 2 * NA * 64) <= 256 ? 256 : 512;
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t sK = base, sV = sK + TILE, sQ = sV + TILE, sdO = sQ + NBUF * QTILE, sPt = sdO + NBUF * QTILE, sdSt = sPt + PT_BYTES,
                 bars = sdSt + PT_BYTES;
  const uint32_t bar_kv = bars, bar_q0 = bars + 8, bar_s = bars + 40, bar_p = bars + 48, bar_o = bars + 56, tmem_slot = bars + 64;
  float2* s_col = reinterpret_cast<float2*>(smem_raw + (bars - base) + 128);          // per-query-column {lse2, delta}
  uint32_t* s_live = reinterpret_cast<uint32_t*>(smem_raw + (bars - base) + 128 + BW * 8);   // live-query bits, BW/32 words
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nq = (p.N + BW - 1) / BW;

  if (threadIdx.x == 0) {
    if (base & 1023u) __trap();
    mbar_init(bar_kv, 1);
    for (int i = 0; i < NBUF; ++i) mbar_init(bar_q0 + 8 * i, 1);
    mbar_init(bar_s, 1); mbar_init(bar_p, 32 * AT_ROW_WARPS); mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == AT_ROW_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - base));
  const uint32_t tSt = tmem, tdPt = tmem + BW, tdV = tmem + 2 * BW, tdK = tdV + NA * 64;

  if (warp == AT_ROW_WARPS) {
    if (elect_one()) {
      const uint32_t idesc_g = make_idesc(128, p.n_o, false, true);
      const uint64_t dk0 = kmajor_desc(sK, 0), dv0 = kmajor_desc(sV, 0), dq0 = kmajor_desc(sQ, 0, QATOM), ddo0 = kmajor_desc(sdO, 0, QATOM),
                     dpt0 = kmajor_desc(sPt, 0), ddst0 = kmajor_desc(sdSt, 0), dqm0 = mnmajor_desc(sQ, 0, QATOM),
                     ddom0 = mnmajor_desc(sdO, 0, QATOM);
      auto issue_q = [&](int i) {
        const int nb = i % NBUF;
        mbar_expect_tx(bar_q0 + 8 * nb, 2 * QTILE);
        load_tile<NA>(sQ + nb * QTILE, &map_q, bar_q0 + 8 * nb, h, i * BW, b, QATOM);
        load_tile<NA>(sdO + nb * QTILE, &map_do, bar_q0 + 8 * nb, h, i * BW, b, QATOM);
      };
      auto issue_s = [&](int i) {
        const int nb = i % NBUF;
        mbar_wait(bar_q0 + 8 * nb, (i / NBUF) & 1);
        tc_fence_after();
        const uint32_t idesc_si = make_idesc(128, kv_cols(p.N, i, BW), false, false);
        const uint64_t boff = (uint64_t)(nb * (QTILE >> 4));
        uint64_t da = dk0, db = dq0 + boff;
        for (int k = 0; k < p.ksteps; ++k) {
          tc_mma_bf16(tSt, da, db, idesc_si, k > 0);
          da += 2; db += 2;
          if ((k & 3) == 3) { da += (AT_ATOM >> 4) - 8; db += (QATOM >> 4) - 8; }
        }
        da = dv0; db = ddo0 + boff;
        for (int k = 0; k < p.ksteps; ++k) {
          tc_mma_bf16(tdPt, da, db, idesc_si, k > 0);
          da += 2; db += 2;
          if ((k & 3) == 3) { da += (AT_ATOM >> 4) - 8; db += (QATOM >> 4) - 8; }
        }
        tc_commit(bar_s);
      };
      mbar_expect_tx(bar_kv, 2 * TILE);
      load_tile<NA>(sK, &map_kv, bar_kv, p.H + h, j0, b, AT_ATOM);
      load_tile<NA>(sV, &map_kv, bar_kv, 2 * p.H + h, j0, b, AT_ATOM);
      for (int i = 0; i < NBUF - 1 && i < nq; ++i) issue_q(i);
      mbar_wait(bar_kv, 0);
      issue_s(0);
      for (int i = 0; i < nq; ++i) {
        mbar_wait(bar_p, i & 1);
        tc_fence_after();
        if (i + NBUF - 1 < nq) issue_q(i + NBUF - 1);
        const int nk = kv_cols(p.N, i, BW) >> 4;
        const uint64_t boff = (uint64_t)((i % NBUF) * (QTILE >> 4));
        uint64_t da = dpt0, db = ddom0 + boff;
        for (int k = 0; k < nk; ++k) {
          tc_mma_bf16(tdV, da, db, idesc_g, (i > 0 || k > 0) ? 1u : 0u);
          da += 2; db += 2048 >> 4;
          if ((k & 3) == 3) da += (AT_ATOM >> 4) - 8;
        }
        da = ddst0; db = dqm0 + boff;
        for (int k = 0; k < nk; ++k) {
          tc_mma_bf16(tdK, da, db, idesc_g, (i > 0 || k > 0) ? 1u : 0u);
          da += 2; db += 2048 >> 4;
          if ((k & 3) == 3) da += (AT_ATOM >> 4) - 8;
        }
        if (i + 1 < nq) issue_s(i + 1); else tc_commit(bar_o);
      }
    }
  } else {
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane, kv = j0 + r;
    const int tid = threadIdx.x;                           // 0..255; threads [0, BW) fetch the per-query statistics
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    const float log2n = log2f((float)p.N);
    const bool kv_ok = kv < p.N;
    const bool quarter_live = j0 + quarter * 32 < p.N;     // any real key row in this quarter?
    // per-query-column statistics:   live query  : {lse*log2e, delta}, live bit 1   p = 2^(s*scale2 - lse2), dS = p (dP - delta)
    //                                masked query: {log2 N, 0},       live bit 0   p = 1/N,                 dS = 0
    //                                out of range: {+inf, 0},         live bit 0   p = 0
    auto fetch = [&](int i, float2& st, bool& lv) {
      const int q = i * BW + tid;
      st = make_float2(INFINITY, 0.f);
      lv = false;
      if (q < p.N) {
        const long long si = ((long long)b * p.H + h) * p.N + q;
        const bool m = p.row_mask && p.row_mask[(long long)b * p.N + q] == 0.f;
        if (m) st = make_float2(log2n, 0.f);
        else { st = make_float2(p.lse[si] * LOG2E, p.delta[si]); lv = true; }
      }
    };
    float2 st_next = make_float2(INFINITY, 0.f);
    bool lv_next = false;
    if (tid < BW) fetch(0, st_next, lv_next);
    for (int i = 0; i < nq; ++i) {
      const int nc = kv_cols(p.N, i, BW);
      named_bar_sync(1, 32 * AT_ROW_WARPS);                // everybody is done with the previous block's statistics
      if (tid < BW) {                                      // warp-uniform (BW is a multiple of 32)
        s_col[tid] = st_next;
        const uint32_t bits = __ballot_sync(0xffffffffu, lv_next);
        if (lane == 0) s_live[tid >> 5] = bits;
        if (i + 1 < nq) fetch(i + 1, st_next, lv_next);    // in flight while this block is processed
      }
      named_bar_sync(1, 32 * AT_ROW_WARPS);
      mbar_wait(bar_s, i & 1);
      tc_fence_after();
      if (quarter_live) {
#pragma unroll 1
        for (int c = half * 32; c < nc; c += 64) {
          uint32_t sreg[32], d[32];
          tc_ld32(tSt + lane_off + c, sreg);
          tc_ld32(tdPt + lane_off + c, d);
          const uint32_t lbits = s_live[c >> 5];
          tc_wait_ld();
          const int ng = min(4, (nc - c) >> 3);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < ng) {
              float pv[8], dv[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const float2 st = s_col[c + g * 8 + t];
                const bool lv = (lbits >> (g * 8 + t)) & 1u;
                const float pr = ex2f(fmaf(__uint_as_float(sreg[g * 8 + t]), lv ? sl2 : 0.f, -st.x));
                pv[t] = pr;
                dv[t] = lv ? pr * (__uint_as_float(d[g * 8 + t]) - st.y) : 0.f;
              }
              store_p_chunk(sPt, r, (c >> 3) + g, pv);
              store_p_chunk(sdSt, r, (c >> 3) + g, dv);
            }
          }
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const long long HD = (long long)p.H * p.hd;
    __nv_bfloat16* krow = p.dqkv + ((long long)b * p.N + kv) * p.ld_dqkv + HD + h * p.hd;
    __nv_bfloat16* vrow = krow + HD;
    for (int c = 0; c < p.n_o / 16; ++c) {   // half 0 stores dV, half 1 stores dK
      uint32_t o[16];
      tc_ld16((half == 0 ? tdV : tdK) + lane_off + c * 16, o);
      tc_wait_ld();
      if (kv_ok) store_row16(half == 0 ? vrow : krow, c * 16, p.hd, o, half == 0 ? 1.f : p.scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AT_ROW_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TCOLS) : "memory");
  }
}

// delta[b,h,i] = sum_d dout*out.  One thread per 8-element (16-byte) chunk of a token row, whole rows per block
// (coalesced row reads); the chunks of a head are then summed through shared memory.
__global__ void __launch_bounds__(1024) attn_delta_tc_kernel(const __nv_bfloat16* __restrict__ out, long long ld_out,
                                                              const __nv_bfloat16* __restrict__ dout, long long ld_dout,
                                                              float* __restrict__ delta, int B, int N, int H, int hd, int rows_per_block) {
  extern __shared__ float s_part[];
  const int cph = hd >> 3, chunks = H * cph;            // chunks per head / per row
  const int lr = threadIdx.x / chunks, ch = threadIdx.x % chunks;
  const long long row = (long long)blockIdx.x * rows_per_block + lr;
  const bool ok = lr < rows_per_block && row < (long long)B * N;
  float s = 0.f;
  if (ok) {
    const uint4 a = *reinterpret_cast<const uint4*>(out + row * ld_out + ch * 8);
    const uint4 c = *reinterpret_cast<const uint4*>(dout + row * ld_dout + ch * 8);
    const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* cp = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 x = __bfloat1622float2(ap[t]), y = __bfloat1622float2(cp[t]);
      s += x.x * y.x + x.y * y.y;
    }
  }
  s_part[threadIdx.x] = s;
  __syncthreads();
  if (ok && ch % cph == 0) {
    float t = 0.f;
    for (int k = 0; k < cph; ++k) t += s_part[threadIdx.x + k];
    const int hh = ch / cph;
    const int bb = (int)(row / N), i = (int)(row % N);
    delta[((long long)bb * H + hh) * N + i] = t;
  }
}

// ---- host ---------------------------------------------------------------------------------------------
struct Map4Key {
  const void* ptr; int hd, slots, N, B; long long ld; int rows;
  bool operator==(const Map4Key& o) const { return ptr == o.ptr && hd == o.hd && slots == o.slots && N == o.N && B == o.B && ld == o.ld && rows == o.rows; }
};
struct Map4Hash {
  size_t operator()(const Map4Key& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&](size_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.hd); mix(k.slots); mix(k.N); mix(k.B); mix((size_t)k.ld); mix(k.rows);
    return h;
  }
};
// 4-D bf16 map {hd, slots, N, B} over a [B*N, ld] row-major matrix whose row holds `slots` runs of hd elements
bool make_map4(CUtensorMap* out, const void* ptr, int hd, int slots, int N, int B, long long ld, int box_rows) {
  static std::mutex mu;
  static std::unordered_map<Map4Key, CUtensorMap, Map4Hash> cache;
  Map4Key key{ptr, hd, slots, N, B, ld, box_rows};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return true; }
  }
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_error("attention_tc: cuTensorMapEncodeTiled unavailable"); return false; }
  cuuint64_t dims[4] = {(cuuint64_t)hd, (cuuint64_t)slots, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)hd * 2, (cuuint64_t)ld * 2, (cuuint64_t)N * ld * 2};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("attention_tc: cuTensorMapEncodeTiled failed (%d) hd=%d slots=%d N=%d B=%d ld=%lld", (int)r, hd, slots, N, B, ld); return false; }
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return true;
}

bool attention_tc_supported(const bfk_attention_args* a, bool bwd, const char** why) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if (a->dtype != BFK_BF16) { *why = "bf16 only"; return false; }
  if (a->hd % 8 != 0 || a->hd < 8 || a->hd > 128) { *why = "head_dim must be a multiple of 8 in [8,128]"; return false; }
  if (a->ld_qkv % 8 != 0 || a->ld_out % 8 != 0 || !al16(a->qkv) || !al16(a->out)) { *why = "qkv/out must be 16-byte aligned"; return false; }
  if (bwd && (a->ld_dout % 8 != 0 || a->ld_dqkv % 8 != 0 || !al16(a->dout) || !al16(a->dqkv))) { *why = "dout/dqkv must be 16-byte aligned"; return false; }
  if (a->H > 65535 || a->B > 65535) { *why = "grid too large"; return false; }
  if (!get_encode_tiled()) { *why = "driver entry point cuTensorMapEncodeTiled missing"; return false; }
  return true;
}

static AttnTcParams make_params(const bfk_attention_args* a) {
  AttnTcParams p;
  p.B = a->B; p.N = a->N; p.H = a->H; p.hd = a->hd;
  p.ksteps = (a->hd + 15) / 16;
  p.n_o = p.ksteps * 16;
  p.scale = a->scale;
  p.row_mask = a->row_mask;
  p.lse = a->lse;
  p.delta = a->delta;
  p.out = (__nv_bfloat16*)a->out; p.ld_out = a->ld_out;
  p.dqkv = (__nv_bfloat16*)a->dqkv; p.ld_dqkv = a->ld_dqkv;
  return p;
}

template <typename Kern>
static int set_attr(Kern k, size_t smem, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("%s: cannot set smem %zu: %s", name, smem, cudaGetErrorString(e)); return BFK_E_CUDA; }
  return BFK_OK;
}

static int bwd_block_width() {   // BFK_ATTN_BW=128 selects the one-CTA-per-SM 128-wide variant (measurement only)
  static int v = -1;
  if (v < 0) { const char* e = getenv("BFK_ATTN_BW"); v = (e && atoi(e) == 128) ? 128 : 64; }
  return v;
}

static int fwd_block_width() {   // BFK_ATTN_FWD_BW=128 selects the 128-wide variant (measurement only)
  static int v = -1;
  if (v < 0) { const char* e = getenv("BFK_ATTN_FWD_BW"); v = (e && atoi(e) == 128) ? 128 : 64; }
  return v;
}

template <int NA, int BW>
static int launch_fwd(const bfk_attention_args* a, const AttnTcParams& p, cudaStream_t s) {
  CUtensorMap mq, mk;
  if (!make_map4(&mq, a->qkv, a->hd, 3 * a->H, a->N, a->B, a->ld_qkv, 128)) return BFK_E_CUDA;
  if (!make_map4(&mk, a->qkv, a->hd, 3 * a->H, a->N, a->B, a->ld_qkv, BW)) return BFK_E_CUDA;
  // NA == 1: 66176 bytes (BW 64, three CTAs per SM) / 115328 bytes (BW 128, two CTAs per SM)
  const size_t smem = NA * AT_ATOM + 4 * NA * BW * 128 + (BW / 64) * AT_ATOM + 640;
  if (int rc = set_attr(attn_fwd_tc_kernel<NA, BW>, smem, "attn_fwd_tc")) return rc;
  dim3 grid((a->N + 127) / 128, a->H, a->B);
  attn_fwd_tc_kernel<NA, BW><<<grid, AT_THREADS, smem, s>>>(mq, mk, p);
  return BFK_OK;
}

int attention_fwd_v2(const bfk_attention_args* a, cudaStream_t s);   // attention_v2.cu: persistent kernels for short sequences
int attention_bwd_v2(const bfk_attention_args* a, cudaStream_t s);   // (return 1 when the problem is not eligible)

int attention_fwd_tc(const bfk_attention_args* a, cudaStream_t s) {
  if (const int rc = attention_fwd_v2(a, s); rc <= 0) return rc;
  AttnTcParams p = make_params(a);
  int rc;
  if (a->hd <= 64) rc = fwd_block_width() == 64 ? launch_fwd<1, 64>(a, p, s) : launch_fwd<1, 128>(a, p, s);
  else rc = fwd_block_width() == 64 ? launch_fwd<2, 64>(a, p, s) : launch_fwd<2, 128>(a, p, s);
  if (rc) return rc;
  BFK_CHECK_LAUNCH("attention_fwd_tc");
  return BFK_OK;
}

template <int NA, int BW, int NBUF>
static int launch_bwd(const bfk_attention_args* a, const AttnTcParams& p, cudaStream_t s) {
  CUtensorMap mq128, mqb, md128, mdb;
  if (!make_map4(&mq128, a->qkv, a->hd, 3 * a->H, a->N, a->B, a->ld_qkv, 128)) return BFK_E_CUDA;
  if (!make_map4(&mqb, a->qkv, a->hd, 3 * a->H, a->N, a->B, a->ld_qkv, BW)) return BFK_E_CUDA;
  if (!make_map4(&md128, a->dout, a->hd, a->H, a->N, a->B, a->ld_dout, 128)) return BFK_E_CUDA;
  if (!make_map4(&mdb, a->dout, a->hd, a->H, a->N, a->B, a->ld_dout, BW)) return BFK_E_CUDA;
  constexpr size_t TILE = NA * AT_ATOM, BTILE = NA * BW * 128, STAGE = (BW / 64) * AT_ATOM;
  // NA == 1, BW == 64: 114944 and 115344 bytes -- two CTAs per SM ((233472 - 2*1024)/2 = 115712)
  const size_t smem_q = 2 * TILE + 2 * NBUF * BTILE + 2 * STAGE + 256;
  const size_t smem_kv = 2 * TILE + 2 * NBUF * BTILE + 2 * STAGE + 128 + BW * 8 + 16;
  if (int rc = set_attr(attn_bwd_dq_tc_kernel<NA, BW, NBUF>, smem_q, "attn_bwd_dq_tc")) return rc;
  if (int rc = set_attr(attn_bwd_dkv_tc_kernel<NA, BW, NBUF>, smem_kv, "attn_bwd_dkv_tc")) return rc;
  dim3 grid((a->N + 127) / 128, a->H, a->B);
  attn_bwd_dq_tc_kernel<NA, BW, NBUF><<<grid, AT_THREADS, smem_q, s>>>(mq128, mqb, md128, p);
  attn_bwd_dkv_tc_kernel<NA, BW, NBUF><<<grid, AT_THREADS, smem_kv, s>>>(mq128, mqb, mdb, p);
  return BFK_OK;
}

int attention_bwd_tc(const bfk_attention_args* a, cudaStream_t s) {
  if (const int rc = attention_bwd_v2(a, s); rc <= 0) return rc;
  AttnTcParams p = make_params(a);
  {
    const int chunks = a->H * (a->hd / 8);
    if (chunks > 1024) { set_error("attention_bwd_tc: H*hd/8 = %d exceeds 1024", chunks); return BFK_E_UNSUPPORTED; }
    const int rpb = std::max(1, 256 / chunks);
    const long long rows = (long long)a->B * a->N;
    attn_delta_tc_kernel<<<(unsigned)((rows + rpb - 1) / rpb), chunks * rpb, chunks * rpb * sizeof(float), s>>>(
        (const __nv_bfloat16*)a->out, a->ld_out, (const __nv_bfloat16*)a->dout, a->ld_dout, a->delta, a->B, a->N, a->H, a->hd, rpb);
  }
  int rc;
  if (a->hd <= 64) rc = bwd_block_width() == 64 ? launch_bwd<1, 64, 3>(a, p, s) : launch_bwd<1, 128, 2>(a, p, s);
  else rc = launch_bwd<2, 64, 3>(a, p, s);   // 128-wide blocks with two head-dim atoms do not fit in shared memory
  if (rc) return rc;
  BFK_CHECK_LAUNCH("attention_bwd_tc");
  return BFK_OK;
}

}  // namespace bfk
