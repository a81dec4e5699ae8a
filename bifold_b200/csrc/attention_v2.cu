// tcgen05 attention for SHORT sequences (N <= ~270 tokens: the SigLIP towers at 224x224 and the single-frame fusion transformer),
// where the generic flash kernels of attention_tc.cu are latency-bound: with 2-5 key blocks per CTA, TMEM allocation, barrier
// set-up, the first TMA round trip and the output epilogue dominate a CTA's life (profiles/r01_attention_ncu.txt: tensor pipe
// 11-18 % busy, 16-25 resident warps per SM).
//
// Both kernels here are PERSISTENT (one CTA per SM loops over (sample, head) items) and keep the whole K and V of an item
// resident in shared memory, so every operand is fetched once per item and the set-up cost is paid once per SM.
//
//   attn_fwd_v2 : 10 warps.  Warps 0-3 / 4-7 are two softmax GROUPS; each group owns one 128-row query tile of the item at a
//                 time (one thread = one query row = one TMEM lane) and ping-pongs with the other group: while one group
//                 exponentiates its S tile the tensor pipe computes the other group's S = Q K^T / O += P V.  Warp 8 issues every
//                 tcgen05.mma from one elected thread, warp 9 issues every TMA load; they run ahead through mbarriers, so the
//                 next item's Q and K land while the current softmax runs (K is released as soon as the last S of the item
//                 is done, V after the last P V).  The key range is cut into `nblk` blocks of `bw` <= 256 columns chosen so that
//                 two S tiles (+ two O tiles when nblk > 1) fit the 512 TMEM columns: N = 196 / 64 run with ONE block (plain
//                 softmax over the whole row, O aliases the S columns), N = 262 with two (online softmax).
//   attn_bwd_v2 : ONE merged kernel for dQ, dK and dV (and the row statistic delta = rowsum(dO * O), which used to be a launch
//                 of its own).  Q, K, V, dO of the item stay in shared memory; per (128 query rows x 128 keys) block the tensor
//                 pipe computes S and dP = dO V^T once, all 8 row warps turn them into P and dS (bf16, swizzled shared memory)
//                 and the three gradient products dQ_i += dS K_j, dV_j += P^T dO_i, dK_j += dS^T Q_i read those tiles as
//                 K-major / MN-major UMMA operands.  Accumulators: S, dP (128 columns each), dV_j, dK_j and one dQ per query
//                 tile — 512 TMEM columns at N = 196 / head_dim 64.  S / dP of block b+1 are issued before the gradient
//                 products of block b, so the exponentials of the next block overlap the tensor work of the current one.
//
// Query-row mask (bifold/models/transformer.py:33-35): a masked row has p == 1/N for every key and dS == 0.
#include "attention_common.cuh"

namespace bfk {

constexpr int A2_THREADS = 320;   // 8 row warps + MMA issuer + TMA loader

struct AttnV2Params {
  int B, N, H, hd;
  int ksteps, n_o;     // ceil(hd/16), ksteps*16
  int ntq;             // 128-row query tiles per item
  int nblk, bw;        // fwd: key blocks per query tile and their width (multiple of 16)
  int krows;           // rows of the resident K / V (bwd: also Q / dO) tiles = nbox * box_rows >= round16(N)
  int nbox, box_rows;  // TMA boxes per resident tile
  int patoms;          // fwd: 64-key atoms of one P tile
  int items;           // B * H
  float scale;
  const float* row_mask;
  float* lse;
  __nv_bfloat16* out; long long ld_out;
  const __nv_bfloat16* dout; long long ld_dout;
  __nv_bfloat16* dqkv; long long ld_dqkv;
};

__device__ __forceinline__ int round16(int x) { return (x + 15) & ~15; }

// =====================================================================================================================
// forward
// =====================================================================================================================
template <int NA>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn_fwd_v2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, AttnV2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  constexpr uint32_t QT = NA * AT_ATOM;                       // one group's Q tile (128 rows)
  const uint32_t KATOM = (uint32_t)p.krows * 128u, KT = NA * KATOM;
  const uint32_t PT = (uint32_t)p.patoms * AT_ATOM;           // one group's P tile (128 rows x bw keys, bf16)
  const uint32_t sQ = base, sK = sQ + 2 * QT, sV = sK + KT, sP = sV + KT, bars = sP + 2 * PT;
  enum { B_Q = 0, B_QFREE = 2, B_K = 4, B_KFREE = 5, B_V = 6, B_VFREE = 7, B_S = 8, B_P = 10, B_O = 12, B_OF = 14, NBAR = 16 };
  auto BAR = [&](int i) { return bars + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = bars + 8u * NBAR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (base & 1023u) __trap();
    for (int i = 0; i < NBAR; ++i) mbar_init(BAR(i), (i == B_P || i == B_P + 1 || i == B_OF || i == B_OF + 1) ? 128u : 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - base));
  // group g: S at column g*256; O behind it, or on top of it when the whole row is one block (S is dead once P is in smem)
  const uint32_t o_off = p.nblk == 1 ? 0u : (uint32_t)p.bw;

  if (warp == 9) {
    // ---------------------------------------------- TMA loader ----------------------------------------------
    if (elect_one()) {
      uint32_t qcnt[2] = {0, 0};
      int n = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
        const int b = item / p.H, h = item % p.H;
        auto load_kv = [&](uint32_t dst, int bar, int slot) {
          mbar_expect_tx(BAR(bar), (uint32_t)(NA * p.nbox * p.box_rows) * 128u);
          for (int a = 0; a < NA; ++a)
            for (int bx = 0; bx < p.nbox; ++bx)
              tma_load_4d(dst + a * KATOM + (uint32_t)(bx * p.box_rows) * 128u, &map_kv, BAR(bar), 64 * a, slot, bx * p.box_rows, b);
        };
        auto load_q = [&](int i) {
          const int g = i & 1;
          if (qcnt[g] > 0) mbar_wait(BAR(B_QFREE + g), (qcnt[g] - 1) & 1);   // the tile's last S product has read the buffer
          mbar_expect_tx(BAR(B_Q + g), QT);
          load_tile<NA>(sQ + g * QT, &map_q, BAR(B_Q + g), h, i * 128, b);
          ++qcnt[g];
        };
        if (n > 0) mbar_wait(BAR(B_KFREE), (n - 1) & 1);
        load_kv(sK, B_K, p.H + h);
        load_q(0);
        if (p.ntq > 1) load_q(1);
        if (n > 0) mbar_wait(BAR(B_VFREE), (n - 1) & 1);
        load_kv(sV, B_V, 2 * p.H + h);
        for (int i = 2; i < p.ntq; ++i) load_q(i);
      }
    }
  } else if (warp == 8) {
    // ---------------------------------------------- MMA issuer ----------------------------------------------
    if (elect_one()) {
      const uint32_t idesc_o = make_idesc(128, p.n_o, false, true);
      auto blk_cols = [&](int j) { return round16(min(p.bw, p.N - j * p.bw)); };
      auto issue_s = [&](int g, int j) {
        const uint32_t idesc = make_idesc(128, blk_cols(j), false, false);
        uint64_t da = kmajor_desc(sQ + g * QT, 0), db = kmajor_desc(sK + (uint32_t)(j * p.bw) * 128u, 0, KATOM);
        for (int k = 0; k < p.ksteps; ++k) {
          tc_mma_bf16(tmem + g * 256, da, db, idesc, k > 0);
          da += 2; db += 2;
          if ((k & 3) == 3) { da += (AT_ATOM >> 4) - 8; db += (KATOM >> 4) - 8; }
        }
      };
      auto issue_pv = [&](int g, int j) {
        const int nk = blk_cols(j) >> 4;
        uint64_t da = kmajor_desc(sP + g * PT, 0), db = mnmajor_desc(sV + (uint32_t)(j * p.bw) * 128u, 0, KATOM);
        for (int k = 0; k < nk; ++k) {
          tc_mma_bf16(tmem + g * 256 + o_off, da, db, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
          da += 2; db += 2048 >> 4;
          if ((k & 3) == 3) da += (AT_ATOM >> 4) - 8;
        }
      };
      uint32_t tcnt[2] = {0, 0}, pcnt[2] = {0, 0};
      int n = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
        mbar_wait(BAR(B_K), n & 1);
        bool v_ready = false;
        for (int r0 = 0; r0 < p.ntq; r0 += 2) {
          const int ng = min(2, p.ntq - r0);
          const bool last_round = r0 + 2 >= p.ntq;
          for (int g = 0; g < ng; ++g) {
            mbar_wait(BAR(B_Q + g), tcnt[g] & 1);
            if (tcnt[g] > 0) mbar_wait(BAR(B_OF + g), (tcnt[g] - 1) & 1);   // the group has read its previous O tile out of TMEM
            tc_fence_after();
            issue_s(g, 0);
            tc_commit(BAR(B_S + g));
            if (p.nblk == 1) {
              tc_commit(BAR(B_QFREE + g));
              if (last_round && g == ng - 1) tc_commit(BAR(B_KFREE));
            }
          }
          for (int j = 0; j < p.nblk; ++j) {
            for (int g = 0; g < ng; ++g) {
              mbar_wait(BAR(B_P + g), pcnt[g] & 1);     // P(j) is in shared memory and S(j) has been read
              ++pcnt[g];
              if (!v_ready) { mbar_wait(BAR(B_V), n & 1); v_ready = true; }
              tc_fence_after();
              issue_pv(g, j);
              if (j + 1 < p.nblk) {
                issue_s(g, j + 1);                       // in-order completion: bar_s(j+1) also says P V(j) is done
                tc_commit(BAR(B_S + g));
                if (j + 2 == p.nblk) {
                  tc_commit(BAR(B_QFREE + g));
                  if (last_round && g == ng - 1) tc_commit(BAR(B_KFREE));
                }
              } else {
                tc_commit(BAR(B_O + g));
                if (last_round && g == ng - 1) tc_commit(BAR(B_VFREE));
              }
            }
          }
          for (int g = 0; g < ng; ++g) ++tcnt[g];
        }
      }
    }
  } else {
    // ---------------------------------------------- softmax groups ----------------------------------------------
    const int g = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t tS = tmem + g * 256 + lane_off, tO = tS + o_off;
    const uint32_t myP = sP + g * PT;
    uint32_t scnt = 0, ocnt = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int b = item / p.H, h = item % p.H;
      for (int i = g; i < p.ntq; i += 2) {
        const int q = i * 128 + r;
        const bool quarter_live = i * 128 + quarter * 32 < p.N;       // warp-uniform: any real query row in this quarter?
        const bool masked = p.row_mask && q < p.N && p.row_mask[(long long)b * p.N + q] == 0.f;
        const float sl2 = masked ? 0.f : p.scale * LOG2E;            // masked row: every score equal => p == 1 for every key
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < p.nblk; ++j) {
          const int nvalid = min(p.bw, p.N - j * p.bw), nc = round16(nvalid);
          mbar_wait(BAR(B_S + g), scnt & 1);
          ++scnt;
          tc_fence_after();
          if (quarter_live) {
            // pass 1: row maximum over the real keys
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < nc; c += 32) {
              uint32_t sreg[32];
              const int lim = nvalid - c;
              if (nc - c >= 32) {
                tc_ld32(tS + c, sreg);
                tc_wait_ld();
                if (lim >= 32) {
#pragma unroll
                  for (int t = 0; t < 32; ++t) mx = fmaxf(mx, __uint_as_float(sreg[t]));
                } else {
#pragma unroll
                  for (int t = 0; t < 32; ++t) if (t < lim) mx = fmaxf(mx, __uint_as_float(sreg[t]));
                }
              } else {
                uint32_t (&s16)[16] = *reinterpret_cast<uint32_t(*)[16]>(sreg);
                tc_ld16(tS + c, s16);
                tc_wait_ld();
#pragma unroll
                for (int t = 0; t < 16; ++t) if (t < lim) mx = fmaxf(mx, __uint_as_float(s16[t]));
              }
            }
            const float m_new = masked ? 0.f : fmaxf(m_run, mx * sl2);
            const float alpha = ex2f(m_run - m_new);
            // pass 2: p = 2^(s*scale*log2e - m) -> bf16 into the swizzled P tile, fp32 row sum
            float rowsum = 0.f;
            auto emit8 = [&](const uint32_t* s8, int col) {        // 8 consecutive keys starting at column `col`
              float v[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) v[t] = ex2f(fmaf(__uint_as_float(s8[t]), sl2, -m_new));
              if (col + 8 > nvalid) {
#pragma unroll
                for (int t = 0; t < 8; ++t) if (col + t >= nvalid) v[t] = 0.f;
              }
#pragma unroll
              for (int t = 0; t < 8; ++t) rowsum += v[t];
              store_p_chunk(myP, r, col >> 3, v);
            };
#pragma unroll 1
            for (int c = 0; c < nc; c += 32) {
              uint32_t sreg[32];
              if (nc - c >= 32) {
                tc_ld32(tS + c, sreg);
                tc_wait_ld();
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) emit8(sreg + 8 * g8, c + 8 * g8);
              } else {
                uint32_t (&s16)[16] = *reinterpret_cast<uint32_t(*)[16]>(sreg);
                tc_ld16(tS + c, s16);
                tc_wait_ld();
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) emit8(sreg + 8 * g8, c + 8 * g8);
              }
            }
            l_run = l_run * alpha + rowsum;
            m_run = m_new;
            if (j > 0 && !__all_sync(0xffffffffu, alpha == 1.f)) {     // bar_s(j) implies P V(j-1) is complete
              for (int c = 0; c < p.n_o / 16; ++c) {
                uint32_t o[16];
                tc_ld16(tO + c * 16, o);
                tc_wait_ld();
#pragma unroll
                for (int t = 0; t < 16; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
                tc_st16(tO + c * 16, o);
              }
              tc_wait_st();
            }
          }
          fence_async_smem();
          tc_fence_before();
          mbar_arrive(BAR(B_P + g));
        }
        mbar_wait(BAR(B_O + g), ocnt & 1);
        ++ocnt;
        tc_fence_after();
        if (quarter_live) {
          const float inv = 1.f / l_run;
          __nv_bfloat16* orow = p.out + ((long long)b * p.N + q) * p.ld_out + h * p.hd;
          for (int c = 0; c < p.n_o / 16; ++c) {
            uint32_t o[16];
            tc_ld16(tO + c * 16, o);
            tc_wait_ld();
            if (q < p.N) store_row16(orow, c * 16, p.hd, o, inv);
          }
          if (q < p.N && p.lse) p.lse[((long long)b * p.H + h) * p.N + q] = m_run * LN2 + logf(l_run);
        }
        tc_fence_before();
        mbar_arrive(BAR(B_OF + g));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// =====================================================================================================================
// backward (dQ, dK, dV and delta in one kernel)
// =====================================================================================================================
template <int NA>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn_bwd_v2_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do, AttnV2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t KATOM = (uint32_t)p.krows * 128u, KT = NA * KATOM;
  constexpr uint32_t BT = 2 * AT_ATOM;                        // P / dS block: 128 query rows x 128 keys (bf16)
  const uint32_t sQ = base, sdO = sQ + KT, sK = sdO + KT, sV = sK + KT, sP = sV + KT, sdS = sP + BT, stat = sdS + BT;
  const uint32_t stat_bytes = (uint32_t)p.ntq * 128u * 12u;   // float2 {lse*log2e, delta} + int flag per query row
  const uint32_t bars = (stat + stat_bytes + 15u) & ~15u;
  enum { B_LD0 = 0, B_LD1 = 1, B_FREE = 2, B_S = 3, B_P = 4, B_G = 5, NBAR = 6 };
  auto BAR = [&](int i) { return bars + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = bars + 8u * NBAR;
  float2* s_stat = reinterpret_cast<float2*>(smem_raw + (stat - base));
  int* s_flag = reinterpret_cast<int*>(smem_raw + (stat - base) + (size_t)p.ntq * 128 * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.N + 127) / 128;              // 128-key blocks (the M tile of dV / dK)
  const int nb = nkb * p.ntq;                     // (query tile, key block) blocks per item, key block outermost

  if (threadIdx.x == 0) {
    if (base & 1023u) __trap();
    for (int i = 0; i < NBAR; ++i) mbar_init(BAR(i), i == B_P ? 256u : 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - base));
  const uint32_t tS = tmem, tdP = tmem + 128, tdV = tmem + 256, tdK = tdV + p.n_o, tdQ = tdK + p.n_o;   // dQ tile i at tdQ + i*n_o

  if (warp == 9) {
    if (elect_one()) {
      int n = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
        const int b = item / p.H, h = item % p.H;
        if (n > 0) mbar_wait(BAR(B_FREE), (n - 1) & 1);       // every product of the previous item has read its operands
        auto load = [&](uint32_t dst, const CUtensorMap* map, int bar, int slot) {
          for (int a = 0; a < NA; ++a)
            for (int bx = 0; bx < p.nbox; ++bx)
              tma_load_4d(dst + a * KATOM + (uint32_t)(bx * p.box_rows) * 128u, map, BAR(bar), 64 * a, slot, bx * p.box_rows, b);
        };
        const uint32_t bytes = (uint32_t)(2 * NA * p.nbox * p.box_rows) * 128u;
        mbar_expect_tx(BAR(B_LD0), bytes);
        load(sQ, &map_qkv, B_LD0, h);
        load(sK, &map_qkv, B_LD0, p.H + h);
        mbar_expect_tx(BAR(B_LD1), bytes);
        load(sdO, &map_do, B_LD1, h);
        load(sV, &map_qkv, B_LD1, 2 * p.H + h);
      }
    }
  } else if (warp == 8) {
    if (elect_one()) {
      const uint32_t idesc_q = make_idesc(128, p.n_o, false, true);    // dQ: A = dS (K-major over keys), B = K (MN-major)
      const uint32_t idesc_g = make_idesc(128, p.n_o, true, true);     // dV / dK: A = P^T / dS^T (MN-major), B = dO / Q (MN-major)
      auto key_cols = [&](int j) { return round16(min(128, p.N - j * 128)); };
      auto issue_sdp = [&](int bidx) {
        const int j = bidx / p.ntq, i = bidx % p.ntq;
        const uint32_t idesc = make_idesc(128, key_cols(j), false, false);
        uint64_t da = kmajor_desc(sQ + (uint32_t)i * AT_ATOM, 0, KATOM), db = kmajor_desc(sK + (uint32_t)j * AT_ATOM, 0, KATOM);
        for (int k = 0; k < p.ksteps; ++k) {
          tc_mma_bf16(tS, da, db, idesc, k > 0);
          da += 2; db += 2;
          if ((k & 3) == 3) { da += (KATOM >> 4) - 8; db += (KATOM >> 4) - 8; }
        }
        da = kmajor_desc(sdO + (uint32_t)i * AT_ATOM, 0, KATOM); db = kmajor_desc(sV + (uint32_t)j * AT_ATOM, 0, KATOM);
        for (int k = 0; k < p.ksteps; ++k) {
          tc_mma_bf16(tdP, da, db, idesc, k > 0);
          da += 2; db += 2;
          if ((k & 3) == 3) { da += (KATOM >> 4) - 8; db += (KATOM >> 4) - 8; }
        }
        tc_commit(BAR(B_S));
      };
      uint32_t pc = 0;
      int n = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
        mbar_wait(BAR(B_LD0), n & 1);
        mbar_wait(BAR(B_LD1), n & 1);
        tc_fence_after();
        issue_sdp(0);
        for (int bidx = 0; bidx < nb; ++bidx) {
          const int j = bidx / p.ntq, i = bidx % p.ntq;
          mbar_wait(BAR(B_P), pc & 1);          // P / dS of this block are in shared memory, S / dP have been read
          ++pc;
          tc_fence_after();
          if (bidx + 1 < nb) issue_sdp(bidx + 1);    // first: the row warps wait for these; the gradient products below only feed accumulators
          const int nkk = key_cols(j) >> 4;                                   // K steps over the keys of the block (dQ)
          const int nkq = round16(min(128, p.N - i * 128)) >> 4;              // K steps over the live query rows (dV, dK)
          uint64_t da = kmajor_desc(sdS, 0), db = mnmajor_desc(sK + (uint32_t)j * AT_ATOM, 0, KATOM);
          for (int k = 0; k < nkk; ++k) {
            tc_mma_bf16(tdQ + i * p.n_o, da, db, idesc_q, (j > 0 || k > 0) ? 1u : 0u);
            da += 2; db += 2048 >> 4;
            if ((k & 3) == 3) da += (AT_ATOM >> 4) - 8;
          }
          da = mnmajor_desc(sP, 0, AT_ATOM); db = mnmajor_desc(sdO + (uint32_t)i * AT_ATOM, 0, KATOM);
          for (int k = 0; k < nkq; ++k) {
            tc_mma_bf16(tdV, da, db, idesc_g, (i > 0 || k > 0) ? 1u : 0u);
            da += 2048 >> 4; db += 2048 >> 4;
          }
          da = mnmajor_desc(sdS, 0, AT_ATOM); db = mnmajor_desc(sQ + (uint32_t)i * AT_ATOM, 0, KATOM);
          for (int k = 0; k < nkq; ++k) {
            tc_mma_bf16(tdK, da, db, idesc_g, (i > 0 || k > 0) ? 1u : 0u);
            da += 2048 >> 4; db += 2048 >> 4;
          }
          tc_commit(BAR(B_G));
        }
        tc_commit(BAR(B_FREE));
      }
    }
  } else {
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const int tid = threadIdx.x;                                     // 0..255
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    const float log2n = log2f((float)p.N);
    const long long HD = (long long)p.H * p.hd;
    uint32_t sc = 0, gc = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int b = item / p.H, h = item % p.H;
      // ---- per-row statistics: {lse * log2e, delta = rowsum(dO * O)}; flag 1 = live, 2 = masked query row (p = 1/N, dS = 0), 0 = beyond N ----
      for (int row = tid; row < p.ntq * 128; row += 256) {
        float2 st = make_float2(INFINITY, 0.f);
        int fl = 0;
        if (row < p.N) {
          if (p.row_mask && p.row_mask[(long long)b * p.N + row] == 0.f) {
            st = make_float2(log2n, 0.f);
            fl = 2;
          } else {
            const __nv_bfloat16* orow = p.out + ((long long)b * p.N + row) * p.ld_out + h * p.hd;
            const __nv_bfloat16* drow = p.dout + ((long long)b * p.N + row) * p.ld_dout + h * p.hd;
            float d = 0.f;
            for (int c = 0; c < p.hd; c += 8) {
              const uint4 a = *reinterpret_cast<const uint4*>(orow + c), e = *reinterpret_cast<const uint4*>(drow + c);
              const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
              const __nv_bfloat162* ep = reinterpret_cast<const __nv_bfloat162*>(&e);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 x = __bfloat1622float2(ap[t]), y = __bfloat1622float2(ep[t]);
                d += x.x * y.x + x.y * y.y;
              }
            }
            st = make_float2(p.lse[((long long)b * p.H + h) * p.N + row] * LOG2E, d);
            fl = 1;
          }
        }
        s_stat[row] = st;
        s_flag[row] = fl;
      }
      named_bar_sync(1, 256);

      auto epilogue_dkv = [&](int j) {       // half 0 stores dV, half 1 stores dK of key block j
        const int kv = j * 128 + r;
        if (j * 128 + quarter * 32 < p.N) {
          __nv_bfloat16* krow = p.dqkv + ((long long)b * p.N + kv) * p.ld_dqkv + HD + h * p.hd;
          __nv_bfloat16* vrow = krow + HD;
          for (int c = 0; c < p.n_o / 16; ++c) {
            uint32_t o[16];
            tc_ld16((half == 0 ? tdV : tdK) + lane_off + c * 16, o);
            tc_wait_ld();
            if (kv < p.N) store_row16(half == 0 ? vrow : krow, c * 16, p.hd, o, half == 0 ? 1.f : p.scale);
          }
        }
      };

      for (int bidx = 0; bidx < nb; ++bidx) {
        const int j = bidx / p.ntq, i = bidx % p.ntq;
        const int nc = round16(min(128, p.N - j * 128));
        mbar_wait(BAR(B_S), sc & 1);
        ++sc;
        if (bidx > 0) {                       // gradient products of the previous block are done: P / dS may be overwritten,
          mbar_wait(BAR(B_G), gc & 1);        // and dV / dK of a finished key block are final
          ++gc;
        }
        tc_fence_after();
        if (bidx > 0 && i == 0) epilogue_dkv(j - 1);
        if (i * 128 + quarter * 32 < p.N) {
          const float2 st = s_stat[i * 128 + r];
          const int fl = s_flag[i * 128 + r];
          const float sl2_eff = fl == 1 ? sl2 : 0.f;
#pragma unroll 1
          for (int c = half * 32; c < nc; c += 64) {
            uint32_t sreg[32], dreg[32];
            const int w = min(32, nc - c);
            if (w == 32) {
              tc_ld32(tS + lane_off + c, sreg);
              tc_ld32(tdP + lane_off + c, dreg);
            } else {
              tc_ld16(tS + lane_off + c, *reinterpret_cast<uint32_t(*)[16]>(sreg));
              tc_ld16(tdP + lane_off + c, *reinterpret_cast<uint32_t(*)[16]>(dreg));
            }
            tc_wait_ld();
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              if (g8 * 8 < w) {
                float pv[8], dv[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                  const float pr = fl == 0 ? 0.f : ex2f(fmaf(__uint_as_float(sreg[g8 * 8 + t]), sl2_eff, -st.x));
                  pv[t] = pr;
                  dv[t] = fl == 1 ? pr * (__uint_as_float(dreg[g8 * 8 + t]) - st.y) : 0.f;
                }
                store_p_chunk(sP, r, (c >> 3) + g8, pv);
                store_p_chunk(sdS, r, (c >> 3) + g8, dv);
              }
            }
          }
        }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(BAR(B_P));
      }
      mbar_wait(BAR(B_G), gc & 1);
      ++gc;
      tc_fence_after();
      epilogue_dkv(nkb - 1);
      for (int i = 0; i < p.ntq; ++i) {
        const int q = i * 128 + r;
        if (i * 128 + quarter * 32 < p.N) {
          __nv_bfloat16* drow = p.dqkv + ((long long)b * p.N + q) * p.ld_dqkv + h * p.hd;
          for (int c = half; c < p.n_o / 16; c += 2) {
            uint32_t o[16];
            tc_ld16(tdQ + i * p.n_o + lane_off + c * 16, o);
            tc_wait_ld();
            if (q < p.N) store_row16(drow, c * 16, p.hd, o, p.scale);
          }
        }
      }
      tc_fence_before();
      named_bar_sync(1, 256);       // the statistics of this item have been read by everybody before the next item overwrites them
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// ---- host ---------------------------------------------------------------------------------------------
constexpr size_t A2_SMEM_MAX = 232448;     // 227 KB per CTA

static bool v2_enabled() {                 // BFK_ATTN_V2=0 keeps the generic flash kernels (A/B measurement)
  static int v = -1;
  if (v < 0) { const char* e = getenv("BFK_ATTN_V2"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v == 1;
}

static void plan_common(const bfk_attention_args* a, AttnV2Params& p) {
  p.B = a->B; p.N = a->N; p.H = a->H; p.hd = a->hd;
  p.ksteps = (a->hd + 15) / 16;
  p.n_o = p.ksteps * 16;
  p.ntq = (a->N + 127) / 128;
  const int npk = (a->N + 15) & ~15;
  p.nbox = (npk + 255) / 256;
  p.box_rows = (((npk + p.nbox - 1) / p.nbox) + 7) & ~7;
  p.krows = p.nbox * p.box_rows;
  p.items = a->B * a->H;
  p.scale = a->scale;
  p.row_mask = a->row_mask;
  p.lse = a->lse;
  p.out = (__nv_bfloat16*)a->out; p.ld_out = a->ld_out;
  p.dout = (const __nv_bfloat16*)a->dout; p.ld_dout = a->ld_dout;
  p.dqkv = (__nv_bfloat16*)a->dqkv; p.ld_dqkv = a->ld_dqkv;
  p.nblk = 1; p.bw = 0; p.patoms = 0;
}

// forward plan: smallest number of key blocks whose S (+ O) tiles fit two groups into 512 TMEM columns
static bool plan_fwd(const bfk_attention_args* a, AttnV2Params& p, size_t& smem) {
  if (!v2_enabled() || a->N > 1024) return false;
  plan_common(a, p);
  const int NA = a->hd <= 64 ? 1 : 2;
  bool ok = false;
  for (int nblk = 1; nblk <= 8 && !ok; ++nblk) {
    const int bw = (((a->N + nblk - 1) / nblk) + 15) & ~15;
    if (bw > 256 || (nblk - 1) * bw >= a->N) continue;
    const int cols = nblk == 1 ? (bw > p.n_o ? bw : p.n_o) : bw + p.n_o;
    if (cols <= 256) { p.nblk = nblk; p.bw = bw; ok = true; }
  }
  if (!ok) return false;
  p.patoms = (p.bw + 63) / 64;
  smem = (size_t)2 * NA * AT_ATOM + (size_t)2 * NA * p.krows * 128 + (size_t)2 * p.patoms * AT_ATOM + 8 * 16 + 16;
  return smem <= A2_SMEM_MAX;
}

static bool plan_bwd(const bfk_attention_args* a, AttnV2Params& p, size_t& smem) {
  if (!v2_enabled() || a->N > 1024) return false;
  plan_common(a, p);
  const int NA = a->hd <= 64 ? 1 : 2;
  if (256 + (2 + p.ntq) * p.n_o > 512) return false;                   // S, dP, dV, dK and one dQ per query tile in TMEM
  smem = (size_t)4 * NA * p.krows * 128 + (size_t)4 * AT_ATOM + (size_t)p.ntq * 128 * 12 + 16 + 8 * 6 + 16;
  return smem <= A2_SMEM_MAX;
}

template <typename Kern>
static int set_smem(Kern k, size_t smem, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("%s: cannot set smem %zu: %s", name, smem, cudaGetErrorString(e)); return BFK_E_CUDA; }
  return BFK_OK;
}

// returns 1 when the problem is not eligible (caller falls back to the generic kernels), 0 on success, < 0 on error
int attention_fwd_v2(const bfk_attention_args* a, cudaStream_t s) {
  AttnV2Params p;
  size_t smem = 0;
  if (!plan_fwd(a, p, smem)) return 1;
  CUtensorMap mq, mkv;
  if (!make_map4(&mq, a->qkv, a->hd, 3 * a->H, a->N, a->B, a->ld_qkv, 128)) return BFK_E_CUDA;
  if (!make_map4(&mkv, a->qkv, a->hd, 3 * a->H, a->N, a->B, a->ld_qkv, p.box_rows)) return BFK_E_CUDA;
  const int grid = p.items < sm_count() ? p.items : sm_count();
  if (a->hd <= 64) {
    if (int rc = set_smem(attn_fwd_v2_kernel<1>, smem, "attn_fwd_v2")) return rc;
    attn_fwd_v2_kernel<1><<<grid, A2_THREADS, smem, s>>>(mq, mkv, p);
  } else {
    if (int rc = set_smem(attn_fwd_v2_kernel<2>, smem, "attn_fwd_v2")) return rc;
    attn_fwd_v2_kernel<2><<<grid, A2_THREADS, smem, s>>>(mq, mkv, p);
  }
  BFK_CHECK_LAUNCH("attention_fwd_v2");
  return BFK_OK;
}

int attention_bwd_v2(const bfk_attention_args* a, cudaStream_t s) {
  AttnV2Params p;
  size_t smem = 0;
  if (!plan_bwd(a, p, smem)) return 1;
  CUtensorMap mqkv, mdo;
  if (!make_map4(&mqkv, a->qkv, a->hd, 3 * a->H, a->N, a->B, a->ld_qkv, p.box_rows)) return BFK_E_CUDA;
  if (!make_map4(&mdo, a->dout, a->hd, a->H, a->N, a->B, a->ld_dout, p.box_rows)) return BFK_E_CUDA;
  const int grid = p.items < sm_count() ? p.items : sm_count();
  if (a->hd <= 64) {
    if (int rc = set_smem(attn_bwd_v2_kernel<1>, smem, "attn_bwd_v2")) return rc;
    attn_bwd_v2_kernel<1><<<grid, A2_THREADS, smem, s>>>(mqkv, mdo, p);
  } else {
    if (int rc = set_smem(attn_bwd_v2_kernel<2>, smem, "attn_bwd_v2")) return rc;
    attn_bwd_v2_kernel<2><<<grid, A2_THREADS, smem, s>>>(mqkv, mdo, p);
  }
  BFK_CHECK_LAUNCH("attention_bwd_v2");
  return BFK_OK;
}

}  // namespace bfk
