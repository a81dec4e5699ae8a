// tcgen05 attention for SHORT sequences (N <= ~270 tokens: the SigLIP towers at 224x224 and the single-frame fusion transformer),
// where the generic flash kernels of attention_tc.cu are latency-bound: with 2-5 key blocks per CTA, TMEM allocation, barrier
// set-up, the first TMA round trip and the output epilogue dominate a CTA's life (profiles/r01_attention_ncu.txt: tensor pipe
// 11-18 % busy, 16-25 resident warps per SM).
//
// Both kernels here are PERSISTENT (one CTA per SM loops over (sample, head) items) and keep the whole K and V of an item
// resident in shared memory, so every operand is fetched once per item and the set-up cost is paid once per SM.
//
//   attn_fwd_v2 : 10 warps: 8 row warps (two per TMEM lane quarter, each owning half of the key columns of its 32 query rows),
//                 warp 8 issues every tcgen05.mma from one elected thread, warp 9 every TMA load.  The work of an item is a
//                 sequence of STEPS (query tile, key block); S of step s+1 is computed into the second TMEM slot while the row
//                 warps work on step s, P V(s) runs while they work on step s+1, and the output epilogue of a tile runs one step
//                 late — in the steady state no warp waits for the tensor pipe (round-2 profile of the first version: 47 % of the
//                 row warps' samples were mbarrier waits).  A row warp reads its half row block from TMEM ONCE into registers
//                 (maximum and exponentials both come from registers), exchanges the row maximum with its sibling warp through
//                 shared memory and writes bf16 P into the swizzled tile the P V product reads.  The key range is cut into `nblk`
//                 blocks of `bw` <= 256 columns so that two S slots (+ two O tiles when nblk > 1) fit the 512 TMEM columns:
//                 N = 196 / 64 run with ONE block (plain softmax over the whole row, O reuses the S columns), N = 262 with two
//                 (online softmax, O rescaled in TMEM).  K is released as soon as the last S of the item is done, V after the last
//                 P V, so the next item's operands land while the current softmax runs.
//   attn_bwd_v2 : ONE merged kernel for dQ, dK and dV (and the row statistic delta = rowsum(dO * O), which used to be a launch
//                 of its own).  Q, K, V, dO of the item stay in shared memory; per (128 query rows x 128 keys) block the tensor
//                 pipe computes S and dP = dO V^T once, all 8 row warps turn them into P and dS (bf16, swizzled shared memory)
//                 and the three gradient products dQ_i += dS K_j, dV_j += P^T dO_i, dK_j += dS^T Q_i read those tiles as
//                 K-major / MN-major UMMA operands.  Accumulators: S, dP (128 columns each), dV_j, dK_j and one dQ per query
//                 tile — 512 TMEM columns at N = 196 / head_dim 64.  The row warps pull S / dP into registers, release the TMEM
//                 columns at once (S / dP of block b+1 are computed while block b is exponentiated), keep P / dS packed in
//                 registers and store them only when the gradient products of block b-1 have released the shared tiles; the
//                 statistics of the CTA's next item are fetched while the last products of the current one run.
//
// Query-row mask (bifold/models/transformer.py:33-35): a masked row has p == 1/N for every key and dS == 0.
#include "attention_common.cuh"

namespace bfk {

constexpr int A2_THREADS = 384;   // three warpgroups: 8 row warps | MMA issuer, TMA loader, two idle warps
// Register budget: the register file is per SM sub-partition (16 K registers, three of the twelve warps each), so a uniform
// allocation stops at 168 registers per thread.  The auxiliary warpgroup hands its registers back right after set-up
// (setmaxnreg.dec) and the two row warpgroups grow to 224 (setmaxnreg.inc): 32 * (2 * 224 + 56) = 16 128 <= 16 384.
constexpr int A2_REGS_ROW = 224, A2_REGS_AUX = 56;
__device__ __forceinline__ void a2_regs_aux() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(A2_REGS_AUX)); }
__device__ __forceinline__ void a2_regs_row() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(A2_REGS_ROW)); }

struct AttnV2Params {
  int B, N, H, hd;
  int ksteps, n_o;     // ceil(hd/16), ksteps*16
  int ntq;             // 128-row query tiles per item
  int nblk, bw;        // fwd: key blocks per query tile and their width (multiple of 16)
  int krows;           // rows of the resident K / V (bwd: also Q / dO) tiles = nbox * box_rows >= round16(N)
  int nbox, box_rows;  // TMA boxes per resident tile
  int patoms;          // fwd: 64-key atoms of one P tile
  int items;           // B * H
  int dbg;             // BFK_ATTN_DBG ablation bits (measurement only; results are garbage)
  float scale;
  const float* row_mask;
  float* lse;
  __nv_bfloat16* out; long long ld_out;
  const __nv_bfloat16* dout; long long ld_dout;
  __nv_bfloat16* dqkv; long long ld_dqkv;
};

__device__ __forceinline__ int round16(int x) { return (x + 15) & ~15; }
// BFK_ATTN_DBG & 64: CTA 0 records SM-clock timestamps of its pipeline events into the lse buffer (measurement only)
struct A2Trace {
  uint32_t* buf; int n, cap;
  __device__ __forceinline__ void init(float* lse, int off, int cap_, bool on) { buf = on ? reinterpret_cast<uint32_t*>(lse) + off : nullptr; n = 0; cap = cap_; }
  __device__ __forceinline__ void mark(uint32_t tag) {
#ifdef BFK_A2_TRACE
    if (buf && n + 2 <= cap) { uint32_t c; asm volatile("mov.u32 %0, %%clock;" : "=r"(c)); buf[n++] = tag; buf[n++] = c; }
#endif
  }
};

// =====================================================================================================================
// forward
// =====================================================================================================================
// TMEM -> register loads of one warp's share of an S block: NC8 groups of 8 columns, group g predicated on g < n8 (warp-uniform)
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

template <int W> __device__ __forceinline__ void ld_o(uint32_t taddr, uint32_t (&o)[W]) {
  if constexpr (W == 16) tc_ld16(taddr, o); else if constexpr (W == 32) tc_ld32(taddr, o); else tc_ld64(taddr, o);
}
template <int W> __device__ __forceinline__ void st_o(uint32_t taddr, const uint32_t (&o)[W]) {
  if constexpr (W == 16) tc_st16(taddr, o); else if constexpr (W == 32) tc_st32(taddr, o); else tc_st64(taddr, o);
}

// LDW: width (32 / 64 / 128 columns) of the single tcgen05.ld that brings one warp's share (bw / 2 <= LDW columns) of an S block
// into registers.
// A "step" is one (query tile, key block) pair of an item; steps run back to back through the 8 row warps:
//   S(step)   lives in TMEM slot step & 1  (two slots: the tensor pipe computes S(step + 1) while the warps work on S(step)),
//   P(step)   in shared-memory slot step & 1,
//   O(tile)   in TMEM behind / on top of the S slot of parity tile & 1.
// Each warp reads its half of the S row block from TMEM ONCE into registers (the TMEM read port, 16 fp32 per clock and SM, is
// as scarce as the MUFU), takes the row maximum there, exchanges it with the sibling warp of its lane quarter through shared
// memory, exponentiates in place and writes bf16 P.  The epilogue of a tile runs one step late, so its P V product has long
// finished; nothing in the steady state waits for the tensor pipe.
template <int NA, int LDW, int HD16>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn_fwd_v2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, AttnV2Params p) {
  pdl_trigger();   // the GEMM behind this kernel may set itself up during its tail (common.cuh)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  constexpr uint32_t QT = NA * AT_ATOM;                       // one Q tile (128 rows)
  const uint32_t KATOM = (uint32_t)p.krows * 128u, KT = NA * KATOM;
  const uint32_t PT = (uint32_t)p.patoms * AT_ATOM;           // one P tile (128 rows x bw keys, bf16)
  const uint32_t sQ = base, sK = sQ + 2 * QT, sV = sK + KT, sP = sV + KT, xch_off = sP + 2 * PT, bars = xch_off + 4 * 1024;
  enum { B_Q = 0, B_QFREE = 2, B_K = 4, B_KFREE = 5, B_V = 6, B_VFREE = 7, B_S = 8, B_P = 10, B_O = 12, B_OF = 14, NBAR = 16 };
  auto BAR = [&](int i) { return bars + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = bars + 8u * NBAR;
  float* xch = reinterpret_cast<float*>(smem_raw + (xch_off - base));   // [2 step parities][2 halves][128] row maxima, then [2 tile parities][2 halves][128] row sums
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ns = p.ntq * p.nblk;                              // steps per item
  // O columns handled by the two warps of a lane quarter: [0, OS0) and [OS0, 16 * HD16); OLx = width of the one load that covers them
  constexpr int OS0 = ((HD16 + 1) / 2) * 16, OS1 = HD16 * 16 - OS0;
  constexpr int OL0 = OS0 <= 16 ? 16 : (OS0 <= 32 ? 32 : 64), OL1 = OS1 <= 16 ? 16 : (OS1 <= 32 ? 32 : 64);

  if (threadIdx.x == 0) {
    if (base & 1023u) __trap();
    for (int i = 0; i < NBAR; ++i) mbar_init(BAR(i), (i == B_P || i == B_P + 1 || i == B_OF || i == B_OF + 1) ? 256u : 1u);
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
  const uint32_t o_off = p.nblk == 1 ? 0u : (uint32_t)p.bw;   // one block per row: O reuses the columns of its (dead) S tile

  if (warp >= 8) {
   a2_regs_aux();        // whole auxiliary warpgroup (warps 8..11): give registers back to the row warps
   if (warp == 9) {
    // ---------------------------------------------- TMA loader ----------------------------------------------
    if (elect_one()) {
      uint32_t qcnt[2] = {0, 0};
      uint32_t gt = 0;                                         // global tile counter: Q buffer = gt & 1
      int n = 0;
      A2Trace tr; tr.init(p.lse, 8192, 2048, (p.dbg & 64) && blockIdx.x == 0);
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
        const int b = item / p.H, h = item % p.H;
        tr.mark(100 + n);
        auto load_kv = [&](uint32_t dst, int bar, int slot) {
          mbar_expect_tx(BAR(bar), (uint32_t)(NA * p.nbox * p.box_rows) * 128u);
          for (int a = 0; a < NA; ++a)
            for (int bx = 0; bx < p.nbox; ++bx)
              tma_load_4d(dst + a * KATOM + (uint32_t)(bx * p.box_rows) * 128u, &map_kv, BAR(bar), 64 * a, slot, bx * p.box_rows, b);
        };
        auto load_q = [&](int i) {
          const int qb = (gt + i) & 1;
          if (qcnt[qb] > 0) mbar_wait(BAR(B_QFREE + qb), (qcnt[qb] - 1) & 1);   // the last S product of the buffer's previous tile is done
          mbar_expect_tx(BAR(B_Q + qb), QT);
          load_tile<NA>(sQ + qb * QT, &map_q, BAR(B_Q + qb), h, i * 128, b);
          ++qcnt[qb];
        };
        if (n > 0) mbar_wait(BAR(B_KFREE), (n - 1) & 1);
        tr.mark(1);
        load_kv(sK, B_K, p.H + h);
        load_q(0);
        if (p.ntq > 1) load_q(1);
        tr.mark(2);
        if (n > 0) mbar_wait(BAR(B_VFREE), (n - 1) & 1);
        tr.mark(3);
        load_kv(sV, B_V, 2 * p.H + h);
        for (int i = 2; i < p.ntq; ++i) load_q(i);
        gt += p.ntq;
      }
    }
  } else if (warp == 8) {
    // ---------------------------------------------- MMA issuer ----------------------------------------------
    if (elect_one()) {
      const uint32_t idesc_o = make_idesc(128, p.n_o, false, true);
      auto blk_cols = [&](int j) { return round16(min(p.bw, p.N - j * p.bw)); };
      uint32_t qwait[2] = {0, 0};      // tiles whose Q buffer this thread has waited for, per buffer
      uint32_t pwait[2] = {0, 0};      // bar_p phases consumed, per S slot
      uint32_t ofwait[2] = {0, 0};     // bar_of phases consumed, per tile parity
      uint32_t tiles_on[2] = {0, 0};   // tiles started on each tile parity (an O buffer is busy until its epilogue has read it)
      uint32_t gs = 0, gt = 0;         // global step / tile counters of this CTA
      int n = 0;
      A2Trace tr; tr.init(p.lse, 4096, 4096, (p.dbg & 64) && blockIdx.x == 0);
      // S(step s of the current item) into TMEM slot (gs0 + s) & 1
      auto issue_s = [&](int s, uint32_t gs0, uint32_t gt0, bool last_of_item) {
        const int i = s / p.nblk, j = s % p.nblk;
        const uint32_t slot = (gs0 + s) & 1u, tp = (gt0 + i) & 1u;
        tr.mark(10);
        if (j == 0) {                                            // first block of a tile: its Q tile must have landed
          mbar_wait(BAR(B_Q + tp), qwait[tp] & 1);
          ++qwait[tp];
        }
        tr.mark(11);
        if (p.nblk == 1 && tiles_on[tp] > ofwait[tp]) {          // O of the tile that last used this slot has not been released yet
          mbar_wait(BAR(B_OF + tp), ofwait[tp] & 1);
          ++ofwait[tp];
        }
        tr.mark(12);
        tc_fence_after();
        const uint32_t idesc = make_idesc(128, blk_cols(j), false, false);
        uint64_t da = kmajor_desc(sQ + tp * QT, 0), db = kmajor_desc(sK + (uint32_t)(j * p.bw) * 128u, 0, KATOM);
        for (int k = 0; k < p.ksteps; ++k) {
          tc_mma_bf16(tmem + slot * 256, da, db, idesc, k > 0);
          da += 2; db += 2;
          if ((k & 3) == 3) { da += (AT_ATOM >> 4) - 8; db += (KATOM >> 4) - 8; }
        }
        tc_commit(BAR(B_S + slot));
        if (j == p.nblk - 1) tc_commit(BAR(B_QFREE + tp));
        if (last_of_item) tc_commit(BAR(B_KFREE));
        if (j == 0) ++tiles_on[tp];
        tr.mark(13);
      };
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
        tr.mark(200 + n);
        mbar_wait(BAR(B_K), n & 1);
        tr.mark(20);
        issue_s(0, gs, gt, ns == 1);
        if (ns > 1) issue_s(1, gs, gt, ns == 2);
        for (int s = 0; s < ns; ++s) {
          const int i = s / p.nblk, j = s % p.nblk;
          const uint32_t slot = (gs + s) & 1u, tp = (gt + i) & 1u;
          tr.mark(30);
          mbar_wait(BAR(B_P + slot), pwait[slot] & 1);           // P(s) is in shared memory and S(s) has been read
          ++pwait[slot];
          tr.mark(31);
          if (s == 0) mbar_wait(BAR(B_V), n & 1);
          tr.mark(32);
          if (p.nblk > 1 && j == 0 && tiles_on[tp] - 1 > ofwait[tp]) {   // separate O buffer: previous tile of this parity read out?
            mbar_wait(BAR(B_OF + tp), ofwait[tp] & 1);
            ++ofwait[tp];
          }
          tc_fence_after();
          const int nk = blk_cols(j) >> 4;
          uint64_t da = kmajor_desc(sP + slot * PT, 0), db = mnmajor_desc(sV + (uint32_t)(j * p.bw) * 128u, 0, KATOM);
          for (int k = 0; k < nk; ++k) {
            tc_mma_bf16(tmem + tp * 256 + o_off, da, db, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
            da += 2; db += 2048 >> 4;
            if ((k & 3) == 3) da += (AT_ATOM >> 4) - 8;
          }
          tc_commit(BAR(B_O + tp));                              // one phase per P V product (rescale of the next block / epilogue wait for it)
          tr.mark(33);
          if (s == ns - 1) tc_commit(BAR(B_VFREE));
          if (s + 2 < ns) issue_s(s + 2, gs, gt, s + 3 == ns);
        }
        gs += ns; gt += p.ntq;
      }
    }
   }
  } else {
    // ---------------------------------------------- row warps ----------------------------------------------
    a2_regs_row();
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int cw = p.bw >> 1;                                   // columns of an S block held by this warp: [half * cw, half * cw + cw)
    uint32_t swait[2] = {0, 0}, owait[2] = {0, 0};
    uint32_t gs = 0, gt = 0;
    float* xsum = xch + 512;                                    // [2][2][128]
    A2Trace tr; tr.init(p.lse, 0, 4096, (p.dbg & 64) && blockIdx.x == 0 && threadIdx.x == 0);
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int b = item / p.H, h = item % p.H;
      float m_run = -INFINITY, l_run = 0.f;
      // epilogue of tile i (global parity tp): O / rowsum -> global, lse; releases the O buffer
      auto epilogue = [&](int i, uint32_t tp, float m_fin, float l_part) {
        const int q = i * 128 + r;
        const bool quarter_live = i * 128 + quarter * 32 < p.N;
        tr.mark(50);
        mbar_wait(BAR(B_O + tp), owait[tp] & 1);
        ++owait[tp];
        tr.mark(51);
        tc_fence_after();
        if (quarter_live) {
          // total row sum = the two halves' partial sums (same rescaling history); buffers alternate by tile parity
          float* xs = xsum + tp * 256;
          xs[half * 128 + r] = l_part;
          named_bar_sync(1 + quarter, 64);
          const float l_tot = l_part + xs[(half ^ 1) * 128 + r];
          const float inv = __fdividef(1.f, l_tot);
          __nv_bfloat16* orow = p.out + ((long long)b * p.N + q) * p.ld_out + h * p.hd;
          const uint32_t tO = tmem + tp * 256 + o_off + lane_off;
          // one load per warp: the two halves split the head dim at a multiple of 16 (compile-time shares OS0 / OS1)
          if (half == 0) {
            uint32_t o[OL0];
            ld_o<OL0>(tO, o);
            tc_wait_ld();
            if (q < p.N) {
#pragma unroll
              for (int c = 0; c < OS0 / 16; ++c) store_row16(orow, c * 16, p.hd, *reinterpret_cast<uint32_t(*)[16]>(o + c * 16), inv);
            }
          } else if constexpr (OS1 > 0) {
            uint32_t o[OL1];
            ld_o<OL1>(tO + OS0, o);
            tc_wait_ld();
            if (q < p.N) {
#pragma unroll
              for (int c = 0; c < OS1 / 16; ++c) store_row16(orow, OS0 + c * 16, p.hd, *reinterpret_cast<uint32_t(*)[16]>(o + c * 16), inv);
            }
          }
          if (half == 0 && q < p.N && p.lse) p.lse[((long long)b * p.H + h) * p.N + q] = (m_fin + __log2f(l_tot)) * LN2;
        }
        tc_fence_before();
        tr.mark(52);
        mbar_arrive(BAR(B_OF + tp));
        tr.mark(53);
      };
      float m_prev = 0.f, l_prev = 0.f;                         // finished tile waiting for its (deferred) epilogue
      int pend_tile = -1;
      for (int s = 0; s < ns; ++s) {
        const int i = s / p.nblk, j = s % p.nblk;
        const uint32_t slot = (gs + s) & 1u, tp = (gt + i) & 1u;
        const int q = i * 128 + r;
        const bool quarter_live = i * 128 + quarter * 32 < p.N;
        const bool masked = p.row_mask && q < p.N && p.row_mask[(long long)b * p.N + q] == 0.f;
        const float sl2 = masked ? 0.f : p.scale * LOG2E;       // masked row: every score equal => p == 1 for every key
        if (j == 0) { m_run = -INFINITY; l_run = 0.f; }
        const int nvalid_blk = min(p.bw, p.N - j * p.bw);       // real keys of the block
        const int c0 = half * cw;                               // first column of this warp
        const int nvalid = max(0, min(cw, nvalid_blk - c0));    // real keys among this warp's columns
        const int n8 = (min(round16(nvalid_blk), c0 + cw) - c0 + 7) >> 3;   // 8-column groups the MMA computed for this warp (<= 0: none)
        tr.mark(40);
        mbar_wait(BAR(B_S + slot), swait[slot] & 1);
        ++swait[slot];
        tr.mark(41);
        tc_fence_after();
        if (quarter_live) {
          uint32_t sv[LDW];
          const uint32_t tS = tmem + slot * 256 + lane_off + (uint32_t)c0;
          // ONE load for the warp's whole share (columns past its share / past the block are read and ignored)
          if constexpr (LDW == 32) tc_ld32(tS, sv); else if constexpr (LDW == 64) tc_ld64(tS, sv); else tc_ld128(tS, sv);
          tc_wait_ld();
          tr.mark(45);
          // Padding keys (columns the tensor pipe computed beyond the sequence end: fewer than 16, i.e. at most two groups) are
          // poisoned to -1e30 once, so the two passes below run without per-element predicates: they cannot win the row maximum
          // and their weight is 2^(-1e30 * scale - m) == 0.  (A masked query row multiplies every score by 0: its padding
          // weights are exactly 1 and are subtracted from the row sum in closed form.)
          const int n8u = __shfl_sync(0xffffffffu, n8, 0);      // computed 8-column groups of this warp (warp-uniform by construction)
          const int nfu = __shfl_sync(0xffffffffu, nvalid >> 3, 0);   // groups made of real keys only
#pragma unroll
          for (int g8 = 0; g8 < LDW / 8; ++g8) {
            if (g8 >= nfu && g8 < n8u) {
#pragma unroll
              for (int t = 0; t < 8; ++t)
                if (g8 * 8 + t >= nvalid) sv[g8 * 8 + t] = __float_as_uint(-1e30f);
            }
          }
          float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
          for (int g8 = 0; g8 < LDW / 8; ++g8) {
            if (g8 < n8u) {
#pragma unroll
              for (int t = 0; t < 8; t += 2) {
                mx0 = fmaxf(mx0, __uint_as_float(sv[g8 * 8 + t]));
                mx1 = fmaxf(mx1, __uint_as_float(sv[g8 * 8 + t + 1]));
              }
            }
          }
          float mx = fmaxf(mx0, mx1);
          // both halves of a row must subtract the same reference: exchange the maxima (buffers alternate by step parity)
          float* xm = xch + (slot * 2) * 128;
          xm[half * 128 + r] = mx;
          tr.mark(46);
          named_bar_sync(1 + quarter, 64);
          mx = fmaxf(mx, xm[(half ^ 1) * 128 + r]);
          tr.mark(47);
          const float m_new = masked ? 0.f : fmaxf(m_run, mx * sl2);
          const float alpha = ex2f(m_run - m_new);
          float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
          const uint32_t prow = sP + slot * PT + (uint32_t)r * 128u;
          const int cb = c0 >> 3, rx = r & 7;
#pragma unroll
          for (int g8 = 0; g8 < LDW / 8; ++g8) {
            if (g8 < n8u) {
              float v[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) v[t] = ex2f(fmaf(__uint_as_float(sv[g8 * 8 + t]), sl2, -m_new));
              rs0 += v[0] + v[4]; rs1 += v[1] + v[5]; rs2 += v[2] + v[6]; rs3 += v[3] + v[7];
              const int c16 = cb + g8;
              const uint32_t addr = prow + (uint32_t)(c16 >> 3) * AT_ATOM + (uint32_t)(((c16 & 7) ^ rx) << 4);
              __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                           "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
                           "r"(*reinterpret_cast<uint32_t*>(&h3)) : "memory");
            }
          }
          const int npad = max(0, n8u * 8 - nvalid);             // computed padding columns of this warp
          const float rowsum = (rs0 + rs1) + (rs2 + rs3) - (masked ? (float)npad : 0.f);
          l_run = l_run * alpha + rowsum;
          m_run = m_new;
          if (j > 0) {                                          // O += needs the old O scaled: wait for P V(j-1), rescale in TMEM
            mbar_wait(BAR(B_O + tp), owait[tp] & 1);
            ++owait[tp];
            tc_fence_after();
            if (!__all_sync(0xffffffffu, alpha == 1.f)) {
              const uint32_t tO = tmem + tp * 256 + o_off + lane_off;
              if (half == 0) {
                uint32_t o[OL0];
                ld_o<OL0>(tO, o);
                tc_wait_ld();
#pragma unroll
                for (int t = 0; t < OL0; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
                st_o<OL0>(tO, o);
              } else if constexpr (OS1 > 0) {
                uint32_t o[OL1];
                ld_o<OL1>(tO + OS0, o);
                tc_wait_ld();
#pragma unroll
                for (int t = 0; t < OL1; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
                st_o<OL1>(tO + OS0, o);
              }
              tc_wait_st();
            }
          }
        } else if (j > 0) {                                     // dead quarter: stay in step with the barrier phases
          mbar_wait(BAR(B_O + tp), owait[tp] & 1);
          ++owait[tp];
        }
        tr.mark(42);
        fence_async_smem();
        tc_fence_before();
        tr.mark(43);
        mbar_arrive(BAR(B_P + slot));
        tr.mark(44);
        // deferred epilogue of the tile finished one step ago: its last P V product ran while this step was processed
        if (pend_tile >= 0) { epilogue(pend_tile, (gt + pend_tile) & 1u, m_prev, l_prev); pend_tile = -1; }
        if (j == p.nblk - 1) { pend_tile = i; m_prev = m_run; l_prev = l_run; }
      }
      if (pend_tile >= 0) epilogue(pend_tile, (gt + pend_tile) & 1u, m_prev, l_prev);
      gs += ns; gt += p.ntq;
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
template <int NA, int HD16>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn_bwd_v2_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do, AttnV2Params p) {
  pdl_trigger();   // the GEMM behind this kernel may set itself up during its tail (common.cuh)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t KATOM = (uint32_t)p.krows * 128u, KT = NA * KATOM;
  constexpr uint32_t BT = 2 * AT_ATOM;                        // P / dS block: 128 query rows x 128 keys (bf16)
  const uint32_t sQ = base, sdO = sQ + KT, sK = sdO + KT, sV = sK + KT, sP = sV + KT, sdS = sP + BT, stat = sdS + BT;
  const uint32_t stat_rows = (uint32_t)p.ntq * 128u;
  const uint32_t stat_bytes = 2u * stat_rows * 12u;           // two buffers of {float2 {lse*log2e, delta}, int flag} per query row
  const uint32_t bars = (stat + stat_bytes + 15u) & ~15u;
  enum { B_LD0 = 0, B_LD1 = 1, B_FREE = 2, B_S = 3, B_C = 4, B_P = 5, B_G = 6, NBAR = 7 };
  auto BAR = [&](int i) { return bars + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = bars + 8u * NBAR;
  float2* s_stat_all = reinterpret_cast<float2*>(smem_raw + (stat - base));                         // [2][stat_rows]
  int* s_flag_all = reinterpret_cast<int*>(smem_raw + (stat - base) + (size_t)2 * stat_rows * 8);   // [2][stat_rows]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NO = HD16 * 16;                   // accumulator columns of the head-dim-wide products (= p.n_o)
  constexpr int OS0 = ((HD16 + 1) / 2) * 16, OS1 = NO - OS0;   // dQ columns of the two warps of a lane quarter
  constexpr int OL0 = OS0 <= 16 ? 16 : (OS0 <= 32 ? 32 : 64), OL1 = OS1 <= 16 ? 16 : (OS1 <= 32 ? 32 : 64);
  constexpr int OLF = NO <= 16 ? 16 : (NO <= 32 ? 32 : (NO <= 64 ? 64 : 128));   // one load for a whole dV / dK row
  const int nkb = (p.N + 127) / 128;              // 128-key blocks (the M tile of dV / dK)
  const int nb = nkb * p.ntq;                     // (query tile, key block) blocks per item, key block outermost

  if (threadIdx.x == 0) {
    if (base & 1023u) __trap();
    for (int i = 0; i < NBAR; ++i) mbar_init(BAR(i), (i == B_P || i == B_C) ? 256u : 1u);
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

  if (warp >= 8) {
   a2_regs_aux();
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
          mbar_wait(BAR(B_C), pc & 1);          // S / dP of this block are in the row warps' registers: their TMEM columns are free
          tc_fence_after();
          if (bidx + 1 < nb) issue_sdp(bidx + 1);    // runs while the warps exponentiate block bidx
          mbar_wait(BAR(B_P), pc & 1);          // P / dS of this block are in shared memory
          ++pc;
          tc_fence_after();
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
   }
  } else {
    a2_regs_row();
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const int tid = threadIdx.x;                                     // 0..255
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    const float log2n = log2f((float)p.N);
    const long long HD = (long long)p.H * p.hd;
    uint32_t sc = 0, gc = 0;
    A2Trace tr; tr.init(p.lse, 0, 4096, (p.dbg & 64) && blockIdx.x == 0 && threadIdx.x == 0);
    // per-row statistics of one item into buffer `buf`: {lse * log2e, delta = rowsum(dO * O)};
    // flag 1 = live, 2 = masked query row (p = 1/N, dS = 0), 0 = beyond N
    auto compute_stats = [&](int item, int buf) {
      const int b = item / p.H, h = item % p.H;
      float2* s_stat = s_stat_all + (size_t)buf * stat_rows;
      int* s_flag = s_flag_all + (size_t)buf * stat_rows;
      for (int row = tid; row < (int)stat_rows; row += 256) {
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
#pragma unroll
            for (int c = 0; c < 2 * HD16; ++c) {              // compile-time bound: the loads of a row are scheduled together
              if (c * 8 < p.hd) {
                const uint4 a = *reinterpret_cast<const uint4*>(orow + c * 8), e = *reinterpret_cast<const uint4*>(drow + c * 8);
                const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
                const __nv_bfloat162* ep = reinterpret_cast<const __nv_bfloat162*>(&e);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float2 x = __bfloat1622float2(ap[t]), y = __bfloat1622float2(ep[t]);
                  d += x.x * y.x + x.y * y.y;
                }
              }
            }
            st = make_float2(p.lse[((long long)b * p.H + h) * p.N + row] * LOG2E, d);
            fl = 1;
          }
        }
        s_stat[row] = st;
        s_flag[row] = fl;
      }
    };
    int n = 0;
    if (blockIdx.x < p.items) compute_stats(blockIdx.x, 0);
    named_bar_sync(1, 256);
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      const int b = item / p.H, h = item % p.H;
      const float2* s_stat = s_stat_all + (size_t)(n & 1) * stat_rows;
      const int* s_flag = s_flag_all + (size_t)(n & 1) * stat_rows;

      auto epilogue_dkv = [&](int j) {       // half 0 stores dV, half 1 stores dK of key block j: one TMEM load per warp
        const int kv = j * 128 + r;
        if (j * 128 + quarter * 32 < p.N) {
          __nv_bfloat16* krow = p.dqkv + ((long long)b * p.N + kv) * p.ld_dqkv + HD + h * p.hd;
          __nv_bfloat16* vrow = krow + HD;
          uint32_t o[OLF];
          if constexpr (OLF == 128) tc_ld128((half == 0 ? tdV : tdK) + lane_off, o); else ld_o<OLF>((half == 0 ? tdV : tdK) + lane_off, o);
          tc_wait_ld();
          if (kv < p.N) {
#pragma unroll
            for (int c = 0; c < HD16; ++c)
              store_row16(half == 0 ? vrow : krow, c * 16, p.hd, *reinterpret_cast<uint32_t(*)[16]>(o + c * 16), half == 0 ? 1.f : p.scale);
          }
        }
      };

      for (int bidx = 0; bidx < nb; ++bidx) {
        const int j = bidx / p.ntq, i = bidx % p.ntq;
        const int nc = round16(min(128, p.N - j * 128));
        const bool quarter_live = i * 128 + quarter * 32 < p.N;
        // this warp's 64 contiguous key columns of the block: [64 * half, 64 * half + 64); one load each for S and dP, results
        // packed to bf16 in registers
        uint32_t pk[32], dk[32];
        const int w = max(0, min(64, nc - 64 * half));          // computed columns of this warp's share (multiple of 16, warp-uniform)
        tr.mark(40);
        mbar_wait(BAR(B_S), sc & 1);
        ++sc;
        tr.mark(41);
        tc_fence_after();
        if (quarter_live && w > 0) {
          const float2 st = s_stat[i * 128 + r];
          const int fl = s_flag[i * 128 + r];
          uint32_t sreg[64], dreg[64];
          tc_ld64(tS + lane_off + 64 * half, sreg);
          tc_ld64(tdP + lane_off + 64 * half, dreg);
          tc_wait_ld();
          if (fl == 1) {                                        // live query row: p = 2^(s * scale * log2e - lse2), dS = p (dP - delta)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (g * 16 < w) {                                 // warp-uniform
#pragma unroll
                for (int t = 8 * g; t < 8 * g + 8; ++t) {
                  const float p0 = ex2f(fmaf(__uint_as_float(sreg[2 * t]), sl2, -st.x));
                  const float p1 = ex2f(fmaf(__uint_as_float(sreg[2 * t + 1]), sl2, -st.x));
                  const float d0 = p0 * (__uint_as_float(dreg[2 * t]) - st.y);
                  const float d1 = p1 * (__uint_as_float(dreg[2 * t + 1]) - st.y);
                  __nv_bfloat162 hp = __floats2bfloat162_rn(p0, p1), hd2 = __floats2bfloat162_rn(d0, d1);
                  pk[t] = *reinterpret_cast<uint32_t*>(&hp);
                  dk[t] = *reinterpret_cast<uint32_t*>(&hd2);
                }
              }
            }
          } else {                                              // masked query row: p = 1/N, dS = 0; row beyond the sequence: zeros
            const float pc = fl == 2 ? ex2f(-st.x) : 0.f;
            __nv_bfloat162 hp = __floats2bfloat162_rn(pc, pc);
            const uint32_t pw = *reinterpret_cast<uint32_t*>(&hp);
#pragma unroll
            for (int t = 0; t < 32; ++t) { pk[t] = pw; dk[t] = 0u; }
          }
        }
        tr.mark(42);
        tc_fence_before();
        mbar_arrive(BAR(B_C));                  // S / dP are in registers: the tensor pipe may overwrite them with the next block
        if (bidx > 0) {                         // gradient products of the previous block are done: P / dS may be overwritten,
          tr.mark(43);
          mbar_wait(BAR(B_G), gc & 1);          // and dV / dK of a finished key block are final
          ++gc;
          tr.mark(44);
          tc_fence_after();
          if (i == 0) epilogue_dkv(j - 1);
          tr.mark(45);
        }
        if (quarter_live) {
#pragma unroll
          for (int g8 = 0; g8 < 8; ++g8) {
            if (g8 * 8 < w) {
              // 16-byte chunk g8 of atom `half` (64 keys per atom), row r, 128-byte swizzle
              const uint32_t off = (uint32_t)half * AT_ATOM + (uint32_t)r * 128u + (uint32_t)((g8 ^ (r & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + off), "r"(pk[4 * g8]), "r"(pk[4 * g8 + 1]),
                           "r"(pk[4 * g8 + 2]), "r"(pk[4 * g8 + 3]) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdS + off), "r"(dk[4 * g8]), "r"(dk[4 * g8 + 1]),
                           "r"(dk[4 * g8 + 2]), "r"(dk[4 * g8 + 3]) : "memory");
            }
          }
        }
        tr.mark(46);
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(BAR(B_P));
        tr.mark(47);
      }
      // the statistics of the CTA's next item are fetched while the last gradient products run
      const int next_item = item + gridDim.x;
      if (next_item < p.items) compute_stats(next_item, (n + 1) & 1);
      tr.mark(50);
      mbar_wait(BAR(B_G), gc & 1);
      ++gc;
      tr.mark(51);
      tc_fence_after();
      epilogue_dkv(nkb - 1);
      tr.mark(52);
      for (int i = 0; i < p.ntq; ++i) {
        const int q = i * 128 + r;
        if (i * 128 + quarter * 32 < p.N) {
          __nv_bfloat16* drow = p.dqkv + ((long long)b * p.N + q) * p.ld_dqkv + h * p.hd;
          const uint32_t tq = tdQ + i * NO + lane_off;
          if (half == 0) {
            uint32_t o[OL0];
            ld_o<OL0>(tq, o);
            tc_wait_ld();
            if (q < p.N) {
#pragma unroll
              for (int c = 0; c < OS0 / 16; ++c) store_row16(drow, c * 16, p.hd, *reinterpret_cast<uint32_t(*)[16]>(o + c * 16), p.scale);
            }
          } else if constexpr (OS1 > 0) {
            uint32_t o[OL1];
            ld_o<OL1>(tq + OS0, o);
            tc_wait_ld();
            if (q < p.N) {
#pragma unroll
              for (int c = 0; c < OS1 / 16; ++c) store_row16(drow, OS0 + c * 16, p.hd, *reinterpret_cast<uint32_t(*)[16]>(o + c * 16), p.scale);
            }
          }
        }
      }
      tr.mark(53);
      tc_fence_before();
      named_bar_sync(1, 256);       // next item's statistics are complete and visible to every row warp
      tr.mark(54);
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
// The persistent FORWARD kernel is opt-in (BFK_ATTN_V2_FWD=1, or backend == BFK_BACKEND_TCGEN05_PERSISTENT): measured on B200
// (profiles/r02_attention_notes.md) it ties the generic kernel at N = 196 (48 vs 49 us) and loses at N = 262 / 64, where the
// 3-CTA-per-SM generic kernel overlaps the phases of different tiles better than one persistent CTA does.  The merged backward
// kernel is the default: 27-31 % faster than the dQ + dK/dV + delta trio.
static bool v2_fwd_enabled(const bfk_attention_args* a) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("BFK_ATTN_V2_FWD"); v = (e && atoi(e) == 1) ? 1 : 0; }
  return v == 1 || a->backend == BFK_BACKEND_TCGEN05_PERSISTENT;
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
  { static int v = -1; if (v < 0) { const char* e = getenv("BFK_ATTN_DBG"); v = e ? atoi(e) : 0; } p.dbg = v; }
}

// forward plan: smallest number of key blocks whose S (+ O) tiles fit two groups into 512 TMEM columns
static bool plan_fwd(const bfk_attention_args* a, AttnV2Params& p, size_t& smem) {
  if (!v2_enabled() || !v2_fwd_enabled(a) || a->N > 1024) return false;
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
  smem = (size_t)2 * NA * AT_ATOM + (size_t)2 * NA * p.krows * 128 + (size_t)2 * p.patoms * AT_ATOM + 4 * 1024 + 8 * 16 + 16;
  return smem <= A2_SMEM_MAX;
}

static bool plan_bwd(const bfk_attention_args* a, AttnV2Params& p, size_t& smem) {
  if (!v2_enabled() || a->N > 1024) return false;
  plan_common(a, p);
  const int NA = a->hd <= 64 ? 1 : 2;
  if (256 + (2 + p.ntq) * p.n_o > 512) return false;                   // S, dP, dV, dK and one dQ per query tile in TMEM
  smem = (size_t)4 * NA * p.krows * 128 + (size_t)4 * AT_ATOM + (size_t)2 * p.ntq * 128 * 12 + 16 + 8 * 7 + 16;
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
  const int cw = p.bw / 2;
#define A2_FWD3(NA_, LDW_, HD16_)                                                                       \
  do {                                                                                                  \
    if (int rc = set_smem(attn_fwd_v2_kernel<NA_, LDW_, HD16_>, smem, "attn_fwd_v2")) return rc;        \
    attn_fwd_v2_kernel<NA_, LDW_, HD16_><<<grid, A2_THREADS, smem, s>>>(mq, mkv, p);                    \
  } while (0)
#define A2_FWD(NA_, HD16_)                                                                              \
  do {                                                                                                  \
    if (cw <= 32) A2_FWD3(NA_, 32, HD16_); else if (cw <= 64) A2_FWD3(NA_, 64, HD16_); else A2_FWD3(NA_, 128, HD16_); \
  } while (0)
  switch (p.n_o / 16) {      // head dims of the policy: 64 (towers), 48 (fusion), 72 -> 80 (so400m); 16 / 32 for small test shapes
    case 1: A2_FWD(1, 1); break;
    case 2: A2_FWD(1, 2); break;
    case 3: A2_FWD(1, 3); break;
    case 4: A2_FWD(1, 4); break;
    case 5: A2_FWD(2, 5); break;
    default: return 1;
  }
#undef A2_FWD3
#undef A2_FWD
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
#define A2_BWD(NA_, HD16_)                                                                          \
  do {                                                                                              \
    if (int rc = set_smem(attn_bwd_v2_kernel<NA_, HD16_>, smem, "attn_bwd_v2")) return rc;          \
    attn_bwd_v2_kernel<NA_, HD16_><<<grid, A2_THREADS, smem, s>>>(mqkv, mdo, p);                    \
  } while (0)
  switch (p.n_o / 16) {
    case 1: A2_BWD(1, 1); break;
    case 2: A2_BWD(1, 2); break;
    case 3: A2_BWD(1, 3); break;
    case 4: A2_BWD(1, 4); break;
    case 5: A2_BWD(2, 5); break;
    default: return 1;
  }
#undef A2_BWD
  BFK_CHECK_LAUNCH("attention_bwd_v2");
  return BFK_OK;
}

}  // namespace bfk
