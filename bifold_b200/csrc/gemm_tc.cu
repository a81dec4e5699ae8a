// tcgen05 GEMM for sm_100a: D = epi(A * B^T), bf16 operands, fp32 accumulation in TMEM.
//
//   * persistent: one CTA per SM loops over 128 x BN output tiles (BN = 256 or 128)
//   * warp 0      : TMA producer  (cp.async.bulk.tensor.2d, 128B swizzle, 64-wide K blocks, mbarrier ring)
//   * warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, kind::f16)
//   * warps 2..5  : epilogue (tcgen05.ld 32x32b -> registers -> fused bias/GELU/residual -> global)
//   * two TMEM accumulator buffers so the epilogue of tile i overlaps the main loop of tile i+1
//   * operands may be K-major (row-major [rows,K]) or MN-major (row-major [K,rows]); the latter serves
//     dgrad (B = W as stored) and wgrad (A = dY, B = X as stored) without any transposed copies.
#include <mutex>
#include <unordered_map>

#include "tc_epilogue.cuh"

namespace bfk {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;

template <int BN> struct TcCfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 256 ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * BN * 4 /*bias*/;
};

// MODE: compile-time EpiMode (see gemm_tc2.cu): EPI_GENERIC = runtime-flag epilogue, 8 epilogue warps; a specialised
// MODE compiles only its own epilogue and runs 16 epilogue warps.
template <int BN, bool AMN, bool BMN, int CL, int MODE>
__global__ void __launch_bounds__(64 + 32 * epi_warps(MODE), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M, int N, int K,
               EpiParams ep, int splits, long long split_stride, int dbg) {
  using Cfg = TcCfg<BN>;
  constexpr int EW = epi_warps(MODE);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
  float* bias_smem = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));   // [2][BN]
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (M + TC_BM - 1) / TC_BM, num_n = (N + BN - 1) / BN;
  // CL = 2: the two CTAs of a cluster work in lock-step on vertically adjacent output tiles (same n-block) and
  // share the B operand: each loads one half of the B tile and multicasts it to both (halves the B traffic from L2).
  const int crank = CL == 2 ? (int)cluster_ctarank() : 0;
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
  const int num_mp = (num_m + CL - 1) / CL;
  const int tiles = num_mp * num_n * splits;         // work items per cluster: (row-pair, n-block, K split)
  const int num_kb_all = (K + TC_BK - 1) / TC_BK;
  const int kb_per = (num_kb_all + splits - 1) / splits;
  auto kb_range = [&](int item, int& kb0, int& kb1) {
    const int sp = item % splits;
    kb0 = sp * kb_per;
    kb1 = min(num_kb_all, kb0 + kb_per);
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CL); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) cluster_sync_all();   // peer barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();     // see common.cuh: a dependent kernel may be scheduled from here on
  pdl_wait();        // PDL launch: the set-up above overlapped the predecessor's tail; its results may be read from here on

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer =====
      int stage = 0; uint32_t phase = 0;
      for (int item = cid; item < tiles; item += ncl) {
        const int tile = item / splits;
        const int m0 = ((tile / num_n) * CL + crank) * TC_BM, n0 = (tile % num_n) * BN;
        int kb0, kb1;
        kb_range(item, kb0, kb1);
        if (dbg & 2) continue;   // diagnostic: epilogue only
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
          const int k0 = kb * TC_BK;
          if constexpr (!AMN) {
            tma_load_2d(sa, &tma_a, full_bar(stage), k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < TC_BM / 64; ++i) tma_load_2d(sa + i * 8192, &tma_a, full_bar(stage), m0 + 64 * i, k0);
          }
          if constexpr (CL == 1) {
            if constexpr (!BMN) {
              tma_load_2d(sb, &tma_b, full_bar(stage), k0, n0);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i) tma_load_2d(sb + i * 8192, &tma_b, full_bar(stage), n0 + 64 * i, k0);
            }
          } else {   // this CTA's half of B, multicast to both CTAs of the cluster
            if constexpr (!BMN) {
              tma_load_2d_mc(sb + crank * (Cfg::B_BYTES / 2), &tma_b, full_bar(stage), k0, n0 + crank * (BN / 2), (uint16_t)0x3);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 128; ++i) {
                const int bi = crank * (BN / 128) + i;
                tma_load_2d_mc(sb + bi * 8192, &tma_b, full_bar(stage), n0 + 64 * bi, k0, (uint16_t)0x3);
              }
            }
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc(TC_BM, BN, AMN, BMN);
      constexpr int KSTEP_A = (AMN ? 2048 : 32) >> 4, KSTEP_B = (BMN ? 2048 : 32) >> 4;   // descriptor step per UMMA_K slice
      const uint64_t da0 = operand_desc<AMN>(smem_base, 0), db0 = operand_desc<BMN>(smem_base + Cfg::A_BYTES, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int item = cid; item < tiles; item += ncl) {
        int kb0, kb1;
        kb_range(item, kb0, kb1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1 && !(dbg & 2); ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t soff = (uint64_t)(stage * (Cfg::STAGE_BYTES >> 4));   // descriptors are built once (da0, db0)
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            tc_mma_bf16(tmem_d, da0 + soff + (uint64_t)(k * KSTEP_A), db0 + soff + (uint64_t)(k * KSTEP_B), idesc,
                        (kb > kb0 || k > 0) ? 1u : 0u);
          if constexpr (CL == 1) tc_commit(empty_bar(stage));
          else tc_commit_mc(empty_bar(stage), (uint16_t)0x3);   // the slot is free in BOTH CTAs' eyes only when both consumed it
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter = warp % 4; the EW/4 warps of a quarter take every (EW/4)-th chunk =====
    const int quarter = warp & 3, sub = (warp - 2) >> 2;
    const int half = sub;   // generic path: two warps per quarter
    const int et = threadIdx.x - 64;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = cid; item < tiles; item += ncl) {
      const int tile = item / splits;
      const int m0 = ((tile / num_n) * CL + crank) * TC_BM, n0 = (tile % num_n) * BN;
      EpiParams epl = ep;
      if (splits > 1) epl.D = reinterpret_cast<float*>(ep.D) + (long long)(item % splits) * split_stride;   // fp32 partial
      float* bias_s = bias_smem + acc * BN;
      for (int i = et; i < BN; i += 32 * EW) bias_s[i] = (ep.bias && n0 + i < N) ? __ldg(ep.bias + n0 + i) : 0.f;
      named_bar_sync(1, 32 * EW);   // bias visible; also orders reuse of the other buffer
      if (ep.vec_ok && item + ncl < tiles) {   // warm L2 with what the NEXT item's epilogue will read
        const int ntile = (item + ncl) / splits;
        const int nm = ((ntile / num_n) * CL + crank) * TC_BM + quarter * 32 + lane, nn0 = (ntile % num_n) * BN;
        for (int c = sub; c < BN / 32; c += EW / 4) epi_prefetch(ep, nm, nn0 + c * 32);
      }
      const int m = m0 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      const bool skip = (dbg & 1) || (ep.vec_ok & 16);   // measurement only
      if constexpr (MODE != EPI_GENERIC) {
        EpiOps<MODE> ops;
        epi_tile_preload<MODE>(epl, m, n0, sub, ops);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        if (!skip) epi_tile<MODE, (BN / 32) / (EW / 4), EW / 4>(epl, m, n0, sub, taddr, bias_s, ops);
      } else {
        EpiOperands cur, nxt;
        if (n0 + half * 32 < N) epi_load(epl, m, n0 + half * 32, cur);   // first chunk's operands: issued before the wait
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          if (n0 + c * 32 >= N || skip) break;  // warp-uniform
          if (c + 2 < BN / 32 && n0 + (c + 2) * 32 < N) epi_load(epl, m, n0 + (c + 2) * 32, nxt);   // one chunk ahead
          epi_chunk32(epl, m, n0 + c * 32, taddr + (uint32_t)(c * 32), bias_s + c * 32, cur);
          cur = nxt;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) cluster_sync_all();   // nobody exits while the peer may still multicast into / signal this CTA
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
  }
}

// ---- host side: tensor maps --------------------------------------------------------------------------
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    else
      (void)cudaGetLastError();
  });
  return fn;
}
static PFN_encodeTiled get_encode() { return get_encode_tiled(); }

bool make_map(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long ld, int box_inner, int box_outer, int swizzle_bytes = 128,
              int elem_bytes = 2);
struct MapKey {
  const void* ptr; long long inner, outer, ld; int box_inner, box_outer;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner && box_outer == o.box_outer;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&](size_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix((size_t)k.inner); mix((size_t)k.outer); mix((size_t)k.ld); mix((size_t)k.box_inner); mix((size_t)k.box_outer);
    return h;
  }
};

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows of pitch ld elements; 128B swizzle.
bool make_map(CUtensorMap* out, const void* ptr, long long inner, long long outer, long long ld, int box_inner, int box_outer, int swizzle_bytes,
              int elem_bytes) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, inner, outer, ld, box_inner, box_outer + (swizzle_bytes << 16) + (elem_bytes << 28)};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return true; }
  }
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("gemm_tc: cuTensorMapEncodeTiled unavailable"); return false; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (cuuint64_t)elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld", (int)r, inner, outer, ld); return false; }
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return true;
}

bool gemm_tc_supported(const bfk_gemm_args* a, const char** why) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if (a->in_dtype != BFK_BF16) { *why = "operands must be bf16"; return false; }
  if (!al16(a->A) || !al16(a->B)) { *why = "operand pointers must be 16-byte aligned"; return false; }
  if ((a->lda * 2) % 16 != 0 || (a->ldb * 2) % 16 != 0) { *why = "leading dimensions must be multiples of 8 elements"; return false; }
  if (!get_encode()) { *why = "driver entry point cuTensorMapEncodeTiled missing"; return false; }
  return true;
}

// D[m*ldd+n] = alpha * sum_s ws[s][m][n]   (fixed order => deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long split_stride, int M, int N, float alpha,
                                     float* __restrict__ D, long long ldd, int accumulate) {
  const long long total = (long long)M * N;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += ws[k * split_stride + i];
    const int m = (int)(i / N), n = (int)(i % N);
    float* d = D + (long long)m * ldd + n;
    *d = accumulate ? (*d + alpha * s) : alpha * s;
  }
}
// same, four columns per thread (N, ldd, split_stride multiples of 4, 16-byte aligned pointers): the partials are read once as
// float4 streams and the row / column of an element comes from one division per four elements
__global__ void __launch_bounds__(256) splitk_reduce4_kernel(const float4* __restrict__ ws, int splits, long long split_stride4, int M,
                                                             int N4, float alpha, float* __restrict__ D, long long ldd, int accumulate) {
  const long long total = (long long)M * N4;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    float4 s = __ldcs(ws + i);
    for (int k = 1; k < splits; ++k) {
      const float4 v = __ldcs(ws + k * split_stride4 + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const int m = (int)(i / N4), n4 = (int)(i - (long long)m * N4);
    float4* d = reinterpret_cast<float4*>(D + (long long)m * ldd) + n4;
    float4 o = make_float4(alpha * s.x, alpha * s.y, alpha * s.z, alpha * s.w);
    if (accumulate) { const float4 c = *d; o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
    *d = o;
  }
}

// K splits for short-and-wide problems (wgrad: few output tiles, very long K) so every SM gets work.
static int choose_splits(const bfk_gemm_args* a, int BN) {
  if (a->split_k == 1 || !a->workspace) return 1;
  const bool plain = !a->bias && a->act == BFK_ACT_NONE && !a->residual && a->out_dtype == BFK_F32;
  if (!plain) return 1;
  const int tiles = ((a->M + TC_BM - 1) / TC_BM) * ((a->N + BN - 1) / BN);
  const int num_kb = (a->K + TC_BK - 1) / TC_BK;
  int splits;
  if (a->split_k > 1) splits = a->split_k;
  else {
    if (tiles * 2 > sm_count()) return 1;
    splits = std::min(16, sm_count() / tiles);
  }
  splits = std::min(splits, std::max(1, num_kb / 4));
  while (splits > 1 && (size_t)splits * a->M * a->N * 4 > (size_t)a->workspace_bytes) --splits;
  // no empty split: ceil-div partition must leave work for the last one
  while (splits > 1 && (splits - 1) * ((num_kb + splits - 1) / splits) >= num_kb) --splits;
  return std::max(1, splits);
}

int gemm_tc2(const bfk_gemm_args* a, cudaStream_t s);   // CTA-pair (cta_group::2) variant, gemm_tc2.cu
static bool pair_enabled() {
  static int v = -1;
  // On by default: the 256x256 pair tile needs 32 KB of operands per SM and K block instead of 48 KB; measured on
  // B200 (profiles/r01_gemm_vs_cublas.txt) it is 2-14% faster than the single-CTA kernel on the policy's big shapes.
  if (v < 0) { const char* e = getenv("BFK_GEMM_2CTA"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
static bool cluster_enabled() {
  static int v = -1;
  // Off by default: measured on B200 (profiles/r01_gemm_cluster_multicast.txt) the cluster-of-2 B multicast is
  // correct but not faster -- L2 already de-duplicates neighbouring unicast requests at this cluster size.
  if (v < 0) { const char* e = getenv("BFK_GEMM_CLUSTER"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

int gemm_dbg() {   // BFK_GEMM_DBG: 1 = skip the epilogue body, 2 = skip TMA + MMA (measurement only; results are garbage)
  static int v = -1;
  if (v < 0) { const char* e = getenv("BFK_GEMM_DBG"); v = e ? atoi(e) : 0; }
  return v;
}

template <int BN, bool AMN, bool BMN, int CL, int MODE>
static int launch_tc_mode(const bfk_gemm_args* a, cudaStream_t s, const CUtensorMap& ma, const CUtensorMap& mb, const EpiParams& ep,
                          int splits, long long split_stride, int grid) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, AMN, BMN, CL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("gemm_tc: cannot set smem %d: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e)); return BFK_E_CUDA; }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + 32 * epi_warps(MODE)); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, AMN, BMN, CL, MODE>, ma, mb, a->M, a->N, a->K, ep, splits, split_stride, gemm_dbg());
  if (le != cudaSuccess) { set_error("gemm_tc: launch failed: %s", cudaGetErrorString(le)); return BFK_E_CUDA; }
  return BFK_OK;
}

// partial_only: leave the fp32 K-split partials [splits][M][N] in the workspace (even for splits == 1) and skip the
// reduction -- the caller fuses it into its own kernel (bfk_lora_grads).
template <int BN, bool AMN, bool BMN, int CL>
static int launch_tc_cl(const bfk_gemm_args* a, cudaStream_t s, int splits, bool partial_only = false) {
  using Cfg = TcCfg<BN>;
  CUtensorMap ma, mb;
  bool ok;
  if (!AMN) ok = make_map(&ma, a->A, a->K, a->M, a->lda, TC_BK, TC_BM);
  else      ok = make_map(&ma, a->A, a->M, a->K, a->lda, 64, TC_BK);
  if (!ok) return BFK_E_CUDA;
  if (!BMN) ok = make_map(&mb, a->B, a->K, a->N, a->ldb, TC_BK, BN / CL);
  else      ok = make_map(&mb, a->B, a->N, a->K, a->ldb, 64, TC_BK);
  if (!ok) return BFK_E_CUDA;
  EpiParams ep;
  epi_from_args(a, ep);
  const long long split_stride = (long long)a->M * a->N;
  if (splits > 1 || partial_only) {   // partial tiles go to the workspace as plain fp32 [splits][M][N]
    ep.D = a->workspace; ep.ldd = a->N; ep.out_dtype = BFK_F32; ep.alpha = 1.f; ep.accumulate = 0;
    ep.vec_ok = ((reinterpret_cast<uintptr_t>(a->workspace) & 31u) == 0 && (a->N % 8) == 0 && (split_stride % 8) == 0) ? 1 : 0;
    ep.mode = ep.vec_ok ? EPI_STORE_F32 : EPI_GENERIC;
  }
  if (gemm_dbg() & 32) ep.mode = EPI_GENERIC;     // measurement only: force the generic epilogue
  if (ep.vec_ok) ep.vec_ok |= gemm_dbg() & 92;     // bit 6 (64): no L2 prefetch of the next tile's epilogue operands
  const int num_m = (a->M + TC_BM - 1) / TC_BM, num_n = (a->N + BN - 1) / BN;
  const int items = ((num_m + CL - 1) / CL) * num_n * splits;
  const int grid = CL * std::min(items, sm_count() / CL);
  // specialised epilogues are instantiated only where the step uses them on this kernel: plain fp32 stores of the
  // wgrad layout (split-K partials) and plain bf16 stores of the dgrad layout
  constexpr bool HAS_F32 = CL == 1 && AMN && BMN, HAS_BF16 = CL == 1 && !AMN && BMN;
  int rc;
  if (HAS_F32 && ep.mode == EPI_STORE_F32) rc = launch_tc_mode<BN, AMN, BMN, CL, HAS_F32 ? EPI_STORE_F32 : EPI_GENERIC>(a, s, ma, mb, ep, splits, split_stride, grid);
  else if (HAS_BF16 && ep.mode == EPI_STORE_BF16) rc = launch_tc_mode<BN, AMN, BMN, CL, HAS_BF16 ? EPI_STORE_BF16 : EPI_GENERIC>(a, s, ma, mb, ep, splits, split_stride, grid);
  else { ep.mode = EPI_GENERIC; rc = launch_tc_mode<BN, AMN, BMN, CL, EPI_GENERIC>(a, s, ma, mb, ep, splits, split_stride, grid); }
  if (rc) return rc;
  if (splits > 1 && !partial_only) {
    const long long total = (long long)a->M * a->N;
    const bool v4 = a->N % 4 == 0 && a->ldd % 4 == 0 && split_stride % 4 == 0 &&
                    ((reinterpret_cast<uintptr_t>(a->workspace) | reinterpret_cast<uintptr_t>(a->D)) & 15u) == 0;
    if (v4) {
      const int rg = (int)std::min<long long>((total / 4 + 255) / 256, (long long)sm_count() * 8);
      splitk_reduce4_kernel<<<rg, 256, 0, s>>>((const float4*)a->workspace, splits, split_stride / 4, a->M, a->N / 4, a->alpha, (float*)a->D,
                                               a->ldd, a->accumulate);
    } else {
      const int rg = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 8);
      splitk_reduce_kernel<<<rg, 256, 0, s>>>((const float*)a->workspace, splits, split_stride, a->M, a->N, a->alpha, (float*)a->D, a->ldd,
                                              a->accumulate);
    }
  }
  BFK_CHECK_LAUNCH("gemm_tc");
  return BFK_OK;
}

template <int BN, bool AMN, bool BMN>
static int launch_tc(const bfk_gemm_args* a, cudaStream_t s, int* partial_splits = nullptr) {
  const int splits = choose_splits(a, BN);
  if (partial_splits) {
    *partial_splits = splits;
    return launch_tc_cl<BN, AMN, BMN, 1>(a, s, splits, true);
  }
  const int num_m = (a->M + TC_BM - 1) / TC_BM, num_n = (a->N + BN - 1) / BN;
  // cluster-of-2 B multicast pays off when every SM has work anyway; tiny problems keep single CTAs
  if (cluster_enabled() && splits == 1 && num_m >= 2 && num_m * num_n >= sm_count() / 2)
    return launch_tc_cl<BN, AMN, BMN, 2>(a, s, splits);
  return launch_tc_cl<BN, AMN, BMN, 1>(a, s, splits);
}

// wgrad-shaped product (A and B both stored [K, rows]) whose K-split partials stay in a->workspace; *splits = count
int gemm_tc_wgrad_partials(const bfk_gemm_args* a, cudaStream_t s, int* splits) {
  if (a->N > 128) return launch_tc<256, true, true>(a, s, splits);
  return launch_tc<128, true, true>(a, s, splits);
}

int gemm_tc(const bfk_gemm_args* a, cudaStream_t s) {
  // CTA-pair kernel (256 x 256 tiles) for every problem that fills the machine with them; a ragged last n-tile
  // (N = 784: 16 live columns of 256) is cheaper there than 128-wide tiles on the single-CTA kernel
  if (pair_enabled() && a->M >= 256 && a->N >= 512 && ((a->M + 255) / 256) * ((a->N + 255) / 256) >= sm_count() / 4 &&
      choose_splits(a, 256) == 1)
    return gemm_tc2(a, s);
  const int w256 = ((a->N + 255) / 256) * 256 - a->N, w128 = ((a->N + 127) / 128) * 128 - a->N;
  const bool bn256 = w256 <= w128 && a->N > 128;
  const bool ta = a->trans_a != 0, tb = a->trans_b != 0;
#define DISPATCH(BN)                                                  \
  do {                                                                \
    if (!ta && !tb) return launch_tc<BN, false, false>(a, s);         \
    if (!ta && tb) return launch_tc<BN, false, true>(a, s);           \
    if (ta && tb) return launch_tc<BN, true, true>(a, s);             \
    return launch_tc<BN, true, false>(a, s);                          \
  } while (0)
  if (bn256) DISPATCH(256);
  DISPATCH(128);
#undef DISPATCH
}

}  // namespace bfk
