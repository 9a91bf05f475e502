// pmi_tc_kernel.cuh -- PMI reciprocal reward on the 5th-generation tensor cores (tcgen05, sm_100a), H = 128.
//
// Same contract as uavsim_pmi_kernel (pmi_kernel.cuh): for every ordered neighbour pair (i, j) of an environment
// the folded PMINetwork (src/models/PMINet.py:41-62) maps la_i * la_j (12 values) to a logit; a float32 softmax
// over each UAV's neighbours mixes their raw rewards (src/agent/uav.py:262-291).
//
// Mapping.  One CTA (256 threads) per SM, persistent over groups of environments.  Pair rows are processed in
// tiles of 128 (threads r and r + 128 share row r = TMEM lane r and split every column range in two).  Both GEMMs of
// the MLP run on the tensor cores as tcgen05.mma.kind::tf32 (M = 128, N = 128, K = 8) issued by one thread, fp32
// accumulators in tensor memory (all 512 columns: 384 for layer 0, 128 for layer 1):
//   layer 0  [128 x 16] x [16 x 384]: the 12 inputs la_i * la_j, a constant-one column carrying the bias, 3 zero
//            columns; the block-diagonal branch structure (PMINet.py:45-58, BN folded) is zeros in the weight tile.
//   layer 1  [128 x 384] x [384 x 128] in 24 K-chunks of 16 hidden units: the threads read their part of the layer-0
//            accumulator (tcgen05.ld), apply ReLU, split, and write the chunk as the next A operand.
// Precision: single-pass TF32 (10-bit mantissa) misses the 1e-5 bar, so every operand is split
// x = hi + lo (hi = tf32(x), lo = tf32(x - hi)) and every K-slice runs three MMAs hi*hi + lo*hi + hi*lo --
// fp32-class products, fp32 accumulation.
// Operand layout: canonical K-major no-swizzle UMMA tiles, byte(r, k) = (k/4)*LBO + (r/8)*128 + (r%8)*16 + (k%4)*4
// (8 x 16-byte core matrices; LBO = 16 * rows).  Thread r writes one 16-byte vector per 4 K values, consecutive
// threads consecutive vectors (no bank conflicts).  The weights are pre-split and pre-tiled on the host: fc_* as one
// 48 KB block fetched once, fc1 as one 16 KB block (hi | lo) per K-chunk fetched with cp.async.bulk (TMA engine,
// mbarrier complete_tx) into one of two stages, so the MMAs of chunk c overlap the ReLU / split of chunk c+1;
// tcgen05.commit releases a stage.
// Epilogue: tcgen05.ld -> bias + ReLU -> dot with fc2 inside the two owning threads -> logit in shared memory;
// then softmax + mix per UAV.
#pragma once
#include "common.cuh"

#define TC_NT 256            // threads per CTA: 2 threads per pair row; warps w and w+4 share a TMEM lane quarter
#define TC_H 128
#define TC_H3 384
#define TC_KC 16             // hidden units per layer-1 K-chunk
#define TC_NCHUNK (TC_H3 / TC_KC)
#define TC_TILE_BYTES 8192   // one 128 x 16 fp32 operand tile
#define TC_W0_BYTES 24576    // one 384 x 16 fp32 weight tile of layer 0
#define TC_AMAX 512          // UAVs per environment group
#define TC_PMAX 8192         // neighbour pairs per environment group
#define TC_TMEM_COLS 512     // layer-1 accumulator at columns 0..127, layer-0 accumulator at 128..511

// instruction descriptor: D = F32, A = B = TF32, K-major both, N = 128 (>>3 at bit 17), M = 128 (>>4 at bit 24)
#define TC_IDESC ((1u << 4) | (2u << 7) | (2u << 10) | ((TC_H >> 3) << 17) | ((128u >> 4) << 24))

struct TcSmem {  // byte offsets inside dynamic shared memory
  static constexpr uint32_t stage = 0;                                   // 2 x (A_hi, A_lo, B_hi, B_lo)
  static constexpr uint32_t w0t = 2 * 4 * TC_TILE_BYTES;                 // layer-0 weight tiles (hi | lo)
  static constexpr uint32_t xt = w0t + 2 * TC_W0_BYTES;                  // layer-0 input tiles (hi | lo)
  static constexpr uint32_t obs = xt + 2 * TC_TILE_BYTES;                // float [AMAX*12]
  static constexpr uint32_t raw = obs + TC_AMAX * 12 * 4;                // double [AMAX]
  static constexpr uint32_t nbr = raw + TC_AMAX * 8;                     // uint64 [AMAX*2]
  static constexpr uint32_t off = nbr + TC_AMAX * 16;                    // uint32 [AMAX+4]
  static constexpr uint32_t logit = off + (TC_AMAX + 4) * 4;             // float [PMAX]
  static constexpr uint32_t part = logit + TC_PMAX * 4;                  // float [128] partial fc2 dots of the upper half
  static constexpr uint32_t b1 = part + 128 * 4;                         // float [128]
  static constexpr uint32_t w2 = b1 + TC_H * 4;                          // float [128]
  static constexpr uint32_t red = w2 + TC_H * 4;                         // double [64]
  static constexpr uint32_t bar = red + 64 * 8;                          // 7 mbarriers + tmem pointer
  static constexpr uint32_t total = bar + 64;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a barrier that never completes (a descriptor bug) traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; spin++) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
  // K-major, no swizzle: start address, LBO = byte distance between the two 16-byte K halves of an MMA (16 * rows of
  // the tile), SBO = 128 B (between 8-row groups), descriptor version 1 (Blackwell)
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) |
         (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t *u = reinterpret_cast<uint32_t *>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
  uint32_t *u = reinterpret_cast<uint32_t *>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// four fp32 values -> tf32 (hi, lo) vectors at the same position of two operand tiles
__device__ __forceinline__ void store_split4(unsigned char *hi_tile, unsigned char *lo_tile, uint32_t byte_off, float a,
                                             float b, float c, float d) {
  const float ha = tf32_rna(a), hb = tf32_rna(b), hc = tf32_rna(c), hd = tf32_rna(d);
  *reinterpret_cast<float4 *>(hi_tile + byte_off) = make_float4(ha, hb, hc, hd);
  *reinterpret_cast<float4 *>(lo_tile + byte_off) = make_float4(tf32_rna(a - ha), tf32_rna(b - hb), tf32_rna(c - hc), tf32_rna(d - hd));
}

struct PmiTcDev {
  const float *b1, *w2;            // [128] [128] folded fp32
  const float *w0_tiles;           // [2][384 x 16] layer 0 (inputs + bias column) pre-split (hi | lo) in the UMMA layout
  const float *w1_tiles;           // [24][2][128 x 16] fc1 pre-split (hi | lo) in the UMMA layout
  float b2;
};

__global__ void __launch_bounds__(TC_NT, 1)
uavsim_pmi_tc_kernel(const KParams P, const UavSimBuffers B, const PmiTcDev W, int64_t env_begin, int64_t env_count,
                     int G, double coop, double *__restrict__ stats_partial) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int n = P.n, tid = threadIdx.x, warp = tid >> 5, row = tid & 127, half = tid >> 7;
  float *s_obs = reinterpret_cast<float *>(smem + TcSmem::obs);
  double *s_raw = reinterpret_cast<double *>(smem + TcSmem::raw);
  uint64_t *s_nbr = reinterpret_cast<uint64_t *>(smem + TcSmem::nbr);
  uint32_t *s_off = reinterpret_cast<uint32_t *>(smem + TcSmem::off);
  float *s_logit = reinterpret_cast<float *>(smem + TcSmem::logit);
  float *s_part = reinterpret_cast<float *>(smem + TcSmem::part);
  float *s_b1 = reinterpret_cast<float *>(smem + TcSmem::b1);
  float *s_w2 = reinterpret_cast<float *>(smem + TcSmem::w2);
  double *s_red = reinterpret_cast<double *>(smem + TcSmem::red);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + TcSmem::bar + 56);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full0 = sbase + TcSmem::bar, bar_free0 = bar_full0 + 16, bar_acc = bar_full0 + 32,
                 bar_l0 = bar_full0 + 40, bar_w0 = bar_full0 + 48;
  unsigned char *x_hi = smem + TcSmem::xt, *x_lo = x_hi + TC_TILE_BYTES;

  for (int k = tid; k < TC_H; k += TC_NT) { s_b1[k] = W.b1[k]; s_w2[k] = W.w2[k]; }
  if (tid == 0) {
    mbar_init(bar_full0, 1); mbar_init(bar_full0 + 8, 1);
    mbar_init(bar_free0, 1); mbar_init(bar_free0 + 8, 1);
    mbar_init(bar_acc, 1); mbar_init(bar_l0, 1); mbar_init(bar_w0, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar_w0, 2 * TC_W0_BYTES);  // layer-0 weights stay resident for the whole kernel
    bulk_g2s(sbase + TcSmem::w0t, W.w0_tiles, 2 * TC_W0_BYTES, bar_w0);
  }
  if (warp == 0) {  // all of tensor memory, allocated by one warp (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *s_tmem;
  const uint32_t tmem_lane = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's lane quarter

  uint32_t use0 = 0, use1 = 0;  // how often each stage has been filled (phase bookkeeping)
  uint32_t tiles_done = 0;
  const int64_t ngroups = (env_count + G - 1) / G;
  double st_r = 0;

  for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int64_t e0 = env_begin + grp * G;
    const int ne = (int)min((int64_t)G, env_begin + env_count - e0);
    const int A = ne * n;  // UAVs in this group (<= TC_AMAX)
    __syncthreads();
    for (int k = tid; k < A * 12; k += TC_NT) s_obs[k] = B.obs[e0 * n * 12 + k];
    for (int a = tid; a < A; a += TC_NT) {
      s_raw[a] = B.raw[e0 * n + a];
      const uint64_t n0 = B.nbr_bits[(e0 * n + a) * 2], n1 = B.nbr_bits[(e0 * n + a) * 2 + 1];
      s_nbr[2 * a] = n0; s_nbr[2 * a + 1] = n1;
      s_off[a + 1] = __popcll(n0) + __popcll(n1);
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the neighbour counts (A <= 512): 16 per lane + warp scan
      const int per = (A + 31) / 32, lo = (tid & 31) * per;
      uint32_t sum = 0;
      for (int k = 0; k < per; k++) if (lo + k < A) sum += s_off[lo + k + 1];
      uint32_t inc = sum;
      for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31) >= o) inc += v; }
      uint32_t run = inc - sum;
      if (tid == 0) s_off[0] = 0;
      for (int k = 0; k < per; k++) if (lo + k < A) { run += s_off[lo + k + 1]; s_off[lo + k + 1] = run; }
    }
    __syncthreads();
    const int npairs = (int)s_off[A];

    for (int p0 = 0; p0 < npairs; p0 += 128) {
      // ---- this thread's pair row: flat index -> (UAV a, its k-th neighbour b), x = la_a * la_b (uav.py:280-281)
      float x[12];
      const int p = p0 + row;
      if (p < npairs) {
        int lo = 0, hi = A;  // largest a with off[a] <= p
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= (uint32_t)p) lo = mid; else hi = mid; }
        const int a = lo;
        int k = p - (int)s_off[a];
        uint64_t w = s_nbr[2 * a];
        int base = (a / n) * n;
        const int c0 = __popcll(w);
        if (k >= c0) { k -= c0; w = s_nbr[2 * a + 1]; base += 64; }
        for (int t = 0; t < k; t++) w &= w - 1;
        const int b = base + __ffsll((long long)w) - 1;
#pragma unroll
        for (int q = 0; q < 12; q++) x[q] = s_obs[a * 12 + q] * s_obs[b * 12 + q];
      } else {
#pragma unroll
        for (int q = 0; q < 12; q++) x[q] = 0.f;
      }

      // ---- layer 0 on the tensor cores: X = [x0..x11, 1, 0, 0, 0] (the one carries the bias), split, K-major tile
      {
        float xv[8];
        if (half == 0) {
#pragma unroll
          for (int q = 0; q < 8; q++) xv[q] = x[q];
        } else {
          xv[0] = x[8]; xv[1] = x[9]; xv[2] = x[10]; xv[3] = x[11];
          xv[4] = (p < npairs) ? 1.f : 0.f; xv[5] = 0.f; xv[6] = 0.f; xv[7] = 0.f;
        }
        store_split4(x_hi, x_lo, (uint32_t)(2 * half) * 2048u + row * 16u, xv[0], xv[1], xv[2], xv[3]);
        store_split4(x_hi, x_lo, (uint32_t)(2 * half + 1) * 2048u + row * 16u, xv[4], xv[5], xv[6], xv[7]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        if (tiles_done == 0) mbar_wait(bar_w0, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t xh = sbase + TcSmem::xt, xl = xh + TC_TILE_BYTES, wh = sbase + TcSmem::w0t, wl = wh + TC_W0_BYTES;
#pragma unroll
        for (int nb = 0; nb < 3; nb++) {      // 384 hidden units = three N = 128 blocks
          const uint32_t d0 = tmem_d + 128u + 128u * nb;
#pragma unroll
          for (int ks = 0; ks < 2; ks++) {    // K = 16 = two MMAs of K = 8
            const uint32_t ao = (uint32_t)ks * 2u * 2048u, bo = (uint32_t)nb * 2048u + (uint32_t)ks * 2u * 6144u;
            umma_tf32(d0, umma_desc(xh + ao, 2048), umma_desc(wh + bo, 6144), ks ? 1u : 0u);
            umma_tf32(d0, umma_desc(xl + ao, 2048), umma_desc(wh + bo, 6144), 1u);
            umma_tf32(d0, umma_desc(xh + ao, 2048), umma_desc(wl + bo, 6144), 1u);
          }
        }
        umma_commit(bar_l0);
      }
      mbar_wait(bar_l0, tiles_done & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

      // ---- layer 1: 24 K-chunks of 16 hidden units through the two stages
#pragma unroll 1
      for (int c = 0; c < TC_NCHUNK; c++) {
        const int s = c & 1;
        const uint32_t stage = sbase + TcSmem::stage + (uint32_t)s * 4u * TC_TILE_BYTES;
        const uint32_t use = s ? use1 : use0;
        if (use > 0) mbar_wait(bar_free0 + 8 * s, (use - 1) & 1);  // MMAs that read this stage have completed
        if (tid == 0) {
          mbar_expect_tx(bar_full0 + 8 * s, 2 * TC_TILE_BYTES);
          bulk_g2s(stage + 2 * TC_TILE_BYTES, W.w1_tiles + (size_t)c * (2 * TC_TILE_BYTES / 4), 2 * TC_TILE_BYTES, bar_full0 + 8 * s);
        }
        // this thread's 8 of the chunk's 16 hidden units: layer-0 accumulator -> ReLU -> split -> A operand
        float v[8];
        tmem_ld8(tmem_lane + 128u + (uint32_t)(c * TC_KC + 8 * half), v);
        unsigned char *a_hi = smem + TcSmem::stage + (size_t)s * 4 * TC_TILE_BYTES, *a_lo = a_hi + TC_TILE_BYTES;
        store_split4(a_hi, a_lo, (uint32_t)(2 * half) * 2048u + row * 16u, fmaxf(v[0], 0.f), fmaxf(v[1], 0.f), fmaxf(v[2], 0.f), fmaxf(v[3], 0.f));
        store_split4(a_hi, a_lo, (uint32_t)(2 * half + 1) * 2048u + row * 16u, fmaxf(v[4], 0.f), fmaxf(v[5], 0.f), fmaxf(v[6], 0.f), fmaxf(v[7], 0.f));
        if (s) use1 = use + 1; else use0 = use + 1;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
          mbar_wait(bar_full0 + 8 * s, use & 1);  // fc1 chunk has landed
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ah = stage, al = stage + TC_TILE_BYTES, bh = stage + 2 * TC_TILE_BYTES, bl = stage + 3 * TC_TILE_BYTES;
#pragma unroll
          for (int ks = 0; ks < TC_KC / 8; ks++) {  // one MMA consumes K = 8 = two 16-byte core-matrix columns
            const uint32_t o = (uint32_t)ks * 2u * 2048u;
            umma_tf32(tmem_d, umma_desc(ah + o, 2048), umma_desc(bh + o, 2048), (c | ks) ? 1u : 0u);
            umma_tf32(tmem_d, umma_desc(al + o, 2048), umma_desc(bh + o, 2048), 1u);
            umma_tf32(tmem_d, umma_desc(ah + o, 2048), umma_desc(bl + o, 2048), 1u);
          }
          umma_commit(bar_free0 + 8 * s);                  // stage reusable when these MMAs are done
          if (c == TC_NCHUNK - 1) umma_commit(bar_acc);    // accumulator complete
        }
      }

      // ---- epilogue: bias + ReLU + fc2 (PMINet.py:59-62), one accumulator row per thread
      mbar_wait(bar_acc, tiles_done & 1);
      tiles_done++;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float part = 0.f;
#pragma unroll 1
      for (int cb = half * 64; cb < half * 64 + 64; cb += 32) {  // this thread's half of the 128 columns
        float v[32];
        tmem_ld32(tmem_lane + (uint32_t)cb, v);
#pragma unroll
        for (int k = 0; k < 32; k++) part = fmaf(s_w2[cb + k], fmaxf(v[k] + s_b1[cb + k], 0.f), part);
      }
      if (half) s_part[row] = part;
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (!half && p < npairs) s_logit[p] = (part + s_part[row]) + W.b2;
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();  // every row has left tensor memory before the next tile overwrites it
    }
    __syncthreads();

    // ---- softmax over each UAV's neighbours (scipy.special.softmax on float32) and the mix (uav.py:284-290)
    for (int a = tid; a < A; a += TC_NT) {
      const uint32_t lo = s_off[a], hi = s_off[a + 1];
      const double raw = s_raw[a];
      double r;
      if (hi > lo) {
        float mx = s_logit[lo];
        for (uint32_t q = lo + 1; q < hi; q++) mx = fmaxf(mx, s_logit[q]);
        float ssum = 0.f;
        for (uint32_t q = lo; q < hi; q++) ssum += expf(s_logit[q] - mx);
        double acc = 0;
        uint32_t q = lo;
        const int base = (a / n) * n;
        for (int half = 0; half < 2; half++) {
          uint64_t w = s_nbr[2 * a + half];
          while (w) {
            const int j = __ffsll((long long)w) - 1;
            w &= w - 1;
            const float wgt = expf(s_logit[q] - mx) / ssum;
            acc += s_raw[base + 64 * half + j] * (double)wgt;
            q++;
          }
        }
        r = (1 - coop) * raw + coop * acc;
      } else {
        r = (1 - coop) * raw;
      }
      r = fmin(fmax(r, -1.0), 1.0);
      B.rew4[e0 * n + a] = (float)r;
      st_r += r;
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)TC_TMEM_COLS) : "memory");
  block_stats_commit(s_red, stats_partial + (size_t)blockIdx.x * STAT_W, st_r, 0, 0, 0, 0, 0, 0, TC_NT);
}
