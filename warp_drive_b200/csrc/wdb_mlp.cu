// wdb_mlp.cu -- fused policy/value MLP forward on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Replaces, for the rollout path, FullyConnected.forward of the reference
// (warp_drive/training/models/fully_connected.py:51-89): obs -> Linear+ReLU -> Linear+ReLU ->
// {head0 softmax, head1 softmax, value}.  The reference (and the torch fallback) runs it as
// 5 GEMMs + 7 elementwise kernels that round-trip the [rows, 256] activations through HBM;
// here one persistent kernel keeps them on chip:
//
//   * all weights (bf16, pre-packed by wdb_mlp_pack_weights into the UMMA canonical K-major
//     no-swizzle layout) are TMA-bulk-loaded into shared memory ONCE per CTA (192 KB);
//   * per 128-row tile: obs (fp32) -> bf16 -> TMEM (through a shared-memory transpose);
//     layer 1 = tcgen05.mma (A from TMEM) into TMEM; epilogue (tcgen05.ld, +bias, ReLU,
//     bf16 pack) writes the hidden activations BACK INTO TMEM (tcgen05.st); layers 2 and 3
//     take their A operand straight from TMEM (tcgen05.mma .ts form) -- hidden activations
//     never touch shared or global memory; final epilogue = bias + two softmaxes + value,
//     staged in shared memory and written with TMA bulk stores.
//   * warp-specialised: lane 0 of a dedicated warp issues every tcgen05.mma (issue is
//     back-pressured by the tensor core, ~100 cycles per instruction); 12 worker warps
//     (4 TMEM lane quadrants x 3 column parts) run the obs path and the epilogues.
//     Workers -> issuer: named barriers (bar.arrive / bar.sync); issuer -> workers:
//     tcgen05.commit onto one mbarrier per layer;
//   * software pipeline: layer 1 of tile t+1 is issued right behind layer 3 of tile t; the
//     next tile's A operand is built, and the PREVIOUS tile's softmax epilogue runs, while
//     the tensor core executes layer 2; the obs of tile t+2 are prefetched into registers
//     while it executes layer 3;
//   * optional TILES mode: the A operand arrives as ready-made bf16 tiles (one TMA bulk
//     copy per tile) instead of fp32 obs rows.
//
// Numerics: bf16 operands, fp32 accumulation (the reference is fp32; SURVEY.md section 8
// row a3 allows TF32/BF16 for the forward that feeds the sampler).  Tested against a
// float32 torch reference in tests/test_gpu_mlp.py (probabilities within 2e-2 abs, sums to 1).
#include <cuda_bf16.h>
#include <math_constants.h>

#include "wdb_common.cuh"

using namespace wdb;

namespace {

constexpr int kTileM = 128;
constexpr int kMlpWeightsStable = 1;   // wdb_mlp_policy_forward_pair flags
__device__ __forceinline__ void griddep_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#ifndef WDB_MLP_WORKERS
#define WDB_MLP_WORKERS 384
#endif
constexpr int kWorkers = WDB_MLP_WORKERS;   // worker warps x 32 (obs path + epilogues); multiple of 128
constexpr int kThreads = kWorkers + 32;  // + 1 warp whose lane 0 issues every tcgen05.mma
constexpr int kWarps = kWorkers / 32;
constexpr int kParts = kWarps / 4;       // worker warps per TMEM lane quadrant
constexpr int kTasks = (48 + kWorkers / 32 - 1) / (kWorkers / 32);       // A-tile build tasks (8 fp32 obs values each) per thread held across a tile
constexpr int kTmemCols = 512;
constexpr int kColD = 0;      // accumulator columns [0, 256)
constexpr int kColH = 256;    // packed bf16 hidden activations [256, 384)
constexpr int kColA = 384;    // packed bf16 obs of the current / next tile [384, 384 + K1/2)
// layer-3 accumulator: behind the obs columns when it fits (then layer 1 of the next tile may
// run during the output epilogue), else on top of D

// Profiling aid (-DWDB_PHASE_CLOCKS): thread 0 of CTA 0 records the SM clock at the phase
// boundaries of its first tiles; read back with wdb_debug_mlp_clocks (debug builds only).
#ifdef WDB_PHASE_CLOCKS
__device__ long long g_mlp_clk[96];
#define MLP_MARK(i)                                                                  \
  if (blockIdx.x == 0 && threadIdx.x == 0 && mark_tile < 6) g_mlp_clk[mark_tile * 16 + (i)] = clock64() - mlp_t0;
#else
#define MLP_MARK(i)
#endif

struct MlpHeader {            // start of the packed weight blob (device memory)
  int F, K1, H, A0, A1, N3;   // input features, padded K of layer 1, hidden width, heads, padded N3
  int off_w1, off_w2, off_w3, off_b, total_bytes, pad;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// UMMA shared-memory descriptor, K-major, no swizzle ("interleaved" 8x16B core matrices):
// element (r, k) of a [rows, K] bf16 tile lives at
//   (r / 8) * SBO + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2     [LBO = 128]
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;     // descriptor version (sm_100)
  return d;                   // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}

__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4)            // D format  = F32
         | (1u << 7)          // A format  = BF16
         | (1u << 10)         // B format  = BF16
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // K-major A and B
}

__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; "
        "selp.u32 %0, 1, 0, p; }"
        : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load(uint32_t dst, const void *src, uint32_t bytes,
                                         uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void tma_store(void *dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),
        "r"(v[14]), "r"(v[15]) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&h);
}

// named barriers: 1 = workers only, 2 = workers arrive / issuer warp waits
__device__ __forceinline__ void bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int n) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// predicated read-only load (0 when off) -- written in PTX so that no select depends on the
// loaded value: the prefetch must not wait for its own loads
__device__ __forceinline__ float ldg_if(const float *p, bool on) {
  float v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t"
               "@p ld.global.nc.f32 %0, [%1];\n\t}"
               : "=f"(v) : "l"(p), "r"((int)on));
  return v;
}

// (lo + b_lo, hi + b_hi) -> ReLU -> packed bf16x2: one FADD2 + one F2FP.RELU per 2 columns
__device__ __forceinline__ uint32_t bias_relu_pack(uint32_t lo, uint32_t hi, float b_lo,
                                                   float b_hi) {
  uint32_t out;
  asm("{\n\t"
      ".reg .b64 v, b, s;\n\t"
      ".reg .f32 x, y;\n\t"
      "mov.b64 v, {%1, %2};\n\t"
      "mov.b64 b, {%3, %4};\n\t"
      "add.rn.f32x2 s, v, b;\n\t"
      "mov.b64 {x, y}, s;\n\t"
      "cvt.rn.relu.bf16x2.f32 %0, y, x;\n\t"
      "}"
      : "=r"(out) : "r"(lo), "r"(hi), "f"(b_lo), "f"(b_hi));
  return out;
}

// hidden epilogue: D[lane, c0..c1) (+bias, ReLU) -> packed bf16 into TMEM columns kColH + c/2
__device__ __forceinline__ void hidden_epilogue(uint32_t tmem_lane_base, uint32_t col_h,
                                                const float *bias, int c0, int c1) {
  for (int c = c0; c < c1; c += 32 * kParts) {   // the warps of a quadrant interleave 32-column chunks
    uint32_t v[32];
    tmem_ld32(tmem_lane_base + kColD + c, v);
    uint32_t out[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float4 b = *reinterpret_cast<const float4 *>(bias + c + 4 * i);   // broadcast LDS.128
      out[2 * i] = bias_relu_pack(v[4 * i], v[4 * i + 1], b.x, b.y);
      out[2 * i + 1] = bias_relu_pack(v[4 * i + 2], v[4 * i + 3], b.z, b.w);
    }
    tmem_st16(tmem_lane_base + col_h + c / 2, out);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 16 warps: warp w owns TMEM lanes 32 (w & 3) .. +31 (= rows of the tile) and, in the hidden
// epilogues, every kParts-th 32-column chunk (part = w >> 2); in the output epilogue part 0
// does head 0 and part 1 head 1 + the value.  (Two warps per scheduler left every dependent
// instruction latency exposed; four hide most of it.)  The A operand of layer 1 also lives in TMEM: the NEXT tile's obs
// (prefetched into registers one tile ahead) are converted and copied there while the tensor
// core runs layer 2 of the current tile, so the obs path never sits on the critical path.
// TILES = false: `input` is the fp32 obs [rows, F]; TILES = true: `input` is the bf16 copy of
// the obs already in A-operand layout (128-row tiles of the canonical K-major layout, K
// padded with zeros -- written by wdb_mlp_pack_obs or by the env-step kernel's epilogue): a
// tile is then ONE 16-byte-aligned contiguous block that the TMA drops into the staging
// buffer, and the whole fp32 obs path (4-byte loads, conversion, scatter) disappears.
// `bid` / `gdim`: this CTA's index and the CTA count of ITS policy (the pair kernel runs two
// policies side by side in one grid).  flags & kMlpWeightsStable: the weight blob was not
// written by the grid this launch programmatically depends on, so its TMA load may start
// before griddepcontrol.wait (see the PDL note at launch_forward_pair).
template <bool TILES>
__device__ __forceinline__ void
mlp_forward_body(const unsigned char *__restrict__ blob, const void *__restrict__ input,
                 long long rows, float *__restrict__ probs0, float *__restrict__ probs1,
                 float *__restrict__ values, const unsigned bid, const unsigned gdim,
                 const int flags) {
  const float *obs = reinterpret_cast<const float *>(input);
  const unsigned char *tiles = reinterpret_cast<const unsigned char *>(input);
  extern __shared__ __align__(128) unsigned char smem[];
#ifdef WDB_PHASE_CLOCKS
  const long long mlp_t0 = clock64();
  int mark_tile = 0;
#endif
  const MlpHeader hd = *reinterpret_cast<const MlpHeader *>(blob);
  const int F = hd.F, K1 = hd.K1, H = hd.H, A0 = hd.A0, A1 = hd.A1, N3 = hd.N3;
  const int w_bytes = hd.total_bytes - hd.off_w1;        // W1 | W2 | W3 | biases, contiguous
  unsigned char *s_w = smem;                             // packed weights + biases
  const unsigned char *s_w2 = s_w + (hd.off_w2 - hd.off_w1);
  const unsigned char *s_w3 = s_w + (hd.off_w3 - hd.off_w1);
  const float *s_b1 = reinterpret_cast<const float *>(s_w + (hd.off_b - hd.off_w1));
  const float *s_b2 = s_b1 + H;
  const float *s_b3 = s_b2 + H;
  unsigned char *s_a = s_w + ((w_bytes + 127) & ~127);   // A tile (layer 1) / output staging
  const int a_bytes = max(kTileM * K1 * 2, kTileM * (A0 + A1 + 1) * 4);
  unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(s_a + ((a_bytes + 15) & ~15));
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const uint32_t bar_w = smem_addr(&s_bar[0]);           // weights landed
  const uint32_t bar_l1 = smem_addr(&s_bar[1]);          // layer-1 / 2 / 3 accumulators complete
  const uint32_t bar_l2 = smem_addr(&s_bar[2]);
  const uint32_t bar_l3 = smem_addr(&s_bar[3]);
  const uint32_t bar_a = smem_addr(&s_bar[4]);           // TILES: next A tile landed in the staging buffer

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_l1, 1);
    mbar_init(bar_l2, 1);
    mbar_init(bar_l3, 1);
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_addr(s_tmem)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16);   // this warp's lanes

  // Programmatic dependent launch: everything above overlapped the tail of the previous grid
  // in the stream.  Nothing that grid wrote is read before griddepcontrol.wait; our own
  // dependents are released only AFTER our wait, so "previous grid complete" is transitive.
  if (!(flags & kMlpWeightsStable)) griddep_wait();
  if (tid == 0) {                                         // weights: TMA bulk copies
    mbar_expect_tx(bar_w, (uint32_t)w_bytes);
    uint32_t off = 0;
    while (off < (uint32_t)w_bytes) {
      const uint32_t n = min((uint32_t)w_bytes - off, 65536u);
      tma_load(smem_addr(s_w + off), blob + hd.off_w1 + off, n, bar_w);
      off += n;
    }
  }

  if (flags & kMlpWeightsStable) griddep_wait();
  griddep_launch_dependents();

  const long long n_tiles = (rows + kTileM - 1) / kTileM;
  // A-tile build: a warp task = 8 rows x 4 k-chunks (chunk = 8 consecutive k = one 16-byte
  // row of a core matrix): lane -> (row = 8 g + lane % 8, chunk = 4 q + lane / 8), so the 8
  // global loads of a task read 128 contiguous bytes of each of 8 obs rows and the one
  // 16-byte shared store per lane is bank-conflict free.  16 * ncg tasks per tile, dealt
  // round-robin to the warps; up to kTasks per thread live in registers (pf).
  const int nchunk = K1 / 8, ncg = (nchunk + 3) / 4;
  const int tasks_per_warp = (16 * ncg + kWarps - 1) / kWarps;
  const bool single_batch = tasks_per_warp <= kTasks;     // whole tile fits the register prefetch
  float pf[kTasks * 8];
  auto load_batch = [&](long long tile, int batch) {
    const long long r0 = tile * kTileM;
    const int valid = (int)min((long long)kTileM, rows - r0);
#pragma unroll
    for (int j = 0; j < kTasks; j++) {
      const int wt = warp + kWarps * (j + kTasks * batch);
      const int g = wt / ncg, q = wt - g * ncg;
      const int r = 8 * g + (lane & 7), c = 4 * q + (lane >> 3);
      const bool on = (g < 16) && (c < nchunk) && (r < valid);
      const float *src = obs + (r0 + r) * F + 8 * c;
#pragma unroll
      for (int i = 0; i < 8; i++) pf[8 * j + i] = ldg_if(src + i, on && 8 * c + i < F);
    }
  };
  auto store_batch = [&](int batch) {
#pragma unroll
    for (int j = 0; j < kTasks; j++) {
      const int wt = warp + kWarps * (j + kTasks * batch);
      const int g = wt / ncg, q = wt - g * ncg;
      const int c = 4 * q + (lane >> 3);
      if (g < 16 && c < nchunk) {
        uint4 o;
        o.x = pack_bf16(pf[8 * j + 0], pf[8 * j + 1]);
        o.y = pack_bf16(pf[8 * j + 2], pf[8 * j + 3]);
        o.z = pack_bf16(pf[8 * j + 4], pf[8 * j + 5]);
        o.w = pack_bf16(pf[8 * j + 6], pf[8 * j + 7]);
        *reinterpret_cast<uint4 *>(s_a + g * (K1 / 8) * 128 + c * 128 + (lane & 7) * 16) = o;
      }
    }
  };
  if (!TILES && single_batch && warp < kWarps && (long long)bid < n_tiles) load_batch(bid, 0);

  const uint32_t idesc_h = make_idesc(kTileM, H);
  const uint32_t idesc_o = make_idesc(kTileM, N3);
  const uint32_t sbo_a = (uint32_t)(K1 / 8) * 128u;       // W1: K1/8 core matrices per row group
  const uint32_t sbo_h = (uint32_t)(H / 8) * 128u;        // W2 / W3: H/8 core matrices per row group
  const bool overlap_l1 = K1 / 2 + N3 <= kTmemCols - kColA;   // room for a separate layer-3 accumulator
  const int col_d3 = overlap_l1 ? kColA + K1 / 2 : kColD;
  float *s_p0 = reinterpret_cast<float *>(s_a);
  float *s_p1 = s_p0 + kTileM * A0;
  float *s_v = s_p1 + kTileM * A1;

  if (warp == kWarps) {
    // ================= MMA issuer warp =================
    // Waits (named barrier 2) until the workers have produced an operand in TMEM, then lane 0
    // issues the layer's MMAs and commits them to that layer's mbarrier.  Issuing is
    // back-pressured by the tensor core (~100+ cycles per instruction), which is why it has
    // its own warp: the workers build the next tile's operand meanwhile.
    auto issue_l1 = [&]() {
      const uint32_t b0 = smem_addr(s_w);
      for (int kk = 0; kk < K1 / 16; kk++)
        mma_ts(tmem_base + kColD, tmem_base + kColA + kk * 8, make_desc(b0 + kk * 256, sbo_a),
               idesc_h, kk > 0);
      mma_commit(bar_l1);
    };
    mbar_wait(bar_w, 0);
    bar_sync(2, kThreads);                                // first A operand in TMEM
    fence_after();
    if (lane == 0 && (long long)bid < n_tiles) issue_l1();
    for (long long tile = bid; tile < n_tiles; tile += gdim) {
      const bool has_next = tile + gdim < n_tiles;
      bar_sync(2, kThreads);                              // hidden activations of layer 1 packed
      fence_after();
      if (lane == 0) {
        const uint32_t b0 = smem_addr(s_w2);
        for (int kk = 0; kk < H / 16; kk++)
          mma_ts(tmem_base + kColD, tmem_base + kColH + kk * 8, make_desc(b0 + kk * 256, sbo_h),
                 idesc_h, kk > 0);
        mma_commit(bar_l2);
      }
      bar_sync(2, kThreads);                              // layer 2 packed; next A operand built
      fence_after();
      if (lane == 0) {
        const uint32_t b0 = smem_addr(s_w3);
        for (int kk = 0; kk < H / 16; kk++)
          mma_ts(tmem_base + col_d3, tmem_base + kColH + kk * 8, make_desc(b0 + kk * 256, sbo_h),
                 idesc_o, kk > 0);
        mma_commit(bar_l3);
      }
      if (!overlap_l1) {
        bar_sync(2, kThreads);                            // output epilogue has read D
        fence_after();
      }
      if (lane == 0 && has_next) issue_l1();              // runs during the output epilogue
    }
  } else {
    // ================= worker warps =================
    // A operand of layer 1 = packed bf16 obs in TMEM columns kColA.. (lane = row): registers
    // (fp32 obs prefetched one tile ahead) -> bf16 canonical tile in the shared staging buffer
    // -> each thread copies its row's 16-byte chunks into TMEM (tcgen05.st).
    const uint32_t tile_bytes = (uint32_t)kTileM * K1 * 2;
    uint32_t a_phase = 0;
    // TILES: thread 0 requests tile `t` (the staging buffer must be free: the previous output
    // store has read it, and every worker is past its reads -- see the end of the tile loop)
    auto request_tile = [&](long long t) {
      if (tid == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_expect_tx(bar_a, tile_bytes);
        tma_load(smem_addr(s_a), tiles + t * tile_bytes, tile_bytes, bar_a);
      }
    };
    auto build_a = [&](long long tile) {
      if (TILES) {
        mbar_wait(bar_a, a_phase); a_phase ^= 1;          // the TMA wrote the canonical tile
        const int row = quad * 32 + lane;
        const unsigned char *rp = s_a + (row >> 3) * (K1 / 8) * 128 + (row & 7) * 16;
        for (int c = part; c < nchunk; c += kParts) {
          const uint4 v = *reinterpret_cast<const uint4 *>(rp + c * 128);
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                       ::"r"(tmem_lane + kColA + 4 * c), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                       : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        return;
      }
      // the staging buffer may still be read by the previous tile's output TMA store
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      bar_sync(1, kWorkers);
      if (single_batch) {
        store_batch(0);
      } else {
        for (int batch = 0; batch * kTasks < tasks_per_warp; batch++) {
          load_batch(tile, batch);
          store_batch(batch);
        }
      }
      bar_sync(1, kWorkers);
      const int row = quad * 32 + lane;
      const unsigned char *rp = s_a + (row >> 3) * (K1 / 8) * 128 + (row & 7) * 16;
      for (int c = part; c < nchunk; c += kParts) {
        const uint4 v = *reinterpret_cast<const uint4 *>(rp + c * 128);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                     ::"r"(tmem_lane + kColA + 4 * c), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                     : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    };

    if ((long long)bid < n_tiles) {
      if (TILES) request_tile(bid);
      build_a(bid);
      if (!TILES && single_batch && (long long)bid + gdim < n_tiles)
        load_batch((long long)bid + gdim, 0);
    }
    fence_before();
    bar_arrive(2, kThreads);                              // first A operand ready
    mbar_wait(bar_w, 0);                                  // biases live in the weight blob

    // ---- output epilogue of tile `tl` (mbarrier parity `ph`): bias, softmax of this part's
    // head (+ value) -> staging -> TMA bulk store.  The caller guarantees that the staging
    // buffer is free (previous store read, no worker still reading it).
    auto output_epilogue = [&](long long tl, uint32_t ph) {
      const long long r0 = tl * kTileM;
      const int valid = (int)min((long long)kTileM, rows - r0);
      mbar_wait(bar_l3, ph);
      fence_after();
      MLP_MARK(7)
      if (part < 2) {
        const int cbase = part ? A0 : 0, cnt = part ? A1 : A0;
        const int row = quad * 32 + lane;                  // TMEM lane == row of the tile
        uint32_t v[32];
        tmem_ld32(tmem_lane + col_d3 + cbase, v);          // columns past N3 hold stale data: unused
        // logits pre-scaled by log2(e): softmax = 2^(l - max) / sum.  Branch-free over 32
        // statically indexed registers (predicated), four independent max / sum chains.
        constexpr float kLog2e = 1.4426950408889634f;
        float lg[32];
        float mx[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
        float value = 0.0f;
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const float x = __uint_as_float(v[i]) + s_b3[min(cbase + i, N3 - 1)];
          lg[i] = i < cnt ? x * kLog2e : -CUDART_INF_F;
          mx[i & 3] = fmaxf(mx[i & 3], lg[i]);
          value = (i == cnt) ? x : value;
        }
        const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        float zz[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 32; i++) {
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(lg[i]) : "f"(lg[i] - m));   // 2^-inf = 0
          zz[i & 3] += lg[i];
        }
        const float inv = 1.0f / ((zz[0] + zz[1]) + (zz[2] + zz[3]));
        float *dst = (part ? s_p1 : s_p0) + row * cnt;
#pragma unroll
        for (int i = 0; i < 32; i++)
          if (i < cnt) dst[i] = lg[i] * inv;
        if (part) s_v[row] = value;
      }
      if (!overlap_l1) {
        fence_before();
        bar_arrive(2, kThreads);                           // D may be overwritten by the next layer 1
      }
      MLP_MARK(8)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      bar_sync(1, kWorkers);
      float *g0 = probs0 + r0 * A0, *g1 = probs1 + r0 * A1;
      const uint32_t n0 = (uint32_t)valid * A0 * 4, n1 = (uint32_t)valid * A1 * 4;
      const bool tma_ok = (((uintptr_t)g0 | (uintptr_t)g1 | n0 | n1) & 15) == 0 &&
                          ((smem_addr(s_p0) | smem_addr(s_p1)) & 15) == 0;
      if (tma_ok) {
        if (tid == 0) {                                    // its completion is awaited before the
          tma_store(g0, smem_addr(s_p0), n0);              // staging buffer is written again
          if (n1) tma_store(g1, smem_addr(s_p1), n1);      // (single-head policies: A1 == 0)
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        for (int i = tid; i < valid * A0; i += kWorkers) g0[i] = s_p0[i];
        for (int i = tid; i < valid * A1; i += kWorkers) g1[i] = s_p1[i];
      }
      if (values && tid < valid) values[r0 + tid] = s_v[tid];
      MLP_MARK(10)
    };

    // With a separate layer-3 accumulator the output epilogue of tile t is DEFERRED into the
    // layer-2 window of tile t+1: the long softmax runs while the tensor core is busy with the
    // 16 biggest MMAs instead of extending the critical path.
    const bool deferred = overlap_l1;
    uint32_t phase = 0;
    long long prev_tile = -1;
    for (long long tile = bid; tile < n_tiles; tile += gdim, phase ^= 1) {
      const bool has_next = tile + gdim < n_tiles;
      MLP_MARK(0)   // tile start
      if (TILES && has_next) {
        bar_sync(1, kWorkers);                             // everyone is done with the staging buffer
        request_tile(tile + gdim);                    // lands during epilogue 1 / layer 2
      }

      // ---- layer 1 epilogue
      mbar_wait(bar_l1, phase);
      fence_after();
      MLP_MARK(1)
      hidden_epilogue(tmem_lane, kColH, s_b1, 32 * part, H);
      fence_before();
      bar_arrive(2, kThreads);
      MLP_MARK(2)

      // ---- while the tensor core runs layer 2: next tile's A operand, previous tile's output
      if (has_next) build_a(tile + gdim);
      MLP_MARK(3)
      if (deferred && prev_tile >= 0) {
        if (!has_next && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        bar_sync(1, kWorkers);                             // staging: store read, A copy finished
        output_epilogue(prev_tile, phase ^ 1);
      }

      // ---- layer 2 epilogue (layer 2 is done with H1: packed in place)
      mbar_wait(bar_l2, phase);
      fence_after();
      MLP_MARK(5)
      hidden_epilogue(tmem_lane, kColH, s_b2, 32 * part, H);
      fence_before();
      bar_arrive(2, kThreads);
      MLP_MARK(6)
      // (the ~12k loads of a tile take a while to queue: do it while layer 3 runs)
      if (!TILES && has_next && single_batch && tile + 2 * (long long)gdim < n_tiles)
        load_batch(tile + 2 * (long long)gdim, 0);
      MLP_MARK(4)

      if (!deferred) {
        if (!has_next && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (!has_next) bar_sync(1, kWorkers);             // (build_a did this when there is a next tile)
        output_epilogue(tile, phase);
      }
      prev_tile = tile;
#ifdef WDB_PHASE_CLOCKS
      mark_tile++;
#endif
    }
    if (deferred && prev_tile >= 0) {
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      bar_sync(1, kWorkers);
      output_epilogue(prev_tile, phase ^ 1);
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }

  fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

template <bool TILES>
__global__ void __launch_bounds__(kThreads, 1)
mlp_forward_kernel(const unsigned char *__restrict__ blob, const void *__restrict__ input,
                   long long rows, float *__restrict__ probs0, float *__restrict__ probs1,
                   float *__restrict__ values, int flags) {
  mlp_forward_body<TILES>(blob, input, rows, probs0, probs1, values, blockIdx.x, gridDim.x, flags);
}

// Two policies in ONE grid: CTAs [0, split) run policy A, the rest policy B (each persistent
// over its own tiles).  Replaces the fork / join of two launches on two streams.
struct MlpPairArgs {
  const unsigned char *blob[2];
  const float *obs[2];
  long long rows[2];
  float *probs0[2], *probs1[2], *values[2];
  unsigned split;
  int flags;
};
__global__ void __launch_bounds__(kThreads, 1)
mlp_forward_pair_kernel(const __grid_constant__ MlpPairArgs a) {
  const int w = blockIdx.x >= a.split ? 1 : 0;
  mlp_forward_body<false>(a.blob[w], a.obs[w], a.rows[w], a.probs0[w], a.probs1[w], a.values[w],
                          w ? blockIdx.x - a.split : blockIdx.x,
                          w ? gridDim.x - a.split : a.split, a.flags);
}

// ---- weight packing: fp32 nn.Linear weights [out, in] -> bf16 canonical K-major tiles ------
__global__ void pack_matrix_kernel(const float *__restrict__ w, int n_rows, int n_cols,
                                   int n_pad, int k_pad, __nv_bfloat16 *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;    // over n_pad * k_pad
  if (i >= n_pad * k_pad) return;
  const int n = i / k_pad, k = i - n * k_pad;
  const float v = (n < n_rows && k < n_cols) ? w[(long long)n * n_cols + k] : 0.0f;
  const int off = (n >> 3) * (k_pad / 8) * 64 + (k >> 3) * 64 + (n & 7) * 8 + (k & 7);
  out[off] = __float2bfloat16_rn(v);
}

__global__ void pack_rows_kernel(const float *__restrict__ w, int rows, int n_cols, int row0,
                                 int k_pad, __nv_bfloat16 *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * n_cols) return;
  const int rr = i / n_cols, k = i - rr * n_cols;
  const int n = row0 + rr;
  const int off = (n >> 3) * (k_pad / 8) * 64 + (k >> 3) * 64 + (n & 7) * 8 + (k & 7);
  out[off] = __float2bfloat16_rn(w[i]);
}

__global__ void pack_misc_kernel(MlpHeader hd, unsigned char *blob, const float *b1,
                                 const float *b2, const float *bh0, const float *bh1,
                                 const float *bv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *reinterpret_cast<MlpHeader *>(blob) = hd;
  float *b = reinterpret_cast<float *>(blob + hd.off_b);
  if (i < hd.H) { b[i] = b1[i]; b[hd.H + i] = b2[i]; }
  if (i < hd.N3) {
    float v = 0.0f;
    if (i < hd.A0) v = bh0[i];
    else if (i < hd.A0 + hd.A1) v = bh1[i - hd.A0];
    else if (i == hd.A0 + hd.A1) v = bv[0];
    b[2 * hd.H + i] = v;
  }
}

MlpHeader make_header(int F, int H, int A0, int A1) {
  MlpHeader hd;
  hd.F = F; hd.K1 = round_up(F, 16); hd.H = H; hd.A0 = A0; hd.A1 = A1;
  hd.N3 = round_up(A0 + A1 + 1, 16);
  hd.off_w1 = 128;
  hd.off_w2 = hd.off_w1 + H * hd.K1 * 2;
  hd.off_w3 = hd.off_w2 + H * H * 2;
  hd.off_b = hd.off_w3 + hd.N3 * H * 2;
  hd.total_bytes = round_up(hd.off_b + (2 * H + hd.N3) * 4, 16);
  hd.pad = 0;
  return hd;
}

size_t mlp_smem_bytes(const MlpHeader &hd) {
  const size_t w = (size_t)hd.total_bytes - hd.off_w1;
  const size_t a = (size_t)max(kTileM * hd.K1 * 2, kTileM * (hd.A0 + hd.A1 + 1) * 4);
  return ((w + 127) & ~(size_t)127) + ((a + 15) & ~(size_t)15) + 64;
}

bool mlp_shape_ok(int F, int H, int A0, int A1) {
  // A1 == 0: a single-head policy (gridworld, classic control); the second output part then
  // carries only the value column
  return F >= 1 && F <= 256 && H >= 16 && H <= 256 && (H % 32) == 0 && A0 >= 1 && A1 >= 0 &&
         A0 <= 32 && A1 + 1 <= 32;
}

}  // namespace

#ifdef WDB_PHASE_CLOCKS
WDB_API int wdb_debug_mlp_clocks(long long *host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_mlp_clk, sizeof(long long) * 96);
}
#endif

// wdb_set_option("mlp_max_ctas", n): CTAs (= SMs, one persistent CTA each) the NEXT forward
// launches may use; 0 = all.  Lets two policies' forwards run side by side on disjoint SMs.
int g_mlp_max_ctas = 0;
namespace wdb {
int g_pdl = 0;            // wdb_set_option("pdl", 0/1): programmatic dependent launches (A/B: slower)
}

WDB_API long long wdb_mlp_blob_bytes(int F, int H, int A0, int A1) {
  if (!mlp_shape_ok(F, H, A0, A1)) return -1;
  const MlpHeader hd = make_header(F, H, A0, A1);
  if (mlp_smem_bytes(hd) > 227 * 1024) return -1;
  return hd.total_bytes;
}

WDB_API int wdb_mlp_pack_weights(void *stream, void *blob, const float *w1, const float *b1,
                                 const float *w2, const float *b2, const float *wh0,
                                 const float *bh0, const float *wh1, const float *bh1,
                                 const float *wv, const float *bv, int F, int H, int A0,
                                 int A1) {
  if (!blob || !w1 || !b1 || !w2 || !b2 || !wh0 || !bh0 || !wv || !bv)
    return (int)cudaErrorInvalidValue;
  if (A1 > 0 && (!wh1 || !bh1)) return (int)cudaErrorInvalidValue;
  if (!mlp_shape_ok(F, H, A0, A1)) return (int)cudaErrorInvalidValue;
  const MlpHeader hd = make_header(F, H, A0, A1);
  cudaStream_t st = as_stream(stream);
  unsigned char *b = reinterpret_cast<unsigned char *>(blob);
  auto bf = [&](int off) { return reinterpret_cast<__nv_bfloat16 *>(b + off); };
  cudaError_t e = cudaMemsetAsync(b + hd.off_w3, 0, (size_t)hd.N3 * H * 2, st);
  if (e != cudaSuccess) return (int)e;
  const int T = 256;
  pack_matrix_kernel<<<(H * hd.K1 + T - 1) / T, T, 0, st>>>(w1, H, F, H, hd.K1, bf(hd.off_w1));
  pack_matrix_kernel<<<(H * H + T - 1) / T, T, 0, st>>>(w2, H, H, H, H, bf(hd.off_w2));
  // W3 = [head0; head1; value] stacked along N (zero padded to N3 rows by the memset)
  pack_rows_kernel<<<(A0 * H + T - 1) / T, T, 0, st>>>(wh0, A0, H, 0, H, bf(hd.off_w3));
  if (A1 > 0)
    pack_rows_kernel<<<(A1 * H + T - 1) / T, T, 0, st>>>(wh1, A1, H, A0, H, bf(hd.off_w3));
  pack_rows_kernel<<<(H + T - 1) / T, T, 0, st>>>(wv, 1, H, A0 + A1, H, bf(hd.off_w3));
  g_launch_count += 5;
  pack_misc_kernel<<<(max(H, hd.N3) + T - 1) / T, T, 0, st>>>(hd, b, b1, b2, bh0, bh1, bv);
  return finish_launch();
}

namespace {

template <bool TILES>
int launch_forward(void *stream, const void *blob, int F, int H, int A0, int A1,
                   const void *input, long long rows, float *probs0, float *probs1,
                   float *values) {
  if (!blob || !input || !probs0 || (A1 > 0 && !probs1) || rows <= 0)
    return (int)cudaErrorInvalidValue;
  if (!mlp_shape_ok(F, H, A0, A1)) return (int)cudaErrorInvalidValue;
  if (TILES && (reinterpret_cast<uintptr_t>(input) & 15)) return (int)cudaErrorInvalidValue;
  const MlpHeader hd = make_header(F, H, A0, A1);
  const size_t smem = mlp_smem_bytes(hd);
  if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp_forward_kernel<TILES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  const long long n_tiles = (rows + kTileM - 1) / kTileM;
  const int sm_budget = (g_mlp_max_ctas > 0 && g_mlp_max_ctas < kNumSMs) ? g_mlp_max_ctas : kNumSMs;
  const int grid = (int)min((long long)sm_budget, n_tiles);
  mlp_forward_kernel<TILES><<<grid, kThreads, smem, as_stream(stream)>>>(
      reinterpret_cast<const unsigned char *>(blob), input, rows, probs0, probs1, values, 0);
  return finish_launch();
}

// Launch with the programmatic-stream-serialization attribute when the "pdl" option is on:
// the grid may start while the previous kernel of the stream drains; the kernels order
// themselves with griddepcontrol.wait (under stream capture this becomes a programmatic edge
// of the CUDA graph).
int launch_forward_pair(void *stream, const wdb_mlp_pair &p) {
  MlpPairArgs a = {};
  size_t smem = 0;
  long long tiles[2];
  for (int w = 0; w < 2; w++) {
    if (!p.blob[w] || !p.obs[w] || !p.probs0[w] || (p.A1[w] > 0 && !p.probs1[w]) || p.rows[w] <= 0)
      return (int)cudaErrorInvalidValue;
    if (!mlp_shape_ok(p.F[w], p.H[w], p.A0[w], p.A1[w])) return (int)cudaErrorInvalidValue;
    const MlpHeader hd = make_header(p.F[w], p.H[w], p.A0[w], p.A1[w]);
    const size_t need = mlp_smem_bytes(hd);
    if (need > smem) smem = need;
    a.blob[w] = reinterpret_cast<const unsigned char *>(p.blob[w]);
    a.obs[w] = p.obs[w];
    a.rows[w] = p.rows[w];
    a.probs0[w] = p.probs0[w]; a.probs1[w] = p.probs1[w]; a.values[w] = p.values[w];
    tiles[w] = (p.rows[w] + kTileM - 1) / kTileM;
  }
  if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp_forward_pair_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  int cb = p.ctas_b;
  if (cb <= 0) {
    // cost in tile-times: the slower of the two shares (the smaller share's tiles run ~1.3x
    // slower: fewer CTAs keep its weights hot in L2)
    double best = 1e30;
    for (int k = 1; k < kNumSMs; k++) {
      const double ca = (double)((tiles[0] + kNumSMs - k - 1) / (kNumSMs - k));
      const double cbt = (double)((tiles[1] + k - 1) / k);
      const double cost = tiles[0] >= tiles[1] ? fmax(ca, 1.3 * cbt) : fmax(1.3 * ca, cbt);
      if (cost < best) { best = cost; cb = k; }
    }
  }
  if (cb >= kNumSMs) cb = kNumSMs - 1;
  long long ca = kNumSMs - cb;
  if (ca > tiles[0]) ca = tiles[0];
  if (cb > tiles[1]) cb = (int)tiles[1];
  a.split = (unsigned)ca;
  a.flags = p.flags;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ca + cb), 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, mlp_forward_pair_kernel, a);
  if (e != cudaSuccess) return (int)e;
  return finish_launch();
}

// fp32 obs [rows, F] -> bf16 A-operand tiles (128 rows x K1, canonical K-major layout, K
// padding and the rows of the last tile beyond `rows` = 0): one thread per 16-byte chunk
__global__ void pack_obs_kernel(const float *__restrict__ obs, long long rows, int F, int K1,
                                unsigned char *__restrict__ tiles) {
  const int nchunk = K1 / 8;
  const long long n_tiles = (rows + kTileM - 1) / kTileM;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tiles * kTileM * nchunk) return;
  // consecutive threads: the 8 rows of a core matrix, then the next chunk of the same rows
  const long long grp = i / (8 * nchunk);                 // 8-row group (global)
  const int in = (int)(i - grp * 8 * nchunk);
  const int c = in >> 3, rr = in & 7;
  const long long row = grp * 8 + rr;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; k++)
    v[k] = (row < rows && 8 * c + k < F) ? obs[row * F + 8 * c + k] : 0.0f;
  uint4 o;
  o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
  const long long t = row >> 7;
  const int r = (int)(row & 127);
  *reinterpret_cast<uint4 *>(tiles + t * ((long long)kTileM * K1 * 2) + (r >> 3) * (K1 * 16) +
                             c * 128 + (r & 7) * 16) = o;
}

}  // namespace

WDB_API long long wdb_mlp_obs_tiles_bytes(int F, long long rows) {
  if (F < 1 || F > 256 || rows <= 0) return -1;
  const long long n_tiles = (rows + kTileM - 1) / kTileM;
  return n_tiles * kTileM * round_up(F, 16) * 2;
}

WDB_API int wdb_mlp_pack_obs(void *stream, const float *obs, long long rows, int F,
                             void *tiles) {
  if (!obs || !tiles || rows <= 0 || F < 1 || F > 256) return (int)cudaErrorInvalidValue;
  const int K1 = round_up(F, 16);
  const long long n = ((rows + kTileM - 1) / kTileM) * kTileM * (K1 / 8);
  const int T = 256;
  pack_obs_kernel<<<(unsigned)((n + T - 1) / T), T, 0, as_stream(stream)>>>(
      obs, rows, F, K1, reinterpret_cast<unsigned char *>(tiles));
  return finish_launch();
}

WDB_API int wdb_mlp_policy_forward(void *stream, const void *blob, int F, int H, int A0, int A1,
                                   const float *obs, long long rows, float *probs0,
                                   float *probs1, float *values) {
  return launch_forward<false>(stream, blob, F, H, A0, A1, obs, rows, probs0, probs1, values);
}

WDB_API int wdb_mlp_policy_forward_pair(void *stream, const wdb_mlp_pair *pair) {
  if (!pair) return (int)cudaErrorInvalidValue;
  return launch_forward_pair(stream, *pair);
}

WDB_API int wdb_mlp_policy_forward_tiles(void *stream, const void *blob, int F, int H, int A0,
                                         int A1, const void *obs_tiles, long long rows,
                                         float *probs0, float *probs1, float *values) {
  return launch_forward<true>(stream, blob, F, H, A0, A1, obs_tiles, rows, probs0, probs1,
                              values);
}
