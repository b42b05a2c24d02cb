// ROUND-1 KERNEL, kept for A/B measurements against the round-2 design in raymarch_tc.cu (IDE3D_TC_V1=1 selects it).
// Fused volume renderer, tensor-core decoder: the same chain as raymarch.cu (rays -> jitter -> cam2world -> 2x tri-plane
// gather -> decoder MLP -> alpha compositing) with the per-sample MLP executed as tcgen05.mma tiles.
//
//   tile   = 128 samples (M = 128): 4 x 32 consecutive samples of 4 neighbouring rays (2x2 pixels)
//   A      = gathered features [128 x 64] (texture 0..31 | shape 32..63), bf16 hi + lo, K-major, 128B-swizzled smem rows
//   layer 1: D1[128 x 64] = A . W1_blk^T      accumulators in TMEM (fp32)
//   epilog : tcgen05.ld D1 -> + b1 -> softplus -> bf16 hi + lo -> A2 [128 x 64] in smem
//   layer 2: D2[128 x n]  (+)= A2 . W2_blk^T  for the output-column range the block feeds
//   final  : tcgen05.ld D2 -> + b2 -> sigma -> warp product scan -> weighted accumulation in registers
// The decoder is processed in hidden blocks of 64 units (three-head decoder: one block per head; dense decoders: HID/64
// blocks).  Zero blocks of W1 (a head reads only the texture or only the shape half) and of W2 (a head feeds only its
// own output columns) are skipped by construction of the MMA program.
//
// Precision: every product is evaluated as hi*hi + hi*lo + lo*hi with bf16 operands and fp32 accumulation ("bf16x3"):
// operands carry 16 mantissa bits, the dropped lo*lo term is 2^-16 relative.  Measured against the fp32 oracle in
// tests/test_gpu_renderer.py (tolerance stated there).
//
// Warp-specialised, one persistent CTA per SM (512 threads):
//   warpgroups 0-1 (8 warps) consumers: wait for a full A stage, issue the UMMAs (one elected thread), run the softplus and
//                          compositing epilogues out of TMEM.  Warp w and w+4 share TMEM lanes 32*(w%4).. and split the
//                          columns: in the softplus epilogue each takes 32 of the 64 hidden units; in the final epilogue the
//                          lower warp takes sigma + the 19 semantic logits (and computes the compositing weight), the upper
//                          warp the 32 colour features (weight handed over through shared memory).  104 registers/thread.
//   warpgroups 2-3 (8 warps) producers: compute sample positions and gather features into a 3-stage ring of A tiles
//                          (152 registers/thread, 24 x LDG.128 in flight per lane); mbarrier full/empty hand-off, the "empty"
//                          arrive is the tcgen05.commit of the last MMA that reads the stage.
#include <stdlib.h>

#include "raymarch_common.cuh"
#include "tc_ptx.cuh"

namespace ide3d {
namespace v1 {

constexpr int kTcConsumerThreads = 256;                       // 8 warps: (TMEM lane quarter) x (column half)
constexpr int kTcProducerWarps = 8;
constexpr int kTcThreads = kTcConsumerThreads + kTcProducerWarps * 32;
constexpr int kTcStages = 3;
constexpr int kTcMaxBlocks = 3;
constexpr int kTileBytes = 128 * 128;                         // [128 rows x 64 bf16]
constexpr int kWTileBytes = 64 * 128;                         // [64 rows x 64 bf16]
constexpr int kTmemCols = 256;                                // D1 3 x 64 (one per hidden block) + D2 64

// ---- shared memory map (bytes) ----
constexpr int kSmW = 0;                                       // per block: W1 hi, W1 lo, W2 hi, W2 lo (8 KB each)
constexpr int kSmA2 = kSmW + kTcMaxBlocks * 4 * kWTileBytes;                // 98304: A2 hi, A2 lo
constexpr int kSmStage = kSmA2 + 2 * kTileBytes;                             // 131072: kTcStages x (A hi, A lo)
constexpr int kSmMisc = kSmStage + kTcStages * 2 * kTileBytes;              // 229376
constexpr int kSmMiscBytes = (kTcMaxBlocks * 64 + 64 + 128 + 8) * 4 + 128;  // b1[192], b2[64], w hand-off[128], wsum[4+4], mbarriers, tmem ptr
constexpr int kTcSmemBytes = kSmMisc + kSmMiscBytes + 1024;                 // + slack for the 1024-byte alignment

struct TcRun { int n0, n, accum; };
struct TcBlock {
    const float* w1; int w1_ld, k0, kcount;        // W1 rows of this hidden block; inputs land at A columns [k0, k0+kcount)
    const float* b1;
    const float* w2; int w2_ld, out0, outc;        // W2[out, hidden cols of this block]; rows feed outputs [out0, out0+outc)
    int nruns;
    TcRun runs[4];                                 // layer-2 MMAs: D2 columns [n0, n0+n), accumulate or overwrite
};
struct TcProgram {
    int nblocks;
    TcBlock blk[kTcMaxBlocks];
    unsigned written;                              // bit g: D2 columns [16g, 16g+16) are produced by some block
};

struct TcArgs {
    PlaneView tex, seg;
    ide3d_decoder dec;
    TcProgram prog;
    const float* cam2world;
    int n, res_w, res_h, steps;
    float cam_z, ray_start, ray_end, box_scale;
    int jitter_mode;
    const float* jitter_u;
    uint32_t seed_lo, seed_hi;
    int clamp_mode, last_back, white_back, fill_weight;
    float max_depth, noise_std;
    const float* noise;
    float *out_feat, *out_depth, *out_weights;
    int tiles_x, tiles_y;
    int debug;                                     // IDE3D_TC_DEBUG: 1 = producers skip the gather, 2 = consumer skips the decoder (timing experiments only)
};

// log2(1 + 2^t): the hidden softplus in base-2 units (see the weight set-up).  ex2 of the clamped argument cannot overflow;
// for t >= 24 the sum rounds to 2^t and lg2 returns t itself, and max(., t) keeps t beyond the clamp (the function is >= t).
__device__ __forceinline__ float softplus2(float t) {
    float e, l;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(t, 126.f)));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.f + e));
    return fmaxf(l, t);
}

// write element (row, k) of a [rows x 64] bf16 swizzle-128B tile
__device__ __forceinline__ void tile_store_bf16(unsigned char* tile, int row, int k, __nv_bfloat16 v) {
    *reinterpret_cast<__nv_bfloat16*>(tile + tc::sw128_offset(row, k >> 3) + (k & 7) * 2) = v;
}

// per-ray constants shared by producer and consumer code
struct RaySetup {
    int n, ray;
    bool ok;
    float dx, dy, dz, dnorm, spacing;
    float m00, m01, m02, m03, m10, m11, m12, m13, m20, m21, m22, m23;
    long long sample_base;
};
__device__ __forceinline__ RaySetup ray_setup(const TcArgs& a, int ptile, int quarter) {
    RaySetup r;
    const int tiles_per_frame = a.tiles_x * a.tiles_y;
    r.n = ptile / tiles_per_frame;
    const int t = ptile - r.n * tiles_per_frame;
    const int px = (t % a.tiles_x) * 2 + (quarter & 1);
    const int py = (t / a.tiles_x) * 2 + (quarter >> 1);
    r.ok = (px < a.res_w) && (py < a.res_h);
    r.ray = r.ok ? py * a.res_w + px : 0;
    const float x = linspace_at(-1.f, 1.f, a.res_w, px);
    const float y = linspace_at(1.f, -1.f, a.res_h, py);
    const float inv = 1.f / sqrtf(x * x + y * y + a.cam_z * a.cam_z);
    r.dx = x * inv; r.dy = y * inv; r.dz = a.cam_z * inv;
    r.dnorm = sqrtf(r.dx * r.dx + r.dy * r.dy + r.dz * r.dz);
    const float* M = a.cam2world + r.n * 16;
    r.m00 = M[0]; r.m01 = M[1]; r.m02 = M[2]; r.m03 = M[3];
    r.m10 = M[4]; r.m11 = M[5]; r.m12 = M[6]; r.m13 = M[7];
    r.m20 = M[8]; r.m21 = M[9]; r.m22 = M[10]; r.m23 = M[11];
    const int S = a.steps;
    r.spacing = (S > 1) ? linspace_at(a.ray_start, a.ray_end, S, 1) - linspace_at(a.ray_start, a.ray_end, S, 0) : 0.f;
    r.sample_base = ((long long)r.n * (a.res_w * a.res_h) + r.ray) * S;
    return r;
}
// jittered depth of sample s and of sample s+1 (z1, only meaningful when s+1 < S), and the jitter offset of s
__device__ __forceinline__ void sample_depths(const TcArgs& a, const RaySetup& r, int s, float& z0, float& off0, float& z1) {
    const int S = a.steps;
    z0 = linspace_at(a.ray_start, a.ray_end, S, s);
    z1 = (s + 1 < S) ? linspace_at(a.ray_start, a.ray_end, S, s + 1) : 0.f;
    off0 = 0.f;
    if (a.jitter_mode == IDE3D_JITTER_TENSOR) {
        off0 = (a.jitter_u[r.sample_base + s] - 0.5f) * r.spacing;
        if (s + 1 < S) z1 += (a.jitter_u[r.sample_base + s + 1] - 0.5f) * r.spacing;
    } else if (a.jitter_mode == IDE3D_JITTER_HASH) {
        const uint32_t gi = (uint32_t)(r.sample_base + s);
        off0 = (jitter_hash(gi, a.seed_lo, a.seed_hi) - 0.5f) * r.spacing;
        if (s + 1 < S) z1 += (jitter_hash(gi + 1u, a.seed_lo, a.seed_hi) - 0.5f) * r.spacing;
    } else if (a.jitter_mode == IDE3D_JITTER_ZVALS) {                  // depths given per sample (hierarchical second pass)
        z0 = a.jitter_u[r.sample_base + s];
        z1 = (s + 1 < S) ? a.jitter_u[r.sample_base + s + 1] : 0.f;
    }
}

__global__ void __launch_bounds__(kTcThreads, 1) raymarch_tc_v1_kernel(const TcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* b1s = reinterpret_cast<float*>(smem + kSmMisc);
    float* b2s = b1s + kTcMaxBlocks * 64;
    float* wbuf = b2s + 64;                                              // [128] compositing weight of each row (half 0 -> half 1)
    float* wsumbuf = wbuf + 128;                                         // [4] weights_sum of the 4 rays
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(wsumbuf + 8);      // [kTcStages] producers -> consumer
    uint64_t* bar_empty = bar_full + kTcStages;                        // [kTcStages] consumer (tcgen05.commit) -> producers
    uint64_t* bar_mma = bar_empty + kTcStages;                         // consumer-internal: MMA batch done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TcProgram& P = a.prog;

    // ---------------- one-time setup: weights -> bf16 hi/lo swizzled tiles, biases, barriers, TMEM
    for (int i = tid; i < P.nblocks * 64 * 64; i += kTcThreads) {
        const int b = i >> 12, j = (i >> 6) & 63, k = i & 63;
        const TcBlock& B = P.blk[b];
        unsigned char* base = smem + kSmW + b * 4 * kWTileBytes;
        // The hidden activation is evaluated in base 2: softplus(x) = ln2 * log2(1 + 2^(x*log2e)).  log2e is folded into W1 / b1
        // and ln2 into W2 here, once, so the per-sample epilogue is add-bias, ex2, +1, lg2 (softplus2 below).
        const float v1 = (k >= B.k0 && k < B.k0 + B.kcount) ? B.w1[j * B.w1_ld + (k - B.k0)] * 1.4426950408889634f : 0.f;   // W1[hidden j][input k]
        const float v2 = (j >= B.out0 && j < B.out0 + B.outc) ? B.w2[(j - B.out0) * B.w2_ld + k] * 0.6931471805599453f : 0.f; // W2[output j][hidden k]
        __nv_bfloat16 hi, lo;
        tc::split_bf16(v1, hi, lo);
        tile_store_bf16(base, j, k, hi);
        tile_store_bf16(base + kWTileBytes, j, k, lo);
        tc::split_bf16(v2, hi, lo);
        tile_store_bf16(base + 2 * kWTileBytes, j, k, hi);
        tile_store_bf16(base + 3 * kWTileBytes, j, k, lo);
    }
    for (int i = tid; i < kTcMaxBlocks * 64; i += kTcThreads) b1s[i] = (i < P.nblocks * 64) ? P.blk[i >> 6].b1[i & 63] * 1.4426950408889634f : 0.f;
    if (tid < 64) {
        float v = 0.f;
        for (int h = 0; h < a.dec.num_heads; ++h) {
            const ide3d_mlp_head& H = a.dec.heads[h];
            if (tid >= H.out_offset && tid < H.out_offset + H.out_count) v = H.b2[tid - H.out_offset];
        }
        b2s[tid] = v;
    }
    if (tid == 0) {
        for (int i = 0; i < kTcStages; ++i) { tc::mbar_init(&bar_full[i], 4); tc::mbar_init(&bar_empty[i], 1); }
        tc::mbar_init(bar_mma, 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, kTmemCols);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int S = a.steps;
    const int num_ptiles = a.tiles_x * a.tiles_y * a.n;
    const int chunks = (S + 31) >> 5;

    if (warp >= 8) {
        // =========================================================================== producers
        tc::setmaxnreg_inc<152>();
        const int pw = warp - 8, pg = pw >> 2, quarter = pw & 3;
        int q = 0;
        for (int pt = blockIdx.x; pt < num_ptiles; pt += gridDim.x) {
            const RaySetup r = ray_setup(a, pt, quarter);
            for (int ch = 0; ch < chunks; ++ch, ++q) {
                if ((q & 1) != pg) continue;
                const int stage = q % kTcStages, use = q / kTcStages;
                tc::mbar_wait(&bar_empty[stage], (use + 1) & 1);             // first use passes immediately
                const int s = ch * 32 + lane;
                const bool live = r.ok && (s < S);
                float cx = 4.f, cy = 4.f, cz = 4.f;
                if (live) {
                    float z0, off0, z1;
                    sample_depths(a, r, s, z0, off0, z1);
                    const float pcx = r.dx * z0 + off0 * r.dx, pcy = r.dy * z0 + off0 * r.dy, pcz = r.dz * z0 + off0 * r.dz;
                    cx = (r.m00 * pcx + r.m01 * pcy + r.m02 * pcz + r.m03) * a.box_scale;
                    cy = (r.m10 * pcx + r.m11 * pcy + r.m12 * pcz + r.m13) * a.box_scale;
                    cz = (r.m20 * pcx + r.m21 * pcy + r.m22 * pcz + r.m23) * a.box_scale;
                }
                unsigned char* a_hi = smem + kSmStage + stage * 2 * kTileBytes;
                unsigned char* a_lo = a_hi + kTileBytes;
                if (a.debug != 1) gather_chunk_axes(a.tex, a.seg, r.n, cx, cy, cz, lane,
                                      [&](int src, int qq, const float (&at)[4], const float (&as)[4]) {
                                          const int row = quarter * 32 + src;
                                          __nv_bfloat16 h[4], l[4];
#pragma unroll
                                          for (int j = 0; j < 4; ++j) tc::split_bf16(at[j], h[j], l[j]);
                                          uint32_t o = tc::sw128_offset(row, qq >> 1) + (qq & 1) * 8;
                                          *reinterpret_cast<uint2*>(a_hi + o) = make_uint2(tc::pack_bf16(h[0], h[1]), tc::pack_bf16(h[2], h[3]));
                                          *reinterpret_cast<uint2*>(a_lo + o) = make_uint2(tc::pack_bf16(l[0], l[1]), tc::pack_bf16(l[2], l[3]));
#pragma unroll
                                          for (int j = 0; j < 4; ++j) tc::split_bf16(as[j], h[j], l[j]);
                                          o = tc::sw128_offset(row, 4 + (qq >> 1)) + (qq & 1) * 8;
                                          *reinterpret_cast<uint2*>(a_hi + o) = make_uint2(tc::pack_bf16(h[0], h[1]), tc::pack_bf16(h[2], h[3]));
                                          *reinterpret_cast<uint2*>(a_lo + o) = make_uint2(tc::pack_bf16(l[0], l[1]), tc::pack_bf16(l[2], l[3]));
                                      });
                tc::fence_async_smem();                                      // my generic-proxy stores -> async proxy (UMMA)
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_full[stage]);
            }
        }
    } else {
        // =========================================================================== consumers
        tc::setmaxnreg_dec<104>();
        const int wg = warp & 3, half = warp >> 2;                 // TMEM lane quarter (= ray of the 2x2 tile), column half
        const uint32_t d1_col = tmem_base, d2_col = tmem_base + 64 * kTcMaxBlocks;
        const uint32_t lane_sel = (uint32_t)(wg * 32) << 16;
        unsigned char* a2_hi = smem + kSmA2;
        unsigned char* a2_lo = a2_hi + kTileBytes;
        const uint32_t a2_hi_u = tc::smem_u32(a2_hi), a2_lo_u = tc::smem_u32(a2_lo);
        const uint32_t w_u = tc::smem_u32(smem + kSmW);
        const uint32_t stage_u = tc::smem_u32(smem + kSmStage);
        const bool issuer = (warp == 0 && lane == 0);
        const int row = wg * 32 + lane;
        uint32_t parity = 0;

        auto issue_l1 = [&](int b, uint32_t a_hi_u, uint32_t a_lo_u) {
            const TcBlock& B = P.blk[b];
            const uint32_t idesc = tc::make_idesc_bf16(128, 64);
            const uint32_t w1hi = w_u + b * 4 * kWTileBytes, w1lo = w1hi + kWTileBytes;
            const int ks0 = B.k0 >> 4, ksn = B.kcount >> 4;
            for (int ks = 0; ks < ksn; ++ks) {
                const uint32_t off = (uint32_t)(ks0 + ks) * 32;                 // 16 bf16 = 32 bytes along K
                tc::umma_bf16(d1_col + b * 64, tc::make_sdesc_sw128(a_hi_u + off), tc::make_sdesc_sw128(w1hi + off), idesc, ks > 0);
                tc::umma_bf16(d1_col + b * 64, tc::make_sdesc_sw128(a_hi_u + off), tc::make_sdesc_sw128(w1lo + off), idesc, 1);
                tc::umma_bf16(d1_col + b * 64, tc::make_sdesc_sw128(a_lo_u + off), tc::make_sdesc_sw128(w1hi + off), idesc, 1);
            }
        };
        auto issue_l2 = [&](int b) {
            const TcBlock& B = P.blk[b];
            const uint32_t w2hi = w_u + b * 4 * kWTileBytes + 2 * kWTileBytes, w2lo = w2hi + kWTileBytes;
            for (int rr = 0; rr < B.nruns; ++rr) {
                const TcRun& R = B.runs[rr];
                const uint32_t idesc = tc::make_idesc_bf16(128, R.n);
                const uint32_t rowoff = (uint32_t)R.n0 * 128;                    // n0 is a multiple of 16 -> atom aligned
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t off = (uint32_t)ks * 32;
                    tc::umma_bf16(d2_col + R.n0, tc::make_sdesc_sw128(a2_hi_u + off), tc::make_sdesc_sw128(w2hi + rowoff + off), idesc, (R.accum || ks > 0));
                    tc::umma_bf16(d2_col + R.n0, tc::make_sdesc_sw128(a2_hi_u + off), tc::make_sdesc_sw128(w2lo + rowoff + off), idesc, 1);
                    tc::umma_bf16(d2_col + R.n0, tc::make_sdesc_sw128(a2_lo_u + off), tc::make_sdesc_sw128(w2hi + rowoff + off), idesc, 1);
                }
            }
        };

        int q = 0;
        for (int pt = blockIdx.x; pt < num_ptiles; pt += gridDim.x) {
            const RaySetup r = ray_setup(a, pt, wg);
            // half 0: semantic logits (19) ; half 1: colour features (32).  acc[32] covers both.
            float acc[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = 0.f;
            float acc_w = 0.f, acc_d = 0.f, carry = 1.f;

            for (int ch = 0; ch < chunks; ++ch, ++q) {
                const int stage = q % kTcStages, use = q / kTcStages;
                const uint32_t a_hi_u = stage_u + stage * 2 * kTileBytes, a_lo_u = a_hi_u + kTileBytes;
                tc::mbar_wait(&bar_full[stage], use & 1);
                tc::tc_fence_after();
                if (a.debug == 2) {                                      // timing experiment: hand the stage straight back
                    tc::bar_sync(1, kTcConsumerThreads);
                    if (issuer) tc::mbar_arrive(&bar_empty[stage]);
                    continue;
                }

                // ---- layer 1 of every hidden block in one batch (D1 has a 64-column slot per block); the A stage is free after it
                if (issuer) {
                    for (int b = 0; b < P.nblocks; ++b) issue_l1(b, a_hi_u, a_lo_u);
                    tc::umma_commit(bar_mma);
                    tc::umma_commit(&bar_empty[stage]);
                }
                tc::mbar_wait(bar_mma, parity);
                parity ^= 1;
                tc::tc_fence_after();
                // ---- per block: softplus epilogue in registers -> (wait until layer 2 of the previous block has read A2) -> A2
                //      -> layer 2 of this block is issued and runs while the next block's softplus is being computed
                for (int b = 0; b < P.nblocks; ++b) {
                    uint32_t ph[16], pl[16];
                    {
                        float v[32];
                        tc::tmem_ld32(d1_col + b * 64 + lane_sel + half * 32, v);
                        const float* bb = b1s + b * 64 + half * 32;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float h0 = softplus2(v[2 * j] + bb[2 * j]);
                            const float h1 = softplus2(v[2 * j + 1] + bb[2 * j + 1]);
                            const __nv_bfloat162 hh = __floats2bfloat162_rn(h0, h1);
                            const float2 back = __bfloat1622float2(hh);
                            const __nv_bfloat162 ll = __floats2bfloat162_rn(h0 - back.x, h1 - back.y);
                            ph[j] = *reinterpret_cast<const uint32_t*>(&hh);
                            pl[j] = *reinterpret_cast<const uint32_t*>(&ll);
                        }
                    }
                    if (b > 0) {                                             // A2 is still being read by layer 2 of block b-1
                        tc::mbar_wait(bar_mma, parity);
                        parity ^= 1;
                    }
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {                       // 8 hidden units = one 16-byte chunk
                        const uint32_t o = tc::sw128_offset(row, half * 4 + c8);
                        *reinterpret_cast<uint4*>(a2_hi + o) = make_uint4(ph[c8 * 4], ph[c8 * 4 + 1], ph[c8 * 4 + 2], ph[c8 * 4 + 3]);
                        *reinterpret_cast<uint4*>(a2_lo + o) = make_uint4(pl[c8 * 4], pl[c8 * 4 + 1], pl[c8 * 4 + 2], pl[c8 * 4 + 3]);
                    }
                    tc::fence_async_smem();
                    tc::tc_fence_before();
                    tc::bar_sync(1, kTcConsumerThreads);
                    tc::tc_fence_after();
                    if (issuer) {
                        issue_l2(b);
                        tc::umma_commit(bar_mma);
                    }
                }
                tc::mbar_wait(bar_mma, parity);
                parity ^= 1;
                tc::tc_fence_after();

                const int s = ch * 32 + lane;
                const bool live = r.ok && (s < S);
                float o32[32];
                if (half == 0) {
                    // ---- sigma + semantic logits (columns 32..63), compositing weight of this sample
                    float z0 = 0.f, off0 = 0.f, z1 = 0.f;
                    if (live) sample_depths(a, r, s, z0, off0, z1);
                    const float zj = z0 + off0;
                    tc::tmem_ld32(d2_col + lane_sel + 32, o32);
                    float sigma = (((P.written >> 3) & 1u) ? o32[19] : 0.f) + b2s[51];
                    if (a.noise != nullptr && live) sigma += a.noise_std * a.noise[r.sample_base + s];
                    const float delta = (s + 1 < S) ? (z1 - zj) * r.dnorm : 1e10f;
                    const float dens = (a.clamp_mode == IDE3D_CLAMP_SOFTPLUS) ? softplus_precise(sigma) : fmaxf(sigma, 0.f);
                    const float alpha = live ? 1.f - expf(-delta * dens) : 0.f;
                    const float keep = live ? (1.f - alpha + 1e-10f) : 1.f;
                    float total;
                    const float T = warp_exclusive_product(keep, lane, total) * carry;
                    carry *= total;
                    float w = alpha * T;
                    acc_w += w;
                    if (a.last_back && ch == chunks - 1) {
                        const float wsum_all = warp_sum(acc_w);
                        if (s == S - 1) w += 1.f - wsum_all;
                    }
                    wbuf[row] = w;                                           // hand the weight to the colour warp
                    tc::bar_sync(2 + wg, 64);
                    if (a.out_weights != nullptr && live) a.out_weights[r.sample_base + s] = w;
                    acc_d = fmaf(w, zj, acc_d);
#pragma unroll
                    for (int c = 0; c < 19; ++c) {
                        const float v = (((P.written >> (2 + (c >> 4))) & 1u) ? o32[c] : 0.f) + b2s[32 + c];
                        acc[c] = fmaf(w, v, acc[c]);
                    }
                } else {
                    // ---- colour features (columns 0..31)
                    tc::tmem_ld32(d2_col + lane_sel, o32);
                    tc::bar_sync(2 + wg, 64);
                    const float w = wbuf[row];
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float v = (((P.written >> (c >> 4)) & 1u) ? o32[c] : 0.f) + b2s[c];
                        acc[c] = fmaf(w, v, acc[c]);
                    }
                }
                tc::tc_fence_before();
                tc::bar_sync(1, kTcConsumerThreads);      // D2 drained and wbuf consumed before the next tile reuses them
            }

            // ---- per-ray reduction and store
            const long long ray_index = (long long)r.n * (a.res_w * a.res_h) + r.ray;
            float* of = a.out_feat + ray_index * (kOut - 1);
            if (half == 0) {
                const float wsum = warp_sum(acc_w);
                if (lane == 0) wsumbuf[wg] = wsum;
                tc::bar_sync(2 + wg, 64);
                float depth = warp_sum(acc_d);
                float mine = 0.f;
#pragma unroll
                for (int c = 0; c < 19; ++c) {
                    const float v = warp_sum(acc[c]);
                    if (c == lane) mine = v;
                }
                if (a.white_back) mine += 1.f - wsum;
                if (a.max_depth != 0.f) depth += (1.f - wsum) * a.max_depth;
                if (a.fill_weight) mine = wsum;
                if (r.ok) {
                    if (lane < 19) of[32 + lane] = mine;
                    if (lane == 0) a.out_depth[ray_index] = depth;
                }
            } else {
                tc::bar_sync(2 + wg, 64);
                const float wsum = wsumbuf[wg];
                float mine = 0.f;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float v = warp_sum(acc[c]);
                    if (c == lane) mine = v;
                }
                if (a.white_back) mine += 1.f - wsum;
                if (a.fill_weight) mine = wsum;
                if (r.ok) of[lane] = mine;
            }
            tc::bar_sync(1, kTcConsumerThreads);          // wsumbuf is reused by the next pixel tile
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, kTmemCols);
}

// Build the hidden-block program from the head list.  Returns false when the decoder does not fit
// (hidden not a multiple of 64, more than kTcMaxBlocks blocks, outputs beyond 64 columns).
static bool build_program(const ide3d_decoder& d, TcProgram& P) {
    P.nblocks = 0;
    P.written = 0;
    for (int h = 0; h < d.num_heads; ++h) {
        const ide3d_mlp_head& H = d.heads[h];
        if (H.hidden <= 0 || H.hidden % 64 != 0) return false;
        if (H.out_offset < 0 || H.out_count <= 0 || H.out_offset + H.out_count > 64) return false;
        if (H.in_sel < 0 || H.in_sel > 2) return false;
        const int in = (H.in_sel == 2) ? 64 : 32;
        for (int c = 0; c < H.hidden / 64; ++c) {
            if (P.nblocks == kTcMaxBlocks) return false;
            TcBlock& B = P.blk[P.nblocks++];
            B.w1 = H.w1 + (size_t)c * 64 * in; B.w1_ld = in;
            B.k0 = (H.in_sel == 1) ? 32 : 0; B.kcount = in;
            B.b1 = H.b1 + c * 64;
            B.w2 = H.w2 + c * 64; B.w2_ld = H.hidden;
            B.out0 = H.out_offset; B.outc = H.out_count;
            // layer-2 column range in units of 16, split into runs of equal "already written" status
            const int g0 = H.out_offset / 16, g1 = (H.out_offset + H.out_count + 15) / 16;
            B.nruns = 0;
            int gi = g0;
            while (gi < g1) {
                const int st = (P.written >> gi) & 1;
                int ge = gi + 1;
                while (ge < g1 && (int)((P.written >> ge) & 1) == st) ++ge;
                B.runs[B.nruns++] = TcRun{gi * 16, (ge - gi) * 16, st};
                gi = ge;
            }
            for (int q = g0; q < g1; ++q) P.written |= 1u << q;
        }
    }
    return P.nblocks > 0;
}

// entry used by ide3d_raymarch_fwd (raymarch.cu); IDE3D_UNSUPPORTED when this decoder / layout has no TC kernel
int launch_raymarch_tc_v1(const ide3d_raymarch_params* p, bool channels_last, cudaStream_t st) {
    if (!channels_last) IDE3D_FAIL(IDE3D_UNSUPPORTED, "raymarch_tc: planes must be channels-last");
    if (p->tex.stride_h != p->seg.stride_h || p->tex.stride_w != p->seg.stride_w)
        IDE3D_FAIL(IDE3D_UNSUPPORTED, "raymarch_tc: tex and seg planes must share strides");
    TcArgs a;
    if (!build_program(p->dec, a.prog)) IDE3D_FAIL(IDE3D_UNSUPPORTED, "raymarch_tc: decoder shape not supported");
    a.tex = make_view(p->tex); a.seg = make_view(p->seg); a.dec = p->dec;
    a.cam2world = p->cam2world;
    a.n = p->n; a.res_w = p->res_w; a.res_h = p->res_h; a.steps = p->num_steps;
    a.cam_z = (float)(-1.0 / tan((2.0 * 3.14159265358979323846 * (double)p->fov_deg / 360.0) / 2.0));
    a.ray_start = p->ray_start; a.ray_end = p->ray_end; a.box_scale = p->box_scale;
    a.jitter_mode = p->jitter_mode; a.jitter_u = p->jitter_u;
    a.seed_lo = (uint32_t)(p->jitter_seed & 0xffffffffu); a.seed_hi = (uint32_t)(p->jitter_seed >> 32);
    a.clamp_mode = p->clamp_mode; a.last_back = p->last_back; a.white_back = p->white_back;
    a.fill_weight = p->fill_weight; a.max_depth = p->max_depth;
    a.noise_std = p->noise_std; a.noise = (p->noise_std != 0.f) ? p->noise : nullptr;
    a.out_feat = p->out_feat; a.out_depth = p->out_depth; a.out_weights = p->out_weights;
    a.tiles_x = ceil_div(p->res_w, 2); a.tiles_y = ceil_div(p->res_h, 2);
    const char* dbg = tuning_env("IDE3D_TC_DEBUG");
    a.debug = dbg ? atoi(dbg) : 0;
    IDE3D_CUDA(cudaFuncSetAttribute(raymarch_tc_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
    const int num_tiles = a.tiles_x * a.tiles_y * a.n;
    int grid = sm_count();
    if (grid > num_tiles) grid = num_tiles;
    raymarch_tc_v1_kernel<<<grid, kTcThreads, kTcSmemBytes, st>>>(a);
    IDE3D_CHECK_LAUNCH("raymarch_tc_v1_kernel");
    return IDE3D_OK;
}

}  // namespace v1
}  // namespace ide3d
