// Fused volume renderer, tensor-core decoder, round-2 second design ("v3"): decoders whose density comes from its OWN hidden block
// (the generator's three-head decoder: texture -> colour, shape -> semantic logits, shape -> sigma).  Same chain as raymarch_tc.cu
// (training/volumetric_rendering.py:34-136 + dnnlib/util.py:580-617 + the decoder), same tiles (4x4 rays x 8 depth samples = 128
// rows), same producers; what changes is WHERE the alpha compositing happens:
//
//   v2  composites in registers: per tile every consumer lane reads its 52 decoder outputs back from TMEM and keeps 51 running sums
//       -- 51 live registers across the whole march, which leaves the compiler no room to overlap the softplus chains (measured:
//       consumer warps issue 12 % of the time, the rest is dependency stalls; profiles/r02d_ncu_raymarch_v2.txt).
//   v3  composites in the TENSOR CORE.  The compositing weight w_s of a sample is a per-ROW scale of the layer-2 input:
//           sum_s w_s * (W2 h_s) = W2 (sum_s w_s h_s)     and the MMA accumulator can do the sum over tiles by itself
//       so the hidden activations of the colour / semantic blocks are multiplied by w_s before they go back to TMEM as the layer-2
//       A operand, layer 2 ACCUMULATES into D2 over all tiles of a unit (12 at 96 samples), and D2 is read once per unit: 8-lane
//       reduction over the depth slots, + b2 * sum(w), store.  w_s needs sigma_s first: the sigma block is processed first and its
//       64 -> 1 second layer runs on the CUDA cores from the fp32 softplus outputs the lane already holds (64 FFMA, no shuffles,
//       no bf16 rounding) -- the sigma block needs no layer-2 MMA and no TMEM write-back at all.
//       Per tile a consumer lane now does 192 softplus + 64 FFMA + 128 FMUL and the transmittance scan; no output read-back, no
//       running sums, nothing live across tiles but T, sum(w) and sum(w z).
//
// Warp roles (TEAMS = 2: 640 threads):  warps 0-3 / 4-7  consumer groups (each owns 256 TMEM columns and one A stage stream),
// warps 8-15  producers (two teams),  warps 16 / 17  MMA issuers of group 0 / 1 (one elected lane each; descriptors by addition),
// warps 18 / 19 idle (they complete the issuers' warpgroup for setmaxnreg).  The issuer runs ahead of the consumers:
//     L1 sigma(t+1)  as soon as the consumers have read the sigma slot of tile t and stage(t+1) is full
//     L2 b(t)        when the four warps have written the scaled A2 of block b
//     L1 colour(t+1) right behind L2(t) (tensor-pipe order keeps the A2 reads ahead of the overwrite) -> releases the stage
#include "raymarch_tc_shared.cuh"

namespace ide3d {
namespace tc3 {

constexpr int kIssuerWarps = 4;                     // one warpgroup: warps 0 / 1 of it issue for group 0 / 1

template <int TEAMS> struct Cfg {
    static constexpr int kThreads = 32 * (kConsumerWarps + 4 * TEAMS + kIssuerWarps);
    // launch: 640 x 96 (TEAMS = 2).  issuers release down to 40, consumers to 88, producers grow to 128: 5120 + 22528 + 32768 <= 61440
    //         768 x 80 (TEAMS = 3).  issuers 40, consumers 72, producers 96: 5120 + 18432 + 36864 <= 61440
    static constexpr int kBaseRegs = (TEAMS == 2) ? 96 : 80;
    static constexpr int kIssuerRegs = 40;
    static constexpr int kConsumerRegs = (TEAMS == 2) ? 88 : 72;
    static constexpr int kProducerRegs = (TEAMS == 2) ? 128 : 96;
};

struct Args3 {
    TcArgs a;
    int coop;               // producers: 1 = all eight warps fill one tile together, 0 = two independent teams
    int stages;             // A buffers in flight (2, or 3 with coop)
    int prefetch;           // coop producers prefetch the next tile's texel lines into L1
    int sb;                 // index of the sigma block in a.prog.blk
    int nc;                 // number of colour / semantic blocks (the others), cb[] their indices in processing order
    int cb[kTcMaxBlocks];
};

__device__ __forceinline__ float4 lds4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// v[i] <- log2(1 + 2^(v[i] + b[i])) for 16 columns, written stage by stage so that the 16 chains overlap
__device__ __forceinline__ void softplus16(float (&v)[16], uint32_t bias_addr) {
    float e[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b = lds4(bias_addr + q * 16);
        v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) e[i] = ex2f(fminf(v[i], 126.f));
#pragma unroll
    for (int i = 0; i < 16; ++i) e[i] = lg2f(1.f + e[i]);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(e[i], v[i]);
}

template <int TEAMS>
__global__ void __launch_bounds__(Cfg<TEAMS>::kThreads, 1) raymarch_tc3_kernel(const Args3 A) {
    constexpr int kTcThreads = Cfg<TEAMS>::kThreads;
    const TcArgs& a = A.a;
    const int kStages = A.stages;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const TcProgram& P = a.prog;
    unsigned char* w_hi = smem;
    unsigned char* w_lo = smem + P.wpart;
    unsigned char* stage_base = smem + 2 * P.wpart;
    unsigned char* misc = stage_base + kStages * kStageBytes;
    float* b1s = reinterpret_cast<float*>(misc);                         // [3 x 64] hidden biases (x log2e)
    float* b2s = b1s + kTcMaxBlocks * 64;                                 // [64] output biases
    float* wsig = b2s + 64;                                               // [64] sigma head, second layer (x ln2)
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(wsig + 64);        // [kStages] producer team -> issuer
    uint64_t* bar_empty = bar_full + kMaxTeams;                          // [kStages] layer 1 of the tile complete -> producers
    uint64_t* bar_d1s = bar_empty + kMaxTeams;                           // [kGroups] layer 1 of the sigma block complete
    uint64_t* bar_d1c = bar_d1s + kGroups;                               // [kGroups] layer 1 of the colour / semantic blocks complete
    uint64_t* bar_s2free = bar_d1c + kGroups;                            // [kGroups] consumers have read the sigma slot (4 arrivals)
    uint64_t* bar_a2 = bar_s2free + kGroups;                             // [kGroups][kTcMaxBlocks] scaled A2 of a block is in TMEM (4 arrivals)
    uint64_t* bar_d2 = bar_a2 + kGroups * kTcMaxBlocks;                  // [kGroups] layer 2 of the unit's last tile complete
    uint64_t* bar_d2free = bar_d2 + kGroups;                             // [kGroups] consumers have read D2 (4 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_d2free + kGroups);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---------------- one-time setup: weights -> bf16 hi/lo swizzled tiles (compacted), biases, barriers, TMEM
    for (int i = tid; i < (2 * P.wpart) / 16; i += kTcThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = tid; i < P.nblocks * 64 * 64; i += kTcThreads) {
        const int b = i >> 12, j = (i >> 6) & 63, k = i & 63;
        const TcBlock& B = P.blk[b];
        __nv_bfloat16 hi, lo;
        if (k < B.kcount) {                                                  // W1[hidden j][input k] -> column k0 + k   (x log2e)
            tc::split_bf16(B.w1[j * B.w1_ld + k] * 1.4426950408889634f, hi, lo);
            tile_store_bf16(w_hi + B.w1_off, j, B.k0 + k, hi);
            tile_store_bf16(w_lo + B.w1_off, j, B.k0 + k, lo);
        }
        if (b != A.sb) {
            const int oc = B.w2_row0 + j;                                    // W2[output oc][hidden k] -> row j of the row block   (x ln2)
            if (oc >= B.out0 && oc < B.out0 + B.outc) {
                tc::split_bf16(B.w2[(oc - B.out0) * B.w2_ld + k] * 0.6931471805599453f, hi, lo);
                tile_store_bf16(w_hi + B.w2_off, j, k, hi);
                tile_store_bf16(w_lo + B.w2_off, j, k, lo);
            }
        }
    }
    for (int i = tid; i < kTcMaxBlocks * 64; i += kTcThreads) b1s[i] = (i < P.nblocks * 64) ? P.blk[i >> 6].b1[i & 63] * 1.4426950408889634f : 0.f;
    if (tid < 64) {
        float v = 0.f;
        for (int h = 0; h < a.dec.num_heads; ++h) {
            const ide3d_mlp_head& H = a.dec.heads[h];
            if (tid >= H.out_offset && tid < H.out_offset + H.out_count) v = H.b2[tid - H.out_offset];
        }
        b2s[tid] = v;
        wsig[tid] = P.blk[A.sb].w2[tid] * 0.6931471805599453f;              // W2_sigma[0][hidden tid] of this 64-unit block
    }
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { tc::mbar_init(&bar_full[i], A.coop ? 8 : 4); tc::mbar_init(&bar_empty[i], 1); }
        for (int g = 0; g < kGroups; ++g) {
            tc::mbar_init(&bar_d1s[g], 1); tc::mbar_init(&bar_d1c[g], 1); tc::mbar_init(&bar_d2[g], 1);
            tc::mbar_init(&bar_s2free[g], 4); tc::mbar_init(&bar_d2free[g], 4);
        }
        for (int i = 0; i < kGroups * kTcMaxBlocks; ++i) tc::mbar_init(&bar_a2[i], 4);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, kGroups * kGroupCols);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int S = a.steps;
    const Schedule sch = make_schedule(a);
    const int unit_stride = kGroups * gridDim.x;
    constexpr int kFirstProducer = kConsumerWarps, kFirstIssuer = kConsumerWarps + 4 * TEAMS;

    if (warp >= kFirstIssuer) {
        // =========================================================================== MMA issuers
        tc::setmaxnreg_dec<Cfg<TEAMS>::kIssuerRegs>();
        const int g = warp - kFirstIssuer;
        if (g < kGroups && lane == 0) {
            const int my_tiles = g ? sch.tiles1 : sch.tiles0;
            const uint32_t d1_col = tmem_base + g * kGroupCols, d2_col = d1_col + kD2Col;
            const uint32_t stage_u = tc::smem_u32(stage_base);
            const uint64_t wh0 = tc::make_sdesc_sw128(tc::smem_u32(w_hi)), wl0 = tc::make_sdesc_sw128(tc::smem_u32(w_lo));
            const uint32_t idesc64 = tc::make_idesc_bf16(128, 64);
            // layer 1 of one hidden block of the tile in `stage`: D1 slot b = A . W1_b^T (three bf16 products per K step).
            // Descriptor start addresses are in 16-byte units: + (bytes >> 4) moves the tile origin.
            auto issue_l1 = [&](int b, int stage) {
                const TcBlock& B = P.blk[b];
                const uint64_t ah0 = tc::make_sdesc_sw128(stage_u + stage * kStageBytes), al0 = ah0 + (kTileBytes >> 4);
                const uint64_t wh = wh0 + (B.w1_off >> 4), wl = wl0 + (B.w1_off >> 4);
                const int ks0 = B.k0 >> 4, ksn = B.kcount >> 4;
                for (int ks = 0; ks < ksn; ++ks) {
                    const uint64_t off = (uint64_t)((ks0 + ks) * 2);                 // 32 bytes along K
                    tc::umma_bf16(d1_col + b * 64, ah0 + off, wh + off, idesc64, ks > 0);
                    tc::umma_bf16(d1_col + b * 64, ah0 + off, wl + off, idesc64, 1);
                    tc::umma_bf16(d1_col + b * 64, al0 + off, wh + off, idesc64, 1);
                }
            };
            // layer 2 of one colour / semantic block: D2 (+)= (w . A2) . W2_b^T, A2 read from TMEM (bf16 hi / lo packed in place)
            auto issue_l2 = [&](int b, bool first_of_unit) {
                const TcBlock& B = P.blk[b];
                for (int rr = 0; rr < B.nruns; ++rr) {
                    const TcRun& R = B.runs[rr];
                    const uint32_t idesc = tc::make_idesc_bf16(128, R.n);
                    const uint64_t wh = wh0 + ((B.w2_off + (R.n0 - B.w2_row0) * 128) >> 4), wl = wl0 + ((B.w2_off + (R.n0 - B.w2_row0) * 128) >> 4);
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t ahi = d1_col + b * 64 + ks * 16, alo = ahi + 8;
                        const uint32_t acc = (!first_of_unit || R.accum || ks > 0) ? 1u : 0u;
                        tc::umma_bf16_ts(d2_col + R.n0, ahi, wh + ks * 2, idesc, acc);
                        tc::umma_bf16_ts(d2_col + R.n0, ahi, wl + ks * 2, idesc, 1);
                        tc::umma_bf16_ts(d2_col + R.n0, alo, wh + ks * 2, idesc, 1);
                    }
                }
            };
            if (my_tiles > 0) {
                {   // prologue: both layer-1 parts of tile 0
                    const int seq = seq_of(sch, g, 0), stage = seq % kStages;
                    tc::mbar_wait(&bar_full[stage], (seq / kStages) & 1);
                    tc::tc_fence_after();
                    issue_l1(A.sb, stage);
                    tc::umma_commit(&bar_d1s[g]);
                    for (int i = 0; i < A.nc; ++i) issue_l1(A.cb[i], stage);
                    tc::umma_commit(&bar_d1c[g]);
                    tc::umma_commit(&bar_empty[stage]);
                }
                uint32_t par_tile = 0, par_unit = 0;            // parities of the per-tile (s2free, a2) and per-unit (d2free) barriers
                int step = 0, unit_i = 0;
                for (int t = 0; t < my_tiles; ++t) {
                    const bool has_next = (t + 1 < my_tiles);
                    int nstage = 0;
                    if (has_next) {
                        const int seq = seq_of(sch, g, t + 1);
                        nstage = seq % kStages;
                        tc::mbar_wait(&bar_s2free[g], par_tile);                     // sigma slot of tile t has been read
                        tc::mbar_wait(&bar_full[nstage], (seq / kStages) & 1);
                        tc::tc_fence_after();
                        issue_l1(A.sb, nstage);
                        tc::umma_commit(&bar_d1s[g]);
                    }
                    for (int i = 0; i < A.nc; ++i) {
                        const int b = A.cb[i];
                        tc::mbar_wait(&bar_a2[g * kTcMaxBlocks + b], par_tile);
                        if (i == 0 && step == 0 && unit_i > 0) { tc::mbar_wait(&bar_d2free[g], par_unit); par_unit ^= 1; }
                        tc::tc_fence_after();
                        issue_l2(b, step == 0);
                    }
                    if (step == a.tiles_per_unit - 1) tc::umma_commit(&bar_d2[g]);
                    if (has_next) {
                        for (int i = 0; i < A.nc; ++i) issue_l1(A.cb[i], nstage);
                        tc::umma_commit(&bar_d1c[g]);
                        tc::umma_commit(&bar_empty[nstage]);
                    }
                    par_tile ^= 1;
                    if (++step == a.tiles_per_unit) { step = 0; ++unit_i; }
                }
            }
        }
    } else if (warp >= kFirstProducer) {
        // =========================================================================== producers
        if constexpr (Cfg<TEAMS>::kProducerRegs > Cfg<TEAMS>::kBaseRegs) tc::setmaxnreg_inc<Cfg<TEAMS>::kProducerRegs>();
        else if constexpr (Cfg<TEAMS>::kProducerRegs < Cfg<TEAMS>::kBaseRegs) tc::setmaxnreg_dec<Cfg<TEAMS>::kProducerRegs>();
        if (A.coop) producer_loop_coop(a, sch, stage_base, bar_full, bar_empty, warp - kFirstProducer, lane, kStages, A.prefetch);
        else producer_loop<TEAMS>(a, sch, stage_base, bar_full, bar_empty, warp - kFirstProducer, lane);
    } else {
        // =========================================================================== consumers
        if constexpr (Cfg<TEAMS>::kConsumerRegs < Cfg<TEAMS>::kBaseRegs) tc::setmaxnreg_dec<Cfg<TEAMS>::kConsumerRegs>();
        const int g = warp >> 2, qw = warp & 3;
        const uint32_t d1_col = tmem_base + g * kGroupCols, d2_col = d1_col + kD2Col;
        const uint32_t lane_sel = (uint32_t)(qw * 32) << 16;
        const uint32_t b1_u = tc::smem_u32(b1s), wsig_u = tc::smem_u32(wsig);
        const int seg0 = lane & ~7, d = lane & 7;
        const int my_units = g ? sch.cnt1 : sch.cnt0;
        uint32_t par_tile = 0, par_unit = 0;
        const float sig_b = b2s[51];

        for (int ui = 0; ui < my_units; ++ui) {
            const int unit = blockIdx.x * kGroups + g + ui * unit_stride;
            const RaySetup r = ray_setup(a, unit, lane >> 3, qw);
            float acc_w = 0.f, acc_wm = 0.f, acc_d = 0.f, T_in = 1.f;

            for (int step = 0; step < a.tiles_per_unit; ++step) {
                // ---- sigma block: softplus of its 64 hidden units, second layer (64 -> 1) on the CUDA cores
                tc::mbar_wait(&bar_d1s[g], par_tile);
                tc::tc_fence_after();
                float sig = 0.f;
#pragma unroll
                for (int c16 = 0; c16 < 4; ++c16) {
                    float v[16];
                    tc::tmem_ld16(d1_col + A.sb * 64 + c16 * 16 + lane_sel, v);
                    if (c16 == 3) {                                                  // the slot has been read: layer 1 of the next tile may overwrite it
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&bar_s2free[g]);
                    }
                    softplus16(v, b1_u + (A.sb * 64 + c16 * 16) * 4);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 ws = lds4(wsig_u + (c16 * 16 + q * 4) * 4);
                        sig = fmaf(v[4 * q], ws.x, sig); sig = fmaf(v[4 * q + 1], ws.y, sig);
                        sig = fmaf(v[4 * q + 2], ws.z, sig); sig = fmaf(v[4 * q + 3], ws.w, sig);
                    }
                }
                // ---- compositing weight of this sample (volumetric_rendering.py:34-74)
                const int s = step * kTileDepth + d;
                const bool live = r.ok && (s < S);
                float z0 = 0.f, off0 = 0.f, z1 = 0.f;
                if (live) sample_depths(a, r, s, z0, off0, z1);
                const float zj = z0 + off0;
                float w;
                {
                    float sigma = sig + sig_b;
                    if (a.noise != nullptr && live) sigma += a.noise_std * a.noise[r.sample_base + s];
                    const float delta = (s + 1 < S) ? (z1 - zj) * r.dnorm : 1e10f;
                    const float dens = (a.clamp_mode == IDE3D_CLAMP_SOFTPLUS) ? softplus_precise(sigma) : fmaxf(sigma, 0.f);
                    const float alpha = live ? 1.f - expf(-delta * dens) : 0.f;
                    const float keep = live ? (1.f - alpha + 1e-10f) : 1.f;
                    float tr = T_in, mine = T_in;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float kj = __shfl_sync(kFull, keep, seg0 + j);
                        if (j == d) mine = tr;
                        tr *= kj;
                    }
                    T_in = tr;
                    w = alpha * mine;
                    acc_w += w;
                    if (a.last_back && step == a.tiles_per_unit - 1) {
                        float ws = acc_w;
                        ws += __shfl_xor_sync(kFull, ws, 1); ws += __shfl_xor_sync(kFull, ws, 2); ws += __shfl_xor_sync(kFull, ws, 4);
                        if (s == S - 1) w += 1.f - ws;
                    }
                    acc_wm += w;
                    if (a.out_weights != nullptr && live) a.out_weights[r.sample_base + s] = w;
                    acc_d = fmaf(w, zj, acc_d);
                }
                // ---- colour / semantic blocks: softplus, x w, bf16 hi / lo back into the same TMEM columns (layer-2 A operand)
                tc::mbar_wait(&bar_d1c[g], par_tile);
                tc::tc_fence_after();
                for (int i = 0; i < A.nc; ++i) {
                    const int b = A.cb[i];
#pragma unroll
                    for (int c16 = 0; c16 < 4; ++c16) {
                        const uint32_t col = d1_col + b * 64 + c16 * 16 + lane_sel;
                        float v[16];
                        tc::tmem_ld16(col, v);
                        softplus16(v, b1_u + (b * 64 + c16 * 16) * 4);
                        uint32_t ph[8], pl[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float h0 = v[2 * j] * w, h1 = v[2 * j + 1] * w;
                            const __nv_bfloat162 hh = __floats2bfloat162_rn(h0, h1);
                            const float2 back = __bfloat1622float2(hh);
                            const __nv_bfloat162 ll = __floats2bfloat162_rn(h0 - back.x, h1 - back.y);
                            ph[j] = *reinterpret_cast<const uint32_t*>(&hh);
                            pl[j] = *reinterpret_cast<const uint32_t*>(&ll);
                        }
                        tc::tmem_st8(col, ph);
                        tc::tmem_st8(col + 8, pl);
                    }
                    tc::tmem_wait_st();
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&bar_a2[g * kTcMaxBlocks + b]);
                }
                par_tile ^= 1;
            }

            // ---- end of the unit: D2 holds sum_s w_s * (W2 h_s) per (ray, depth slot); reduce the 8 depth slots, add b2 * sum(w), store
            //      (weights_sum is the sum BEFORE the last_back correction, volumetric_rendering.py:56-72)
            tc::mbar_wait(&bar_d2[g], par_unit);
            par_unit ^= 1;
            tc::tc_fence_after();
            float wsum = acc_w, wused = acc_wm, depth = acc_d;
#pragma unroll
            for (int m = 1; m < 8; m <<= 1) {
                wsum += __shfl_xor_sync(kFull, wsum, m); wused += __shfl_xor_sync(kFull, wused, m); depth += __shfl_xor_sync(kFull, depth, m);
            }
            if (a.max_depth != 0.f) depth += (1.f - wsum) * a.max_depth;
            const long long ray_index = (long long)r.n * (a.res_w * a.res_h) + r.ray;
            float* of = a.out_feat + ray_index * (kOut - 1);
#pragma unroll
            for (int c16 = 0; c16 < 4; ++c16) {
                float o[16];
                tc::tmem_ld16(d2_col + lane_sel + c16 * 16, o);
                if (c16 == 3) {                                                      // D2 has been read: the next unit may overwrite it
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&bar_d2free[g]);
                }
                const bool written = (P.written >> c16) & 1u;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int ch = c16 * 16 + c;
                    if (ch < kOut - 1) {
                        float v = written ? o[c] : 0.f;
                        v += __shfl_xor_sync(kFull, v, 1); v += __shfl_xor_sync(kFull, v, 2); v += __shfl_xor_sync(kFull, v, 4);
                        v = fmaf(b2s[ch], wused, v);
                        if (a.white_back) v += 1.f - wsum;
                        if (a.fill_weight) v = wsum;
                        if (r.ok && (ch & 7) == d) of[ch] = v;
                    }
                }
            }
            if (r.ok && d == 0) a.out_depth[ray_index] = depth;
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, kGroups * kGroupCols);
}

}  // namespace tc3

// Does the program have a density head of its own?  One block whose only output is channel 51 (sigma), 64 hidden units, and no other
// block writing channel 51.  Fills the block order of the v3 kernel.
static bool tc3_eligible(tc3::Args3& A) {
    TcProgram& P = A.a.prog;
    A.sb = -1; A.nc = 0;
    for (int b = 0; b < P.nblocks; ++b) {
        const TcBlock& B = P.blk[b];
        if (B.out0 == kOut - 1 && B.outc == 1 && B.w2_ld == 64) {
            if (A.sb >= 0) return false;
            A.sb = b;
        } else {
            if (B.out0 + B.outc > kOut - 1) return false;
            A.cb[A.nc++] = b;
        }
    }
    if (A.sb < 0 || A.nc < 1) return false;
    // layer-2 runs again, over the colour / semantic blocks only (the sigma block issues no layer 2 here): a 16-column group is
    // overwritten by its first writer in the unit's first tile and accumulated into by everything after
    unsigned written = 0;
    for (int i = 0; i < A.nc; ++i) {
        TcBlock& B = P.blk[A.cb[i]];
        const int g0 = B.out0 / 16, g1 = (B.out0 + B.outc + 15) / 16;
        B.nruns = 0;
        int gi = g0;
        while (gi < g1) {
            const int st = (written >> gi) & 1;
            int ge = gi + 1;
            while (ge < g1 && (int)((written >> ge) & 1) == st) ++ge;
            B.runs[B.nruns++] = TcRun{gi * 16, (ge - gi) * 16, st};
            gi = ge;
        }
        for (int q = g0; q < g1; ++q) written |= 1u << q;
    }
    P.written = written;
    return true;
}

int launch_raymarch_tc3(const TcArgs& a, int teams, cudaStream_t st, bool& handled) {
    tc3::Args3 A;
    A.a = a;
    // (the three-team instantiation passes the small parity cases but does not finish at the full-size configuration -- not root-caused;
    //  IDE3D_TC_TEAMS=3 therefore selects the raymarch_tc.cu kernel)
    handled = (teams == 2) && tc3_eligible(A);
    if (!handled) return IDE3D_OK;
    A.coop = env_int("IDE3D_TC_COOP", 1, 0, 1);
    A.stages = A.coop ? env_int("IDE3D_TC_STAGES", 3, 2, 3) : 2;
    A.prefetch = env_int("IDE3D_TC_PREFETCH", 0, 0, 1);      // measured on B200: 1.34 ms with the prefetch, 1.11 ms without (profiles/r02i_*)
    const int smem = 2 * a.prog.wpart + A.stages * kStageBytes + (kTcMaxBlocks * 64 + 64 + 64) * 4 + (2 * kMaxTeams + (6 + kTcMaxBlocks) * kGroups) * 8 + 16 + 1024;
    int grid = sm_count();
    if (grid * kGroups > a.num_units) grid = ceil_div(a.num_units, kGroups);
    IDE3D_CUDA(cudaFuncSetAttribute(tc3::raymarch_tc3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc3::raymarch_tc3_kernel<2><<<grid, tc3::Cfg<2>::kThreads, smem, st>>>(A);
    IDE3D_CHECK_LAUNCH("raymarch_tc3_kernel");
    return IDE3D_OK;
}

}  // namespace ide3d
