// Fused volume renderer, tensor-core decoder (round-2 design): rays -> jitter -> cam2world -> 2x tri-plane gather -> decoder MLP
// -> alpha compositing in ONE kernel, no per-sample intermediate in HBM (training/volumetric_rendering.py:34-136,
// dnnlib/util.py:580-617 and the generator's decoder in one pass).
//
//   unit   = 4x4 neighbouring rays, marched front to back; tile = 8 consecutive depth samples of the 16 rays = 128 rows (M = 128)
//            row r: warp quarter r/32 = pixel row of the unit, lane = (pixel column)*8 + depth index   -> the 8 samples of a ray
//            sit in 8 adjacent lanes (compositing = 8-lane segmented product), rays that share a plane row / column share texels
//   A      = gathered features [128 x 64] (texture 0..31 | shape 32..63), bf16 hi + lo, K-major, 128B-swizzled smem rows
//   layer 1: D1[128 x 64] per hidden block = A . W1_blk^T, fp32 accumulators in TMEM
//   epilog : tcgen05.ld D1 -> + b1 -> softplus -> bf16 hi / lo -> tcgen05.st back INTO THE SAME TMEM COLUMNS (16 fp32 columns
//            become 8 packed hi + 8 packed lo columns: one K = 16 step)
//   layer 2: D2[128 x n] (+)= A2 . W2_blk^T with the A operand read FROM TMEM (tcgen05.mma [d], [a_tmem], b_desc): no shared-memory
//            round trip, no proxy fence, no block-wide barrier between the two layers
//   final  : tcgen05.ld D2 -> + b2 -> sigma -> alpha -> 8-lane product scan (+ carried transmittance) -> weighted accumulation in
//            registers; one 8-lane reduction per ray at the end of the march
// Every product is bf16 hi*hi + hi*lo + lo*hi with fp32 accumulation ("bf16x3": 16 mantissa bits per operand).
//
// Warp-specialised persistent CTA, 640 threads, one CTA per SM:
//   warps 0-3 / 4-7    two consumer GROUPS; a group owns 256 TMEM columns (D1 3x64 + D2 64) and marches its own units, so the MMA /
//                      commit latency of one group's tile is covered by the softplus / compositing work of the other group;
//                      lane 0 of a group's first warp issues its MMAs
//   warps 8-19         producers, three teams of four warps; a team fills one 32 KB A stage (a tile) at a time: sample positions,
//                      per-axis bilinear footprints (once per sample, by the lane that owns it), 8 lanes per texel x LDG.128,
//                      blend, bf16 split, swizzled st.shared; mbarrier full / empty ring, "empty" = tcgen05.commit of layer 1
// Shared memory: compacted weight tiles (52 KB for the three-head decoder) + kStages x 32 KB, so that 64+ KB stay L1 for the gather.
#include "raymarch_tc_shared.cuh"

namespace ide3d {

template <int TEAMS>
__global__ void __launch_bounds__(TcCfg<TEAMS>::kThreads, 1) raymarch_tc_kernel(const TcArgs a) {
    constexpr int kStages = TEAMS, kTeams = TEAMS, kTcThreads = TcCfg<TEAMS>::kThreads;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const TcProgram& P = a.prog;
    unsigned char* w_hi = smem;
    unsigned char* w_lo = smem + P.wpart;
    unsigned char* stage_base = smem + 2 * P.wpart;
    unsigned char* misc = stage_base + kStages * kStageBytes;
    float* b1s = reinterpret_cast<float*>(misc);                         // [3 x 64] hidden biases (x log2e)
    float* b2s = b1s + kTcMaxBlocks * 64;                                 // [64] output biases
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(b2s + 64);         // [kStages] producer team -> consumers
    uint64_t* bar_empty = bar_full + kMaxTeams;                      // [kStages] layer-1 commit -> producers
    uint64_t* bar_d1 = bar_empty + kMaxTeams;                        // [kGroups] layer 1 done
    uint64_t* bar_a2 = bar_d1 + kGroups;                                // [kGroups][kTcMaxBlocks] A2 of a hidden block is in TMEM (4 warp arrivals)
    uint64_t* bar_d2 = bar_a2 + kGroups * kTcMaxBlocks;                 // [kGroups] layer 2 done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_d2 + kGroups);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---------------- one-time setup: weights -> bf16 hi/lo swizzled tiles (compacted), biases, barriers, TMEM
    for (int i = tid; i < (2 * P.wpart) / 16; i += kTcThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = tid; i < P.nblocks * 64 * 64; i += kTcThreads) {
        const int b = i >> 12, j = (i >> 6) & 63, k = i & 63;
        const TcBlock& B = P.blk[b];
        __nv_bfloat16 hi, lo;
        // hidden activation in base 2: softplus(x) = ln2 * log2(1 + 2^(x*log2e)); log2e goes into W1 / b1, ln2 into W2, once
        if (k < B.kcount) {                                                  // W1[hidden j][input k] -> column k0 + k
            tc::split_bf16(B.w1[j * B.w1_ld + k] * 1.4426950408889634f, hi, lo);
            tile_store_bf16(w_hi + B.w1_off, j, B.k0 + k, hi);
            tile_store_bf16(w_lo + B.w1_off, j, B.k0 + k, lo);
        }
        const int oc = B.w2_row0 + j;                                        // W2[output oc][hidden k] -> row j of the row block
        if (oc >= B.out0 && oc < B.out0 + B.outc) {
            tc::split_bf16(B.w2[(oc - B.out0) * B.w2_ld + k] * 0.6931471805599453f, hi, lo);
            tile_store_bf16(w_hi + B.w2_off, j, k, hi);
            tile_store_bf16(w_lo + B.w2_off, j, k, lo);
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
    }
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { tc::mbar_init(&bar_full[i], 4); tc::mbar_init(&bar_empty[i], 1); }
        for (int g = 0; g < kGroups; ++g) { tc::mbar_init(&bar_d1[g], 1); tc::mbar_init(&bar_d2[g], 1); }
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

    if (warp >= kConsumerWarps) {
        // =========================================================================== producers
        if constexpr (TcCfg<TEAMS>::kProducerRegs < TcCfg<TEAMS>::kBaseRegs) tc::setmaxnreg_dec<TcCfg<TEAMS>::kProducerRegs>();
        else if constexpr (TcCfg<TEAMS>::kProducerRegs > TcCfg<TEAMS>::kBaseRegs) tc::setmaxnreg_inc<TcCfg<TEAMS>::kProducerRegs>();
        producer_loop<TEAMS>(a, sch, stage_base, bar_full, bar_empty, warp - kConsumerWarps, lane);
    } else {
        // =========================================================================== consumers
        if constexpr (TcCfg<TEAMS>::kConsumerRegs < TcCfg<TEAMS>::kBaseRegs) tc::setmaxnreg_dec<TcCfg<TEAMS>::kConsumerRegs>();
        else if constexpr (TcCfg<TEAMS>::kConsumerRegs > TcCfg<TEAMS>::kBaseRegs) tc::setmaxnreg_inc<TcCfg<TEAMS>::kConsumerRegs>();
        const int g = warp >> 2, qw = warp & 3;
        const uint32_t d1_col = tmem_base + g * kGroupCols, d2_col = d1_col + kD2Col;
        const uint32_t lane_sel = (uint32_t)(qw * 32) << 16;
        const uint32_t w_hi_u = tc::smem_u32(w_hi), w_lo_u = tc::smem_u32(w_lo);
        const uint32_t stage_u = tc::smem_u32(stage_base);
        const bool issuer = (qw == 0 && lane == 0);
        const int seg0 = lane & ~7, d = lane & 7;
        const int my_tiles = g ? sch.tiles1 : sch.tiles0, my_units = g ? sch.cnt1 : sch.cnt0;
        uint32_t par_d1 = 0, par_d2 = 0, par_a2 = 0;

        // layer 1 of every hidden block of one tile (D1 has a 64-column slot per block); the A stage is free when these MMAs are done
        auto issue_l1 = [&](int seq) {
            const int stage = seq % kStages;
            const uint32_t a_hi_u = stage_u + stage * kStageBytes, a_lo_u = a_hi_u + kTileBytes;
            const uint32_t idesc = tc::make_idesc_bf16(128, 64);
            for (int b = 0; b < P.nblocks; ++b) {
                const TcBlock& B = P.blk[b];
                const int ks0 = B.k0 >> 4, ksn = B.kcount >> 4;
                for (int ks = 0; ks < ksn; ++ks) {
                    const uint32_t off = (uint32_t)(ks0 + ks) * 32;                 // 16 bf16 = 32 bytes along K, same column in A and W1
                    const uint64_t ah = tc::make_sdesc_sw128(a_hi_u + off), al = tc::make_sdesc_sw128(a_lo_u + off);
                    const uint64_t wh = tc::make_sdesc_sw128(w_hi_u + B.w1_off + off), wl = tc::make_sdesc_sw128(w_lo_u + B.w1_off + off);
                    tc::umma_bf16(d1_col + b * 64, ah, wh, idesc, ks > 0);
                    tc::umma_bf16(d1_col + b * 64, ah, wl, idesc, 1);
                    tc::umma_bf16(d1_col + b * 64, al, wh, idesc, 1);
                }
            }
            tc::umma_commit(&bar_d1[g]);
            tc::umma_commit(&bar_empty[stage]);
        };
        // layer 2 of one hidden block: A2 (bf16 hi / lo, packed) is read from TMEM columns [64b + 16ks, +8) / [64b + 16ks + 8, +8)
        auto issue_l2 = [&](int b) {
            const TcBlock& B = P.blk[b];
            for (int rr = 0; rr < B.nruns; ++rr) {
                const TcRun& R = B.runs[rr];
                const uint32_t idesc = tc::make_idesc_bf16(128, R.n);
                const uint32_t rowoff = (uint32_t)(R.n0 - B.w2_row0) * 128;          // multiples of 16 rows: 1024-byte atom aligned
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t off = B.w2_off + rowoff + (uint32_t)ks * 32;
                    const uint32_t ahi = d1_col + b * 64 + ks * 16, alo = ahi + 8;
                    const uint64_t wh = tc::make_sdesc_sw128(w_hi_u + off), wl = tc::make_sdesc_sw128(w_lo_u + off);
                    tc::umma_bf16_ts(d2_col + R.n0, ahi, wh, idesc, (R.accum || ks > 0));
                    tc::umma_bf16_ts(d2_col + R.n0, ahi, wl, idesc, 1);
                    tc::umma_bf16_ts(d2_col + R.n0, alo, wh, idesc, 1);
                }
            }
        };

        bool l1_issued = false;                              // layer 1 of the NEXT tile already in flight (issuer only)
        int t = 0;
        for (int ui = 0; ui < my_units; ++ui) {
            const int unit = blockIdx.x * kGroups + g + ui * unit_stride;
            const RaySetup r = ray_setup(a, unit, lane >> 3, qw);
            float acc[kOut - 1];
#pragma unroll
            for (int c = 0; c < kOut - 1; ++c) acc[c] = 0.f;
            float acc_w = 0.f, acc_d = 0.f, T_in = 1.f;

            for (int step = 0; step < a.tiles_per_unit; ++step, ++t) {
                if (issuer && !l1_issued) {
                    const int seq = seq_of(sch, g, t);
                    tc::mbar_wait(&bar_full[seq % kStages], (seq / kStages) & 1);
                    tc::tc_fence_after();
                    issue_l1(seq);
                }
                l1_issued = false;
                __syncwarp();
                tc::mbar_wait(&bar_d1[g], par_d1);
                par_d1 ^= 1;
                tc::tc_fence_after();

                // ---- hidden blocks: softplus epilogue in place, 16 columns at a time, then layer 2 of the block
                for (int b = 0; b < P.nblocks; ++b) {
                    const float* bb = b1s + b * 64;
#pragma unroll
                    for (int c16 = 0; c16 < 4; ++c16) {
                        const uint32_t col = d1_col + b * 64 + c16 * 16 + lane_sel;
                        float v[16];
                        tc::tmem_ld16(col, v);
                        uint32_t ph[8], pl[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float h0 = softplus2(v[2 * j] + bb[c16 * 16 + 2 * j]);
                            const float h1 = softplus2(v[2 * j + 1] + bb[c16 * 16 + 2 * j + 1]);
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
                    if (qw == 0) {
                        if (lane == 0) {
                            tc::mbar_wait(&bar_a2[g * kTcMaxBlocks + b], par_a2);
                            tc::tc_fence_after();
                            issue_l2(b);
                            if (b == P.nblocks - 1) {
                                tc::umma_commit(&bar_d2[g]);
                                // layer 1 of this group's next tile right behind (MMAs execute in issue order: it overwrites D1 only
                                // after layer 2 above has read A2 from it) -- if its stage is already full; else after the epilogue
                                if (t + 1 < my_tiles) {
                                    const int seq = seq_of(sch, g, t + 1);
                                    if (tc::mbar_try_wait(&bar_full[seq % kStages], (seq / kStages) & 1)) {
                                        tc::tc_fence_after();
                                        issue_l1(seq);
                                        l1_issued = true;
                                    }
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
                par_a2 ^= 1;                             // every a2 barrier completes once per tile

                tc::mbar_wait(&bar_d2[g], par_d2);
                par_d2 ^= 1;
                tc::tc_fence_after();

                // ---- sigma + semantic logits (D2 columns 32..55), compositing weight of this sample
                const int s = step * kTileDepth + d;
                const bool live = r.ok && (s < S);
                float z0 = 0.f, off0 = 0.f, z1 = 0.f;
                if (live) sample_depths(a, r, s, z0, off0, z1);
                const float zj = z0 + off0;
                float w;
                {
                    float o16[16], o8[8];
                    tc::tmem_ld16(d2_col + lane_sel + 32, o16);
                    tc::tmem_ld8(d2_col + lane_sel + 48, o8);
                    float sigma = (((P.written >> 3) & 1u) ? o8[3] : 0.f) + b2s[51];
                    if (a.noise != nullptr && live) sigma += a.noise_std * a.noise[r.sample_base + s];
                    const float delta = (s + 1 < S) ? (z1 - zj) * r.dnorm : 1e10f;
                    const float dens = (a.clamp_mode == IDE3D_CLAMP_SOFTPLUS) ? softplus_precise(sigma) : fmaxf(sigma, 0.f);
                    const float alpha = live ? 1.f - expf(-delta * dens) : 0.f;
                    const float keep = live ? (1.f - alpha + 1e-10f) : 1.f;
                    // exclusive product over the 8 samples of the ray in sample order, times the transmittance carried in
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
                    if (a.out_weights != nullptr && live) a.out_weights[r.sample_base + s] = w;
                    acc_d = fmaf(w, zj, acc_d);
#pragma unroll
                    for (int c = 0; c < 16; ++c) acc[32 + c] = fmaf(w, (((P.written >> 2) & 1u) ? o16[c] : 0.f) + b2s[32 + c], acc[32 + c]);
#pragma unroll
                    for (int c = 0; c < 3; ++c) acc[48 + c] = fmaf(w, (((P.written >> 3) & 1u) ? o8[c] : 0.f) + b2s[48 + c], acc[48 + c]);
                }
                // ---- colour features (D2 columns 0..31)
#pragma unroll
                for (int c16 = 0; c16 < 2; ++c16) {
                    float o16[16];
                    tc::tmem_ld16(d2_col + lane_sel + c16 * 16, o16);
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        acc[c16 * 16 + c] = fmaf(w, (((P.written >> c16) & 1u) ? o16[c] : 0.f) + b2s[c16 * 16 + c], acc[c16 * 16 + c]);
                }
                tc::tc_fence_before();                   // D2 reads ordered before the a2 arrival that lets layer 2 of the next tile overwrite it
            }

            // ---- per-ray reduction over the 8 depth lanes and store (weights_sum is the sum BEFORE the last_back correction,
            //      volumetric_rendering.py:56-72)
            float wsum = acc_w, depth = acc_d;
#pragma unroll
            for (int m = 1; m < 8; m <<= 1) { wsum += __shfl_xor_sync(kFull, wsum, m); depth += __shfl_xor_sync(kFull, depth, m); }
            if (a.max_depth != 0.f) depth += (1.f - wsum) * a.max_depth;
            const long long ray_index = (long long)r.n * (a.res_w * a.res_h) + r.ray;
            float* of = a.out_feat + ray_index * (kOut - 1);
#pragma unroll
            for (int c = 0; c < kOut - 1; ++c) {
                float v = acc[c];
                v += __shfl_xor_sync(kFull, v, 1); v += __shfl_xor_sync(kFull, v, 2); v += __shfl_xor_sync(kFull, v, 4);
                if (a.white_back) v += 1.f - wsum;
                if (a.fill_weight) v = wsum;
                if (r.ok && (c & 7) == d) of[c] = v;
            }
            if (r.ok && d == 0) a.out_depth[ray_index] = depth;
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, kGroups * kGroupCols);
}

int launch_raymarch_tc3(const TcArgs& a, int teams, cudaStream_t st, bool& handled);     // raymarch_tc3.cu

// entry used by ide3d_raymarch_fwd (raymarch.cu); IDE3D_UNSUPPORTED when this decoder / layout has no TC kernel
int launch_raymarch_tc(const ide3d_raymarch_params* p, bool channels_last, cudaStream_t st) {
    if (!channels_last) IDE3D_FAIL(IDE3D_UNSUPPORTED, "raymarch_tc: planes must be channels-last");
    if (p->tex.stride_h != p->seg.stride_h || p->tex.stride_w != p->seg.stride_w)
        IDE3D_FAIL(IDE3D_UNSUPPORTED, "raymarch_tc: tex and seg planes must share strides");
    if (((long long)p->tex.h * p->tex.stride_h + (long long)p->tex.w * p->tex.stride_w + 96) * 4 >= (1ll << 31))
        IDE3D_FAIL(IDE3D_UNSUPPORTED, "raymarch_tc: plane too large for 32-bit byte offsets");
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
    a.units_x = ceil_div(p->res_w, kUnitW); a.units_y = ceil_div(p->res_h, kUnitH);
    a.num_units = a.units_x * a.units_y * a.n;
    a.tiles_per_unit = ceil_div(p->num_steps, kTileDepth);
    // tuning knob: gather instruction shape
    a.ray_major = env_int("IDE3D_TC_RAY_MAJOR", 1, 0, 1);
    const int teams = env_int("IDE3D_TC_TEAMS", 2, 2, kMaxTeams);
    const int smem = 2 * a.prog.wpart + teams * kStageBytes + (kTcMaxBlocks * 64 + 64) * 4 + (2 * kMaxTeams + (2 + kTcMaxBlocks) * kGroups) * 8 + 16 + 1024;
    int grid = sm_count();
    if (grid * kGroups > a.num_units) grid = ceil_div(a.num_units, kGroups);
    // decoders with a density head of their own take the kernel that composites inside the tensor core (raymarch_tc3.cu)
    if (env_int("IDE3D_TC_V2", 0, 0, 1) == 0) {
        bool handled = false;
        const int rc3 = launch_raymarch_tc3(a, teams, st, handled);
        if (handled) return rc3;
    }
    if (teams == 2) {
        IDE3D_CUDA(cudaFuncSetAttribute(raymarch_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        raymarch_tc_kernel<2><<<grid, TcCfg<2>::kThreads, smem, st>>>(a);
    } else {
        IDE3D_CUDA(cudaFuncSetAttribute(raymarch_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        raymarch_tc_kernel<3><<<grid, TcCfg<3>::kThreads, smem, st>>>(a);
    }
    IDE3D_CHECK_LAUNCH("raymarch_tc_kernel");
    return IDE3D_OK;
}

}  // namespace ide3d
