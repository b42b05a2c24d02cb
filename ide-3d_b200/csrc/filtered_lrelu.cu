// filtered_lrelu for sm_100a.
//   ide3d_filtered_lrelu_act : in-place gain * lrelu * clamp with 2-bit sign write / read -- the activation
//                              stage of the generic composition (filtered_lrelu.cu:1105-1211,
//                              filtered_lrelu.cpp:213-290).
//   ide3d_filtered_lrelu     : fused bias -> up-FIR -> act -> down-FIR (filtered_lrelu.cu:139-1099).
//                              Returns IDE3D_UNSUPPORTED for configurations without a fused kernel, which the
//                              caller resolves exactly like the reference's return code -1
//                              (filtered_lrelu.cpp:52-56, filtered_lrelu.py:223-229): upfirdn2d + act + upfirdn2d.
//
// Sign tensor (filtered_lrelu.cpp:82-96): uint8 [N, C, s_h, s_w/4], 2 bits per element, 4 elements per byte
// in x order; code 1 = value was negative (backward multiplies by slope), code 2 = value was clamped
// (backward gradient is 0; clamp wins over sign).  Filters are kernel arguments / shared memory here, never
// global __constant__ state, so concurrent streams are safe (the reference is not: filtered_lrelu.cu:77-78).
#include "common.cuh"

namespace ide3d {

struct ActArgs {
    void* x;
    unsigned char* s;
    int xw, xh, xc, xn;
    long long sxw, sxh, sxc, sxn;
    int sw, sh, sox, soy;
    float gain, slope, clamp;
};

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return (float)(*p); }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v) { *p = (T)v; }
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half(v); }

// MODE 0: plain forward, 1: write signs, 2: read signs.  One thread per element in x; a 16-lane group
// owns one 32-bit word of the sign tensor (16 elements * 2 bits).
template <typename T, int MODE>
__global__ void __launch_bounds__(256) lrelu_act_kernel(const ActArgs p) {
    const int lane16 = threadIdx.x & 15;
    const int width = (MODE == 1) ? p.sw : p.xw;              // launch covers the sign row when writing
    const int height = (MODE == 1) ? p.sh : p.xh;
    const long long planes = (long long)p.xc * p.xn;
    const long long words_x = (width + 15) >> 4;              // 16-element groups per row
    const long long total = planes * height * words_x * 16;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long grp = i >> 4;
        const int gx = (int)(grp % words_x);
        long long r = grp / words_x;
        const int y = (int)(r % height);
        const long long q = r / height;                       // n*C + c
        const int x = gx * 16 + lane16;
        const int n = (int)(q / p.xc), c = (int)(q - (long long)n * p.xc);
        T* pv = (T*)p.x + n * p.sxn + c * p.sxc + (long long)y * p.sxh + (long long)x * p.sxw;
        if (MODE == 1) {
            unsigned s = 0;
            if (x < p.xw && y < p.xh) {
                float v = ldf<T>(pv) * p.gain;
                if (v < 0.f) { v *= p.slope; s = 1; }
                if (fabsf(v) > p.clamp) { v = (v < 0.f) ? -p.clamp : p.clamp; s = 2; }
                stf<T>(pv, v);
            }
            s <<= (lane16 << 1);
            const unsigned m = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;
            s |= __shfl_xor_sync(m, s, 1);
            s |= __shfl_xor_sync(m, s, 2);
            s |= __shfl_xor_sync(m, s, 4);
            s |= __shfl_xor_sync(m, s, 8);
            if (lane16 == 0 && x < p.sw) {
                const long long is = x + (long long)p.sw * (y + (long long)p.sh * q);
                reinterpret_cast<unsigned*>(p.s)[is >> 4] = s;
            }
        } else if (x < p.xw) {
            float v = ldf<T>(pv) * p.gain;
            if (MODE == 2) {
                const unsigned sx = (unsigned)(x + p.sox), sy = (unsigned)(y + p.soy);
                if (sx < (unsigned)p.sw && sy < (unsigned)p.sh) {
                    const long long is = (sx >> 2) + (long long)(p.sw >> 2) * (sy + (long long)p.sh * q);
                    unsigned s = p.s[is];
                    s >>= (sx & 3) << 1;
                    if (s & 1) v *= p.slope;
                    if (s & 2) v = 0.f;
                }
            } else {
                if (v < 0.f) v *= p.slope;
                if (fabsf(v) > p.clamp) v = (v < 0.f) ? -p.clamp : p.clamp;
            }
            stf<T>(pv, v);
        }
    }
}

template <typename T>
static int launch_act(const ActArgs& a, int write_signs, int read_signs, cudaStream_t st) {
    const int width = write_signs ? a.sw : a.xw, height = write_signs ? a.sh : a.xh;
    const long long total = (long long)a.xc * a.xn * height * ((width + 15) >> 4) * 16;
    long long grid = ceil_div<long long>(total, 256);
    const long long cap = (long long)sm_count() * 16;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    if (write_signs) lrelu_act_kernel<T, 1><<<(unsigned)grid, 256, 0, st>>>(a);
    else if (read_signs) lrelu_act_kernel<T, 2><<<(unsigned)grid, 256, 0, st>>>(a);
    else lrelu_act_kernel<T, 0><<<(unsigned)grid, 256, 0, st>>>(a);
    IDE3D_CHECK_LAUNCH("lrelu_act_kernel");
    return IDE3D_OK;
}

}  // namespace ide3d

using namespace ide3d;

extern "C" int ide3d_filtered_lrelu_act(const ide3d_filtered_lrelu_act_params* q, ide3d_stream_t stream) {
    IDE3D_REQUIRE(q && q->x, "filtered_lrelu_act: null tensor");
    IDE3D_REQUIRE(q->x_w > 0 && q->x_h > 0 && q->x_c > 0 && q->x_n > 0, "x is empty");
    IDE3D_REQUIRE(!(q->write_signs && q->read_signs), "filtered_lrelu_act: cannot read and write signs at once");
    if (q->write_signs || q->read_signs) {
        IDE3D_REQUIRE(q->s != nullptr, "signs tensor missing");
        IDE3D_REQUIRE(q->s_w > 0 && q->s_h > 0 && (q->s_w & 3) == 0, "signs width must be a positive multiple of 4 elements");
        if (q->write_signs) {
            IDE3D_REQUIRE((q->s_w & 15) == 0, "written signs width must be a multiple of 16 elements");
            IDE3D_REQUIRE(q->s_w >= q->x_w && q->s_h >= q->x_h, "signs tensor smaller than x");
            IDE3D_REQUIRE((reinterpret_cast<uintptr_t>(q->s) & 3) == 0, "signs tensor must be 4-byte aligned");
        }
    }
    ActArgs a;
    a.x = q->x; a.s = q->s;
    a.xw = q->x_w; a.xh = q->x_h; a.xc = q->x_c; a.xn = q->x_n;
    a.sxw = q->x_stride_w; a.sxh = q->x_stride_h; a.sxc = q->x_stride_c; a.sxn = q->x_stride_n;
    a.sw = q->s_w; a.sh = q->s_h; a.sox = q->s_ofs_x; a.soy = q->s_ofs_y;
    a.gain = q->gain; a.slope = q->slope; a.clamp = q->clamp;
    cudaStream_t st = (cudaStream_t)stream;
    switch (q->dtype) {
        case IDE3D_F32: return launch_act<float>(a, q->write_signs, q->read_signs, st);
        case IDE3D_F16: return launch_act<__half>(a, q->write_signs, q->read_signs, st);
        case IDE3D_F64: return launch_act<double>(a, q->write_signs, q->read_signs, st);
    }
    IDE3D_FAIL(IDE3D_INVALID, "filtered_lrelu_act: unsupported dtype %d", q->dtype);
}

// ------------------------------------------------------------------------------------------------------------
// Fused kernel for separable filters: one block = one 32x32 output tile of one (n, c) plane; the 4x-sized intermediate
// lives only in shared memory.
//   s_in [IH][IW]   input tile + bias (zero outside the image)
//   h1   [IH][UW]   horizontal polyphase up-FIR
//   u    [UH][UW]   vertical up-FIR, * up^2 * gain, leaky ReLU, clamp (+ sign write / read)
//   v    [TOH][UW]  vertical down-FIR
//   y    [TOH][TOW] horizontal down-FIR -> global
// Filter taps (flipped here unless `flip`) are per-launch shared memory: no global filter state.
struct FusedArgs {
    const void* x; const void* b; const float* fu; const float* fd; void* y; unsigned char* s;
    int px0, py0, flip;
    float gain, slope, clamp;
    int xw, xh, xc, xn;
    long long sxw, sxh, sxc, sxn;
    int yw, yh;
    long long syw, syh, syc, syn;
    int sw, sh, sox, soy, mode;          // mode 0: plain, 1: write signs, 2: read signs
    int one_u, one_d;                    // 1x1 "full" filters carry their value once, not once per axis
};

template <int UP, int DOWN, int FU, int FD>
struct FusedGeom {
    static constexpr int TOW = 32, TOH = 32;
    static constexpr int UW = (TOW - 1) * DOWN + FD, UH = (TOH - 1) * DOWN + FD;     // intermediate tile
    static constexpr int TU = (FU + UP - 1) / UP;                                     // up-FIR taps per output phase
    static constexpr int IW = (UW + UP - 1) / UP + TU + 4, IH = (UH + UP - 1) / UP + TU + 4;   // input tile (+ slack for the 4-wide windows)
    static constexpr int IWP = IW | 1, UWP = UW | 1;                                  // odd pitches: rows and columns both conflict-free
    static constexpr int H1R = IH;                                                    // rows of the horizontally filtered buffer
    static constexpr int kFloats = IH * IWP + H1R * UWP + (UH + 4) * UWP + TOH * UWP + TOH * (TOW + 1) + 2 * 32;
};

// Each pass is register-tiled: a thread produces 4 outputs of one polyphase branch from a sliding window, with the
// filter taps of that branch in registers -> ~0.4 shared-memory loads per FMA instead of 2.
template <typename T, int UP, int DOWN, int FU, int FD>
__global__ void __launch_bounds__(256) filtered_lrelu_fused_kernel(const FusedArgs p, int tiles_x, int tiles_y) {
    using G = FusedGeom<UP, DOWN, FU, FD>;
    constexpr int LOG_UP = (UP == 1) ? 0 : (UP == 2 ? 1 : 2);
    constexpr int TU = G::TU;
    constexpr int WU = TU + 3;                    // window of the up passes (4 outputs of one phase)
    constexpr int WD = 3 * DOWN + FD;             // window of the down passes (4 outputs)
    extern __shared__ __align__(16) float fsm[];
    float* s_in = fsm;
    float* h1 = s_in + G::IH * G::IWP;
    float* u = h1 + G::H1R * G::UWP;
    float* v = u + (G::UH + 4) * G::UWP;
    float* yt = v + G::TOH * G::UWP;             // [TOH][TOW+1] output tile for coalesced stores
    float* fus = yt + G::TOH * (G::TOW + 1);     // [32] up filter, as applied (flipped unless p.flip)
    float* fds = fus + 32;                       // [32] down filter
    const int tid = threadIdx.x;
    if (tid < 32) {
        fus[tid] = (tid < FU) ? p.fu[p.flip ? tid : FU - 1 - tid] : 0.f;
        fds[tid] = (tid < FD) ? p.fd[p.flip ? tid : FD - 1 - tid] : 0.f;
    }
    const long long tiles_plane = (long long)tiles_x * tiles_y;
    const long long total = tiles_plane * p.xc * p.xn;
    const float act_gain = p.gain * (float)(UP * UP) * (p.one_u ? 1.f / p.fu[0] : 1.f);
    const float fd_scale_y = p.one_d ? 1.f / p.fd[0] : 1.f;
    __syncthreads();
    float fdr[FD];
#pragma unroll
    for (int k = 0; k < FD; ++k) fdr[k] = fds[k];

    for (long long blk = blockIdx.x; blk < total; blk += gridDim.x) {
        const long long plane = blk / tiles_plane;
        const int t = (int)(blk - plane * tiles_plane);
        const int n = (int)(plane / p.xc), c = (int)(plane - (long long)n * p.xc);
        const int ox_t = (t % tiles_x) * G::TOW, oy_t = (t / tiles_x) * G::TOH;
        const int ux_t = ox_t * DOWN, uy_t = oy_t * DOWN;                  // origin of the intermediate tile
        const int ix_t = (ux_t - p.px0) >> LOG_UP, iy_t = (uy_t - p.py0) >> LOG_UP;   // arithmetic shift = floor
        const T* xin = (const T*)p.x + n * p.sxn + c * p.sxc;
        const float bias = p.b ? ldf<T>((const T*)p.b + c) : 0.f;

        __syncthreads();
        for (int i = tid; i < G::IH * G::IW; i += 256) {
            const int ly = i / G::IW, lx = i - ly * G::IW;
            const int gx = ix_t + lx, gy = iy_t + ly;
            float val = 0.f;
            if ((unsigned)gx < (unsigned)p.xw && (unsigned)gy < (unsigned)p.xh) val = ldf<T>(xin + gy * p.sxh + gx * p.sxw) + bias;
            s_in[ly * G::IWP + lx] = val;
        }
        __syncthreads();

        // ---- horizontal up-FIR, one polyphase branch r at a time: outputs jx = j0 + UP*m use taps r, r+UP, ...
#pragma unroll
        for (int r = 0; r < UP; ++r) {
            float f[TU];
#pragma unroll
            for (int tt = 0; tt < TU; ++tt) f[tt] = (r + tt * UP < FU) ? fus[r + tt * UP] : 0.f;
            const int j0 = (p.px0 - r - ux_t) & (UP - 1);
            const int cnt = (G::UW - j0 + UP - 1) >> LOG_UP;                 // outputs of this branch
            const int mblocks = (cnt + 3) >> 2;
            for (int i = tid; i < G::IH * mblocks; i += 256) {
                const int mb = i / G::IH, ly = i - mb * G::IH;                // lanes walk down the rows
                const int m0 = mb * 4;
                const int first = ((ux_t + j0 + UP * m0 - p.px0 + r) >> LOG_UP) - ix_t;
                const float* row = s_in + ly * G::IWP + first;
                float w[WU];
#pragma unroll
                for (int q = 0; q < WU; ++q) w[q] = row[q];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float acc = 0.f;
#pragma unroll
                    for (int tt = 0; tt < TU; ++tt) acc = fmaf(f[tt], w[e + tt], acc);
                    if (m0 + e < cnt) h1[ly * G::UWP + j0 + UP * (m0 + e)] = acc;
                }
            }
        }
        __syncthreads();

        // ---- vertical up-FIR + activation (+ signs), branch by branch; lanes walk along x
#pragma unroll
        for (int r = 0; r < UP; ++r) {
            float f[TU];
#pragma unroll
            for (int tt = 0; tt < TU; ++tt) f[tt] = (r + tt * UP < FU) ? fus[r + tt * UP] : 0.f;
            const int j0 = (p.py0 - r - uy_t) & (UP - 1);
            const int cnt = (G::UH - j0 + UP - 1) >> LOG_UP;
            const int mblocks = (cnt + 3) >> 2;
            for (int i = tid; i < mblocks * G::UW; i += 256) {
                const int mb = i / G::UW, jx = i - mb * G::UW;
                const int m0 = mb * 4;
                const int first = ((uy_t + j0 + UP * m0 - p.py0 + r) >> LOG_UP) - iy_t;
                const float* col = h1 + first * G::UWP + jx;
                float w[WU];
#pragma unroll
                for (int q = 0; q < WU; ++q) w[q] = (first + q < G::H1R) ? col[q * G::UWP] : 0.f;
                const int ux = ux_t + jx;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (m0 + e >= cnt) break;
                    float acc = 0.f;
#pragma unroll
                    for (int tt = 0; tt < TU; ++tt) acc = fmaf(f[tt], w[e + tt], acc);
                    const int jy = j0 + UP * (m0 + e);
                    const int uy = uy_t + jy;
                    float val = acc * act_gain;
                    if (p.mode == 2) {
                        const unsigned sx = (unsigned)(ux + p.sox), sy = (unsigned)(uy + p.soy);
                        if (sx < (unsigned)p.sw && sy < (unsigned)p.sh) {
                            const long long is = (sx >> 2) + (long long)(p.sw >> 2) * (sy + (long long)p.sh * plane);
                            const unsigned sb = p.s[is] >> ((sx & 3) << 1);
                            if (sb & 1) val *= p.slope;
                            if (sb & 2) val = 0.f;
                        }
                    } else {
                        unsigned sg = 0;
                        if (val < 0.f) { val *= p.slope; sg = 1; }
                        if (fabsf(val) > p.clamp) { val = (val < 0.f) ? -p.clamp : p.clamp; sg = 2; }
                        // a block OWNS intermediate columns/rows [u_t, u_t + 32*DOWN) (plus the tail on the last tile):
                        // byte-aligned, so no two blocks ever touch the same sign byte; bits are ORed into the zero-filled tensor
                        const bool own = (jx < G::TOW * DOWN || ox_t + G::TOW >= p.yw) && (jy < G::TOH * DOWN || oy_t + G::TOH >= p.yh);
                        if (p.mode == 1 && sg && own && ux < p.sw && uy < p.sh) {
                            const long long is = (ux >> 2) + (long long)(p.sw >> 2) * (uy + (long long)p.sh * plane);
                            atomicOr(reinterpret_cast<unsigned*>(p.s + (is & ~3ll)), sg << (((is & 3) << 3) + ((ux & 3) << 1)));
                        }
                    }
                    u[jy * G::UWP + jx] = val;
                }
            }
        }
        __syncthreads();

        // ---- vertical down-FIR: 4 output rows per thread from one column window
        for (int i = tid; i < (G::TOH / 4) * G::UW; i += 256) {
            const int ob = i / G::UW, jx = i - ob * G::UW;
            const float* col = u + (ob * 4 * DOWN) * G::UWP + jx;
            float w[WD];
#pragma unroll
            for (int q = 0; q < WD; ++q) w[q] = col[q * G::UWP];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < FD; ++k) acc = fmaf(fdr[k], w[e * DOWN + k], acc);
                v[(ob * 4 + e) * G::UWP + jx] = acc * fd_scale_y;
            }
        }
        __syncthreads();
        // ---- horizontal down-FIR: 4 output columns per thread; lanes walk down the rows
        for (int i = tid; i < G::TOH * (G::TOW / 4); i += 256) {
            const int xb = i / G::TOH, oy = i - xb * G::TOH;
            const float* row = v + oy * G::UWP + xb * 4 * DOWN;
            float w[WD];
#pragma unroll
            for (int q = 0; q < WD; ++q) w[q] = row[q];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < FD; ++k) acc = fmaf(fdr[k], w[e * DOWN + k], acc);
                yt[oy * (G::TOW + 1) + xb * 4 + e] = acc;
            }
        }
        __syncthreads();
        T* yout = (T*)p.y + n * p.syn + c * p.syc;
        for (int i = tid; i < G::TOH * G::TOW; i += 256) {
            const int oy = i / G::TOW, ox = i - oy * G::TOW;
            if (oy_t + oy < p.yh && ox_t + ox < p.yw)
                stf<T>(yout + (long long)(oy_t + oy) * p.syh + (long long)(ox_t + ox) * p.syw, yt[oy * (G::TOW + 1) + ox]);
        }
    }
}

template <typename T, int UP, int DOWN, int FU, int FD>
static int launch_fused(const FusedArgs& a, cudaStream_t st) {
    using G = FusedGeom<UP, DOWN, FU, FD>;
    const size_t smem = (size_t)G::kFloats * sizeof(float);
    auto kern = filtered_lrelu_fused_kernel<T, UP, DOWN, FU, FD>;
    if (smem > 48 * 1024) IDE3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles_x = ceil_div(a.yw, G::TOW), tiles_y = ceil_div(a.yh, G::TOH);
    const long long total = (long long)tiles_x * tiles_y * a.xc * a.xn;
    int per_sm = 1;
    IDE3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > total) grid = total;
    kern<<<(unsigned)grid, 256, smem, st>>>(a, tiles_x, tiles_y);
    IDE3D_CHECK_LAUNCH("filtered_lrelu_fused_kernel");
    return IDE3D_OK;
}

template <typename T>
static int dispatch_fused(const FusedArgs& a, int up, int down, int fu, int fd, cudaStream_t st) {
#define IDE3D_FL_CASE(UP, DOWN, FU, FD) if (up == UP && down == DOWN && fu == FU && fd == FD) return launch_fused<T, UP, DOWN, FU, FD>(a, st);
    IDE3D_FL_CASE(2, 2, 12, 12)     // StyleGAN3 default (filter_size 6), filtered_lrelu.cu:1262
    IDE3D_FL_CASE(2, 2, 8, 8)
    IDE3D_FL_CASE(2, 2, 16, 16)
    IDE3D_FL_CASE(4, 2, 24, 12)
    IDE3D_FL_CASE(4, 2, 16, 8)
    IDE3D_FL_CASE(2, 4, 12, 24)
    IDE3D_FL_CASE(2, 1, 12, 1)
    IDE3D_FL_CASE(2, 1, 8, 1)
    IDE3D_FL_CASE(1, 2, 1, 12)
    IDE3D_FL_CASE(1, 2, 1, 8)
    IDE3D_FL_CASE(1, 1, 1, 1)
#undef IDE3D_FL_CASE
    IDE3D_FAIL(IDE3D_UNSUPPORTED, "filtered_lrelu: no fused kernel for up=%d down=%d fu=%d fd=%d", up, down, fu, fd);
}

extern "C" int ide3d_filtered_lrelu(const ide3d_filtered_lrelu_params* q, ide3d_stream_t stream) {
    IDE3D_REQUIRE(q && q->x && q->y && q->fu && q->fd, "filtered_lrelu: null tensor");
    IDE3D_REQUIRE(q->dtype == IDE3D_F32 || q->dtype == IDE3D_F16, "x and b must be float16 or float32");
    IDE3D_REQUIRE(q->up >= 1 && q->down >= 1, "up and down must be at least 1");
    IDE3D_REQUIRE(!(q->write_signs && q->read_signs), "filtered_lrelu: cannot read and write signs at once");
    // separable filters only (fu_h == 0), or 1x1 "full" filters when the factor is 1 (filtered_lrelu.py:178-181)
    const bool fu_ok = (q->fu_h == 0) || (q->fu_h == 1 && q->fu_w == 1);
    const bool fd_ok = (q->fd_h == 0) || (q->fd_h == 1 && q->fd_w == 1);
    if (!fu_ok || !fd_ok) IDE3D_FAIL(IDE3D_UNSUPPORTED, "filtered_lrelu: no fused kernel for 2-D filters %dx%d / %dx%d", q->fu_w, q->fu_h, q->fd_w, q->fd_h);
    if (q->write_signs || q->read_signs) {
        IDE3D_REQUIRE(q->s != nullptr && q->s_w > 0 && q->s_h > 0 && (q->s_w & 3) == 0, "signs tensor missing or malformed");
        if (q->write_signs) IDE3D_REQUIRE((q->s_w & 15) == 0 && (reinterpret_cast<uintptr_t>(q->s) & 3) == 0, "written signs need 16-element rows, 4-byte alignment");
    }
    FusedArgs a;
    a.x = q->x; a.b = q->b; a.fu = q->fu; a.fd = q->fd; a.y = q->y; a.s = q->s;
    a.px0 = q->pad_x0; a.py0 = q->pad_y0; a.flip = q->flip;
    a.gain = q->gain; a.slope = q->slope; a.clamp = q->clamp;
    a.xw = q->x_w; a.xh = q->x_h; a.xc = q->x_c; a.xn = q->x_n;
    a.sxw = q->x_stride_w; a.sxh = q->x_stride_h; a.sxc = q->x_stride_c; a.sxn = q->x_stride_n;
    a.yw = q->y_w; a.yh = q->y_h;
    a.syw = q->y_stride_w; a.syh = q->y_stride_h; a.syc = q->y_stride_c; a.syn = q->y_stride_n;
    a.sw = q->s_w; a.sh = q->s_h; a.sox = q->s_ofs_x; a.soy = q->s_ofs_y;
    a.mode = q->write_signs ? 1 : (q->read_signs ? 2 : 0);
    a.one_u = (q->fu_h == 1 && q->fu_w == 1); a.one_d = (q->fd_h == 1 && q->fd_w == 1);   // value applied once, not per axis
    cudaStream_t st = (cudaStream_t)stream;
    if (q->write_signs) IDE3D_CUDA(cudaMemsetAsync(q->s, 0, (size_t)q->x_n * q->x_c * q->s_h * (q->s_w >> 2), st));
    const int rc = (q->dtype == IDE3D_F32) ? dispatch_fused<float>(a, q->up, q->down, q->fu_w, q->fd_w, st)
                                            : dispatch_fused<__half>(a, q->up, q->down, q->fu_w, q->fd_w, st);
    return rc;
}
