/*
 * ide3d_b200 -- C-ABI of the B200-native IDE-3D hot path (libide3d_b200.so).
 *
 * Drop-in boundary.  Every entry point below replaces one pybind11/Python boundary of the
 * reference (citations are relative to the reference tree, MrTornado24/IDE-3D @ b2ee653):
 *
 *   ide3d_bias_act            torch_utils/ops/bias_act.cpp:32       (params: bias_act.h:12-31)
 *   ide3d_modconv_epilogue    torch_utils/ops/fma.py:15 + bias_act.cpp:32 fused (extension; inversion/networks.py:104-105,512)
 *   ide3d_upfirdn2d           torch_utils/ops/upfirdn2d.cpp:16      (params: upfirdn2d.h:14-40)
 *   ide3d_upfirdn2d_add       upfirdn2d + the skip-connection add (extension; inversion/networks.py:841-844)
 *   ide3d_upfirdn2d_epilogue  upfirdn2d + demodulation/noise/bias_act tail (extension; inversion/networks.py:104-105,512)
 *   ide3d_filtered_lrelu      torch_utils/ops/filtered_lrelu.cpp:16 (params: filtered_lrelu.h:14-51)
 *   ide3d_filtered_lrelu_act  torch_utils/ops/filtered_lrelu.cpp:213 (params: filtered_lrelu.h:53-68)
 *   ide3d_initial_rays        training/volumetric_rendering.py:77   get_initial_rays_trig
 *   ide3d_transform_points    training/volumetric_rendering.py:99,108  perturb_points + transform_sampled_points
 *   ide3d_sample_triplane     dnnlib/util.py:580                    sample_from_triplane
 *   ide3d_integrate           training/volumetric_rendering.py:34   fancy_integration
 *   ide3d_sample_pdf          training/volumetric_rendering.py:224  sample_pdf
 *   ide3d_mask2color          dnnlib/seg_tools.py:75                mask2color
 *   ide3d_sample_voxel        generator.synthesis.renderer.sample_voxel (call site extract_shapes.py:146)
 *   ide3d_sigma_grid          extract_shapes.py:99-150 (create_samples :74-96 + the sample_voxel loop :144-148)
 *   ide3d_raymarch_fwd        the per-frame chain the generator class runs: rays -> jitter -> world
 *                             transform -> 2x tri-plane gather -> decoder MLP -> compositing, fused.
 *   ide3d_planes_to_nhwc      layout helper for the two kernels above (no reference counterpart).
 *
 * Conventions
 *   - plain C: raw device pointers, sizes, strides (in ELEMENTS), a cudaStream_t passed as void*.
 *   - the caller owns every buffer; the library never allocates or frees device memory, keeps no
 *     global device state (filters/weights travel as kernel arguments or per-launch shared memory,
 *     unlike filtered_lrelu.cu:77-78), and is therefore thread- and stream-safe.
 *   - kernels run on the CURRENT device of the calling thread, on the given stream.
 *   - return value: IDE3D_OK, or a negative code; nothing is thrown across the ABI.
 *     IDE3D_UNSUPPORTED (-1) keeps the reference meaning "no specialised kernel, caller may fall
 *     back to the generic composition" (filtered_lrelu.cpp:52-56).
 *   - ide3d_last_error() returns a thread-local, human-readable message for the last failure.
 */
#ifndef IDE3D_B200_H_
#define IDE3D_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDE3D_ABI_VERSION 1

enum ide3d_status {
    IDE3D_OK = 0,
    IDE3D_UNSUPPORTED = -1,   /* no kernel for this configuration (caller may fall back) */
    IDE3D_INVALID = -2,       /* malformed arguments (the reference raises via TORCH_CHECK) */
    IDE3D_CUDA_ERROR = -3     /* launch / runtime failure; see ide3d_last_error() */
};

enum ide3d_dtype { IDE3D_F32 = 0, IDE3D_F16 = 1, IDE3D_F64 = 2 };

typedef void* ide3d_stream_t; /* cudaStream_t */

int ide3d_abi_version(void);
const char* ide3d_last_error(void);
/* number of kernel launches issued by this library since load (all threads); for bench bookkeeping */
uint64_t ide3d_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * bias_act: y = clamp(gain * act(x + b)) and its 1st/2nd-order gradient forms.
 * Mirrors bias_act(x,b,xref,yref,dy,grad,dim,act,alpha,gain,clamp) (bias_act.cpp:32); the
 * caller resolves `dim` into size_b / step_b = x.stride(dim) exactly like bias_act.cpp:70-73.
 * x, xref, yref, dy, y: [size_x] dense, same dtype/layout; b: [size_b] or NULL.
 * act: 1 linear 2 relu 3 lrelu 4 tanh 5 sigmoid 6 elu 7 selu 8 softplus 9 swish (bias_act.py:21-31)
 * clamp < 0 disables clamping. */
int ide3d_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy,
                   void* y, int dtype, int grad, int act, float alpha, float gain, float clamp,
                   int64_t size_x, int64_t size_b, int64_t step_b, ide3d_stream_t stream);

/* Extension (no reference plugin): the tail of an activation-scaled modulated convolution in one pass,
 *     y = bias_act(x * scale[n,c] + noise[(n),h,w], b[c], act, alpha, gain, clamp)
 * i.e. fma.fma (inversion/networks.py:104-105; torch_utils/ops/fma.py:15) followed by bias_act (:512).
 * x, y: [n, c, h*w] dense (channels_last = 0) or [n, h*w, c] dense (channels_last = 1), dtype as bias_act;
 * scale [n*c], noise [noise_batch * h*w] (noise_batch = 1 or n), b [c]: same dtype as x, each may be NULL.
 * Optional second output y2 = y * scale2[n,c] (scale2 [n*c]; y2 like y): the style modulation `x * styles` that opens
 * the NEXT modulated convolution (inversion/networks.py:100), written in the same pass; y may then be NULL.
 * Forward only.  IDE3D_UNSUPPORTED when the vector width does not divide h*w (NCHW) or c (channels_last). */
int ide3d_modconv_epilogue(const void* x, const void* scale, const void* noise, const void* b, void* y,
                           const void* scale2, void* y2, int dtype, int act, float alpha, float gain, float clamp,
                           int64_t n, int64_t c, int64_t hw, int64_t noise_batch, int channels_last,
                           ide3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * upfirdn2d: pad -> zero-upsample -> FIR -> decimate.  Field meaning = upfirdn2d_kernel_params
 * (upfirdn2d.h:14-40); sizes/strides in the reference order [W, H, C, N]; strides in elements.
 * f is float32 [f_h, f_w] with element strides f_stride_{h,w}.  flip != 0 = correlation. */
typedef struct ide3d_upfirdn2d_params {
    const void* x;
    const float* f;
    void* y;
    int dtype;
    int up_x, up_y, down_x, down_y, pad_x0, pad_y0;
    int flip;
    float gain;
    int in_w, in_h, in_c, in_n;
    int64_t in_stride_w, in_stride_h, in_stride_c, in_stride_n;
    int f_w, f_h;
    int64_t f_stride_w, f_stride_h;
    int out_w, out_h;
    int64_t out_stride_w, out_stride_h, out_stride_c, out_stride_n;
} ide3d_upfirdn2d_params;
int ide3d_upfirdn2d(const ide3d_upfirdn2d_params* p, ide3d_stream_t stream);

/* Extension (no reference plugin): the skip-connection step of a 'skip' synthesis block in one pass,
 *     y = upfirdn2d(x, f, ...) + add + bias[c]
 * i.e. upsample2d of the running image (inversion/networks.py:841) fused with the `img.add_(y)` that follows (:844) and
 * with the ToRGB bias (:707).  add: same dtype, logical shape of y, element strides add_stride_{n,h,w}, channel stride 1
 * (it may be a channel slice of a wider channels_last tensor); bias [c] or NULL.  Only the channels_last patch kernel
 * implements it (x, y channels_last, C % 4 == 0, 4x4 filter, up/down in {1,2}); otherwise IDE3D_UNSUPPORTED. */
int ide3d_upfirdn2d_add(const ide3d_upfirdn2d_params* p, const void* add, int64_t add_stride_n, int64_t add_stride_h,
                        int64_t add_stride_w, const void* bias, ide3d_stream_t stream);

/* Extension (no reference plugin): upfirdn2d with the modulated-convolution tail applied to the filter output before it is
 * stored -- the up=2 SynthesisLayer is conv_transpose2d -> FIR -> x*dcoefs + noise -> bias_act (inversion/networks.py:104-105,
 * :512; conv2d_resample.py:112-126) and this runs everything after the transposed convolution in one pass:
 *     v  = clamp(gain * act(fir * scale[n,c] + noise[(n),oy,ox] + b[c]))     act: 1 linear, 3 lrelu(alpha)
 *     y  = v                      (params->y; may be NULL when only y2 is wanted)
 *     y2 = v * scale2[n,c]        (optional: the next layer's style modulation, layout of y)
 * scale, b, scale2: dtype of x, [n*c] / [c] / [n*c]; noise: dtype of x, [noise_batch, out_h, out_w] dense; any may be NULL.
 * Same kernel restrictions as ide3d_upfirdn2d_add (channels_last, C % 4 == 0, 4x4 filter); otherwise IDE3D_UNSUPPORTED. */
typedef struct ide3d_fir_epilogue {
    const void *scale, *noise, *b, *scale2;
    void* y2;
    int act;
    float alpha, gain, clamp;      /* clamp < 0: off */
    int64_t noise_batch;           /* 1 or n */
} ide3d_fir_epilogue;
int ide3d_upfirdn2d_epilogue(const ide3d_upfirdn2d_params* p, const ide3d_fir_epilogue* e, ide3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * filtered_lrelu: bias -> up-FIR -> gain*lrelu*clamp (+ 2-bit sign tensor) -> down-FIR, fused.
 * Field meaning = filtered_lrelu_kernel_params (filtered_lrelu.h:14-51).  fu/fd are float32, 1-D
 * (separable, *_h == 0) or 2-D.  s is the uint8 sign tensor [N,C,s_h,s_w/4] (filtered_lrelu.cpp:82-96)
 * written when write_signs, read when read_signs, ignored when both 0.  Returns IDE3D_UNSUPPORTED
 * for configurations without a fused kernel (the caller then composes upfirdn2d +
 * ide3d_filtered_lrelu_act exactly like filtered_lrelu.py:223-229). */
typedef struct ide3d_filtered_lrelu_params {
    const void* x;
    const void* b;          /* [C] same dtype as x, or NULL */
    const float* fu;
    const float* fd;
    void* y;
    unsigned char* s;
    int dtype;
    int up, down;
    int fu_w, fu_h;         /* fu_h == 0 -> separable 1-D filter of fu_w taps */
    int fd_w, fd_h;
    int pad_x0, pad_y0;
    int flip;
    float gain, slope, clamp;
    int x_w, x_h, x_c, x_n;
    int64_t x_stride_w, x_stride_h, x_stride_c, x_stride_n;
    int y_w, y_h;
    int64_t y_stride_w, y_stride_h, y_stride_c, y_stride_n;
    int s_w, s_h;           /* sign tensor extent in ELEMENTS (s_w multiple of 4) */
    int s_ofs_x, s_ofs_y;
    int write_signs, read_signs;
} ide3d_filtered_lrelu_params;
int ide3d_filtered_lrelu(const ide3d_filtered_lrelu_params* p, ide3d_stream_t stream);

/* In-place activation + sign handling of the generic fallback (filtered_lrelu.cpp:213). */
typedef struct ide3d_filtered_lrelu_act_params {
    void* x;                /* in/out */
    unsigned char* s;
    int dtype;
    int x_w, x_h, x_c, x_n;
    int64_t x_stride_w, x_stride_h, x_stride_c, x_stride_n;
    int s_w, s_h;
    int s_ofs_x, s_ofs_y;
    float gain, slope, clamp;
    int write_signs, read_signs;
} ide3d_filtered_lrelu_act_params;
int ide3d_filtered_lrelu_act(const ide3d_filtered_lrelu_act_params* p, ide3d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Volume renderer.
 * A tri-plane tensor is [N, 3*32, H, W] float32: planes (xy, yz, xz) of 32 channels each
 * (dnnlib/util.py:586-596), addressed through element strides so that both NCHW and
 * channels_last tensors are accepted.  The fused kernels take their fast path when stride_c == 1
 * (one texel = one 128-byte line per plane); otherwise convert once with ide3d_planes_to_nhwc. */
typedef struct ide3d_triplane {
    const float* data;
    int n, h, w;            /* batch, plane height, plane width */
    int64_t stride_n, stride_c, stride_h, stride_w;
} ide3d_triplane;

/* Decoder MLP, as a list of independent heads.  Each head reads the 32 texture features
 * (in_sel 0), the 32 shape features (in_sel 1) or their concatenation (in_sel 2, 64 inputs):
 *     h = softplus(W1 @ f + b1)   W1 [hidden, in] row-major
 *     o[out_offset : out_offset+out_count] = W2 @ h + b2   W2 [out_count, hidden] row-major
 * Output channels not covered by any head are 0.  Channel 51 (the last of 52) is sigma
 * (extract_shapes.py:146-147).  Fused kernels exist for:
 *     1 head : in_sel 2, out 0..52,           hidden in {64, 128}
 *     3 heads: (tex -> 0..32), (seg -> 32..51), (seg -> 51..52), hidden 64 each
 * anything else returns IDE3D_UNSUPPORTED. */
typedef struct ide3d_mlp_head {
    int in_sel, hidden, out_offset, out_count;
    const float *w1, *b1, *w2, *b2;
} ide3d_mlp_head;
typedef struct ide3d_decoder {
    int num_heads;
    ide3d_mlp_head heads[4];
} ide3d_decoder;

/* ZVALS: jitter_u holds the depth of every sample itself, [N, R, S] ascending along S (the merged coarse + importance samples of
 * a hierarchical second pass, volumetric_rendering.py:224-265): point = ray direction * z, no linspace, no jitter. */
enum ide3d_jitter { IDE3D_JITTER_NONE = 0, IDE3D_JITTER_TENSOR = 1, IDE3D_JITTER_HASH = 2, IDE3D_JITTER_ZVALS = 3 };
enum ide3d_clamp { IDE3D_CLAMP_SOFTPLUS = 0, IDE3D_CLAMP_RELU = 1 };
/* AUTO: tensor cores when the decoder / layout allows, else CUDA cores.  FP32: CUDA-core FFMA, plain fp32.
 * TC: tcgen05 tensor cores, every product as bf16 hi*hi + hi*lo + lo*hi with fp32 accumulation (16-bit operand
 * mantissas); IDE3D_UNSUPPORTED if the decoder shape or plane layout has no tensor-core kernel. */
enum ide3d_precision { IDE3D_PRECISION_AUTO = 0, IDE3D_PRECISION_FP32 = 1, IDE3D_PRECISION_TC = 2 };

typedef struct ide3d_raymarch_params {
    ide3d_triplane tex, seg;
    ide3d_decoder dec;
    const float* cam2world;     /* [N,16] row-major 4x4 (c[:, :16], gen_images.py:105-107) */
    int n;                      /* frames */
    int res_w, res_h;           /* neural render resolution (64 x 64) */
    int num_steps;              /* depth samples per ray */
    float fov_deg, ray_start, ray_end;
    float box_scale;            /* world -> plane grid units (2 / box_warp) */
    int jitter_mode;            /* ide3d_jitter */
    const float* jitter_u;      /* [N, R, S] uniforms in [0,1) when jitter_mode == TENSOR; sample depths when ZVALS */
    uint64_t jitter_seed;       /* when jitter_mode == HASH */
    int clamp_mode;             /* ide3d_clamp */
    int last_back, white_back;
    float max_depth;            /* 0 = off (volumetric_rendering.py:66) */
    int fill_weight;            /* fill_mode == 'weight' (:71-72) */
    float noise_std;            /* with noise != NULL: sigma += noise_std * noise (:45) */
    const float* noise;         /* [N, R, S] standard normal, or NULL */
    float* out_feat;            /* [N, R, 51]  composited colour features + semantic logits */
    float* out_depth;           /* [N, R] */
    float* out_weights;         /* [N, R, S] or NULL */
    int precision;              /* ide3d_precision: how the decoder MLP is evaluated */
} ide3d_raymarch_params;
int ide3d_raymarch_fwd(const ide3d_raymarch_params* p, ide3d_stream_t stream);

/* Backward of ide3d_raymarch_fwd: what autograd derives in the reference from grid_sample (grid_sample_gradfix.py:55-61), the decoder
 * layers and fancy_integration (volumetric_rendering.py:34-74), in one kernel that recomputes the per-sample chain instead of storing it.
 * p: the forward's parameters (outputs ignored).  grad_feat [N,R,51], grad_depth [N,R] or NULL.  grad_tex / grad_seg: fp32 buffers with
 * the planes' (channels-last) strides, ZERO-INITIALISED by the caller, accumulated into with red.global.add; either may be NULL.
 * grad_params: NULL, or 12 pointers {dW1, db1, dW2, db2} x 3 heads (dense, the heads' shapes), zero-initialised, accumulated into.
 * No gradient for cam2world / jitter / noise.  IDE3D_UNSUPPORTED for decoders other than the three-head one, NCHW planes, > 256 samples. */
int ide3d_raymarch_bwd(const ide3d_raymarch_params* p, const float* grad_feat, const float* grad_depth, float* grad_tex,
                       float* grad_seg, float* const* grad_params, ide3d_stream_t stream);

/* sample_voxel: decode `points` [N, P, 3] (world units) -> out [N, P, 52], or, with sigma_only,
 * out [N, P] holding channel 51 only. */
int ide3d_sample_voxel(const ide3d_triplane* tex, const ide3d_triplane* seg, const ide3d_decoder* dec,
                       const float* points, int64_t num_points, float box_scale, int sigma_only,
                       float* out, ide3d_stream_t stream);

/* sigma grid for extract_shapes: generates the points of 0.9 * create_samples(N, origin, cube) in
 * the kernel (including the float-division index quirk, extract_shapes.py:84-86) for flat voxel
 * indices [first, first + count) of every batch item and writes sigma to out [n, count]. */
int ide3d_sigma_grid(const ide3d_triplane* tex, const ide3d_triplane* seg, const ide3d_decoder* dec,
                     int grid_n, const float voxel_origin[3], float cube_length, float pre_scale,
                     float box_scale, int64_t first, int64_t count, float* out, ide3d_stream_t stream);

/* [N, C, H, W] (any strides) -> dense [N, H, W, C] float32. */
int ide3d_planes_to_nhwc(const float* src, int n, int c, int h, int w, int64_t stride_n, int64_t stride_c,
                         int64_t stride_h, int64_t stride_w, float* dst, ide3d_stream_t stream);

/* ---- stand-alone stages (the reference's free functions on materialised tensors) ---- */

/* get_initial_rays_trig: points [n,R,S,3], z_vals [n,R,S], rays_d_cam [n,R,3]. */
int ide3d_initial_rays(int n, int num_steps, float fov_deg, int res_w, int res_h, float ray_start,
                       float ray_end, float* points, float* z_vals, float* rays_d_cam,
                       ide3d_stream_t stream);

/* perturb_points + the camera transform of transform_sampled_points.  u may be NULL (no jitter).
 * In: points [n,R,S,3], z_vals [n,R,S], dirs [n,R,3], cam2world [n,16].
 * Out: points_world [n,R,S,3], z_out [n,R,S], dirs_world [n,R,3], origins [n,R,3]. */
int ide3d_transform_points(const float* points, const float* z_vals, const float* dirs, const float* u,
                           const float* cam2world, int n, int num_rays, int num_steps, float* points_world,
                           float* z_out, float* dirs_world, float* origins, ide3d_stream_t stream);

/* sample_from_triplane: coords [N, P, 3] in grid units -> out [N*P, 32]. */
int ide3d_sample_triplane(const ide3d_triplane* planes, const float* coords, int64_t num_points,
                          float* out, ide3d_stream_t stream);

/* fancy_integration on a materialised rgb_sigma [n,R,S,C] (sigma = last channel).
 * Out: rgb [n,R,C-1], depth [n,R], weights [n,R,S]. */
int ide3d_integrate(const float* rgb_sigma, const float* rays_d_cam, const float* z_vals, const float* noise,
                    float noise_std, int n, int num_rays, int num_steps, int channels, int clamp_mode,
                    int last_back, int white_back, float max_depth, int fill_weight, float* rgb,
                    float* depth, float* weights, ide3d_stream_t stream);

/* sample_pdf: bins [R, S+1], weights [R, S], u [R, n_imp] -> samples [R, n_imp]. */
int ide3d_sample_pdf(const float* bins, const float* weights, const float* u, int num_rays, int num_bins,
                     int n_importance, float eps, float* samples, ide3d_stream_t stream);

/* mask2color (dnnlib/seg_tools.py:75-82): per pixel argmax over the c semantic logits (first maximum wins, like
 * torch.argmax) and colour look-up, one pass.  masks [n, c, h, w] float32 with element strides; lut [c, 3] float32 (the
 * COLOR_MAP rows, seg_tools.py:13-32); out [n, 3, h, w] float32 dense NCHW -- or, with out_u8 != 0, uint8 (the
 * `.to(torch.uint8)` that follows in gen_videos.py:24-38 folded in). */
int ide3d_mask2color(const float* masks, int n, int c, int h, int w, int64_t stride_n, int64_t stride_c, int64_t stride_h,
                     int64_t stride_w, const float* lut, void* out, int out_u8, ide3d_stream_t stream);

/* Marching cubes on a density grid (render_mesh.py:30-32 / dnnlib/geometry.py:282-286 call PyMCubes' marching_cubes on the host).
 * volume [nx, ny, nz] fp32 dense (index (x*ny + y)*nz + z); a corner is inside where value >= threshold.  Tables (device memory) come
 * from ide3d_b200/mesh.py::build_tables: ntri_table [256] int32, tri_table [256*16] int8 (edge triples, -1 terminated), edge_corner
 * [12*2] int32 (lower / upper corner of each cube edge).
 *   ide3d_mc_classify: counts[cell] = number of triangles of the cell, cell = (x*(ny-1) + y)*(nz-1) + z.
 *   ide3d_mc_emit:     offsets = inclusive scan of counts (int64); per emitted vertex k of triangle t: edge_ids[3t+k] = global id of the
 *                      cut grid edge ((lower corner flat index)*3 + axis), verts[(3t+k)*3 ..] = its position in index units. */
int ide3d_mc_classify(const float* volume, int nx, int ny, int nz, float threshold, const int* ntri_table, unsigned char* counts,
                      ide3d_stream_t stream);
int ide3d_mc_emit(const float* volume, int nx, int ny, int nz, float threshold, const signed char* tri_table, const int* edge_corner,
                  const unsigned char* counts, const int64_t* offsets, int64_t* edge_ids, float* verts, ide3d_stream_t stream);

/* Style vectors and demodulation coefficients of every modulated convolution of one synthesis call, two launches
 * (replaces per layer: FullyConnectedLayer.forward of the affine, inversion/networks.py:136-165 / :476, and the dcoefs
 * reduction of modulated_conv2d, :89-90).
 *   styles[style_off + n*in_ch + i] = ((affine_w[i,:] . ws[n, w_index, :]) * w_gain + affine_b[i] * b_gain) * out_scale
 *   dcoefs[dcoef_off + n*out_ch + o] = rsqrt(sum_i styles[n,i]^2 * wsq[o,i] + 1e-8)        (layers with wsq != NULL)
 * ws [n, num_ws, w_dim] fp32 dense; wsq[o,i] = sum over the kernel taps of weight[o,i,:,:]^2 (a constant of the weights). */
typedef struct ide3d_style_layer {
    const float* affine_w;      /* [in_ch, w_dim] */
    const float* affine_b;      /* [in_ch] or NULL */
    const float* wsq;           /* [out_ch, in_ch] or NULL (no demodulation: ToRGB) */
    float w_gain, b_gain;       /* FullyConnectedLayer runtime gains */
    float out_scale;            /* extra factor on the style (ToRGB weight_gain), 1 otherwise */
    int w_index;                /* which ws[:, w_index] feeds the layer */
    int in_ch, out_ch;
    int64_t style_off, dcoef_off;
} ide3d_style_layer;
int ide3d_style_plan(const float* ws, int n, int num_ws, int w_dim, const ide3d_style_layer* layers, int num_layers,
                     float* styles, float* dcoefs, ide3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* IDE3D_B200_H_ */
