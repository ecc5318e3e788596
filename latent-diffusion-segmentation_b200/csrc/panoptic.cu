// Panoptic post-processing of the decoded logits on the GPU.
//
// Replaces the per-image tail of TrainerDiffusion.compute_pq
// (/root/reference/ldmseg/trainers/trainers_ldm_cond.py:1243-1313): the reference materialises the
// [B,128,512,512] f32 logits (134 MB / image), bilinear-resizes them per image to the original (h, w), takes
// argmax / softmax-max / sigmoid on the device, copies the FULL [128,h,w] sigmoid map plus the ids to the host and
// filters segments in numpy.  Here the seg decoder's 256x256 logits (before its own bilinear x2, vae.py:270) are
// resampled once through the composition of both bilinear maps, the per-pixel decisions are taken in registers,
// the two per-class area histograms are accumulated on the fly, and a second tiny kernel applies the
// count_th / overlap_th rules -- the host receives uint8 ids and a 128-entry table.
//
//   panoptic_resample : logits -> pred (int16, -1 = below mask_th), area[c] = #pixels with argmax c (after the
//                       threshold), orig_area[c] = #pixels with sigmoid(logit_c) >= mask_th        (:1264-1289, 1301)
//   panoptic_filter   : keep[c] per rules (:1293-1304), ids = keep[pred] ? pred + 1 : 0               (:1296-1313)
#include "common.h"
#include "ptx.cuh"
#include "../../include/ldmseg_b200.h"

namespace ldm {

// PyTorch's area_pixel_compute_source_index (align_corners = False): src = scale * (dst + 0.5) - 0.5, clamped at 0
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l1) {
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = static_cast<int>(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - static_cast<float>(i0);
}

// value of the x2-upsampled logits (vae.py:270, scale_factor = 2 -> scale 0.5) at pixel (Y, X) of the 2s x 2s grid,
// four channels per lane
__device__ __forceinline__ float4 sample_x2(const float* __restrict__ img, int s, int ld, int Y, int X, int c4) {
  int y0, y1, x0, x1;
  float ly, lx;
  src_index(0.5f, Y, s, y0, y1, ly);
  src_index(0.5f, X, s, x0, x1, lx);
  const float4 v00 = __ldg(reinterpret_cast<const float4*>(img + (static_cast<size_t>(y0) * s + x0) * ld) + c4);
  const float4 v01 = __ldg(reinterpret_cast<const float4*>(img + (static_cast<size_t>(y0) * s + x1) * ld) + c4);
  const float4 v10 = __ldg(reinterpret_cast<const float4*>(img + (static_cast<size_t>(y1) * s + x0) * ld) + c4);
  const float4 v11 = __ldg(reinterpret_cast<const float4*>(img + (static_cast<size_t>(y1) * s + x1) * ld) + c4);
  const float hy0 = 1.f - ly, hx0 = 1.f - lx;
  float4 o;
  o.x = hy0 * (hx0 * v00.x + lx * v01.x) + ly * (hx0 * v10.x + lx * v11.x);
  o.y = hy0 * (hx0 * v00.y + lx * v01.y) + ly * (hx0 * v10.y + lx * v11.y);
  o.z = hy0 * (hx0 * v00.z + lx * v01.z) + ly * (hx0 * v10.z + lx * v11.z);
  o.w = hy0 * (hx0 * v00.w + lx * v01.w) + ly * (hx0 * v10.w + lx * v11.w);
  return o;
}

// geom per image: {h, w, crop_y0, crop_x0, crop_h, crop_w} -- target size and the padding crop on the 2s x 2s grid.
// One warp per output pixel, lane = 4 consecutive classes (c = 128).  grid (ceil(max_hw / 8), nb), block 256.
__global__ void __launch_bounds__(256)
panoptic_resample_kernel(const float* __restrict__ logits, int s, int ld, const int* __restrict__ geom,
                         int out_stride, float mask_th, int threshold_output, int16_t* __restrict__ pred,
                         int* __restrict__ area, int* __restrict__ orig_area) {
  pdl_sync();
  __shared__ int s_area[128], s_orig[128];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 128; i += blockDim.x) s_area[i] = s_orig[i] = 0;
  __syncthreads();
  const int h = geom[b * 6 + 0], w = geom[b * 6 + 1];
  const int cy0 = geom[b * 6 + 2], cx0 = geom[b * 6 + 3], ch = geom[b * 6 + 4], cw = geom[b * 6 + 5];
  const int lane = threadIdx.x & 31;
  const int pix = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix < h * w) {
    const int oy = pix / w, ox = pix - oy * w;
    const float* img = logits + static_cast<size_t>(b) * s * s * ld;
    // second resize (:1267-1272): cropped 2s-grid -> (h, w), scale = in / out
    int Y0, Y1, X0, X1;
    float LY, LX;
    src_index(static_cast<float>(ch) / static_cast<float>(h), oy, ch, Y0, Y1, LY);
    src_index(static_cast<float>(cw) / static_cast<float>(w), ox, cw, X0, X1, LX);
    const float4 a00 = sample_x2(img, s, ld, cy0 + Y0, cx0 + X0, lane);
    const float4 a01 = sample_x2(img, s, ld, cy0 + Y0, cx0 + X1, lane);
    const float4 a10 = sample_x2(img, s, ld, cy0 + Y1, cx0 + X0, lane);
    const float4 a11 = sample_x2(img, s, ld, cy0 + Y1, cx0 + X1, lane);
    const float HY0 = 1.f - LY, HX0 = 1.f - LX;
    float v[4];
    v[0] = HY0 * (HX0 * a00.x + LX * a01.x) + LY * (HX0 * a10.x + LX * a11.x);
    v[1] = HY0 * (HX0 * a00.y + LX * a01.y) + LY * (HX0 * a10.y + LX * a11.y);
    v[2] = HY0 * (HX0 * a00.z + LX * a01.z) + LY * (HX0 * a10.z + LX * a11.z);
    v[3] = HY0 * (HX0 * a00.w + LX * a01.w) + LY * (HX0 * a10.w + LX * a11.w);
    // argmax (first maximum, as torch.argmax) and the softmax denominator
    float best = v[0];
    int bi = lane * 4;
#pragma unroll
    for (int i = 1; i < 4; ++i)
      if (v[i] > best) {
        best = v[i];
        bi = lane * 4 + i;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    float sum = expf(v[0] - best) + expf(v[1] - best) + expf(v[2] - best) + expf(v[3] - best);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    int label = bi;
    if (threshold_output && 1.f / sum < mask_th) label = -1;             // :1276-1284
    // sigmoid(logit_c) >= mask_th, evaluated like torch.sigmoid (:1288, 1301)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (1.f / (1.f + expf(-v[i])) >= mask_th) atomicAdd(&s_orig[lane * 4 + i], 1);
    if (lane == 0) {
      pred[static_cast<size_t>(b) * out_stride + pix] = static_cast<int16_t>(label);
      if (label >= 0) atomicAdd(&s_area[label], 1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    if (s_area[i]) atomicAdd(&area[b * 128 + i], s_area[i]);
    if (s_orig[i]) atomicAdd(&orig_area[b * 128 + i], s_orig[i]);
  }
}

// keep[c] (:1293-1304): drop a segment if its area is below count_th, if it is the ignore label, or if
// area / orig_area < overlap_th (numpy int / int -> float64 true division, compared with a Python float);
// then ids = keep[pred] ? pred + 1 : 0 (:1296-1313).  grid (chunks, nb).
__global__ void panoptic_filter_kernel(const int16_t* __restrict__ pred, const int* __restrict__ geom,
                                       int out_stride, const int* __restrict__ area,
                                       const int* __restrict__ orig_area, int count_th, double overlap_th,
                                       int ignore_label, uint8_t* __restrict__ ids, int* __restrict__ keep) {
  pdl_sync();
  __shared__ int s_keep[128];
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < 128; c += blockDim.x) {
    const int a = area[b * 128 + c];
    int k = a > 0 && a >= count_th && c != ignore_label;
    if (k && static_cast<double>(a) / static_cast<double>(orig_area[b * 128 + c]) < overlap_th) k = 0;
    s_keep[c] = k;
    if (blockIdx.x == 0) keep[b * 128 + c] = k;
  }
  __syncthreads();
  const int n = geom[b * 6 + 0] * geom[b * 6 + 1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int l = pred[static_cast<size_t>(b) * out_stride + i];
    ids[static_cast<size_t>(b) * out_stride + i] = (l >= 0 && s_keep[l]) ? static_cast<uint8_t>(l + 1) : 0;
  }
}

}  // namespace ldm

using namespace ldm;

extern "C" int ldmseg_panoptic_resample(const float* logits, int nb, int s, int c, int ld, const int* geom_dev,
                                        int max_hw, int out_stride, float mask_th, int threshold_output,
                                        int16_t* pred, int* area, int* orig_area, void* stream) {
  LDM_REQUIRE(logits && geom_dev && pred && area && orig_area, "panoptic_resample: null pointer");
  LDM_REQUIRE(c == 128 && ld % 4 == 0 && ld >= c, "panoptic_resample: 128 classes (the reference's out_channels), ld %% 4 == 0");
  LDM_REQUIRE((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "panoptic_resample: logits not 16-byte aligned");
  LDM_REQUIRE(nb > 0 && s > 0 && max_hw > 0 && out_stride >= max_hw, "panoptic_resample: bad geometry");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LDM_CUDA(cudaMemsetAsync(area, 0, sizeof(int) * 128 * nb, st));
  LDM_CUDA(cudaMemsetAsync(orig_area, 0, sizeof(int) * 128 * nb, st));
  launch_kernel(panoptic_resample_kernel, dim3((max_hw + 7) / 8, nb), dim3(256), 0, st, logits, s, ld, geom_dev,
                out_stride, mask_th, threshold_output, pred, area, orig_area);
  return check_launch("panoptic_resample_kernel");
}

extern "C" int ldmseg_panoptic_filter(const int16_t* pred, int nb, const int* geom_dev, int max_hw, int out_stride,
                                      const int* area, const int* orig_area, int count_th, double overlap_th,
                                      int ignore_label, uint8_t* ids, int* keep, void* stream) {
  LDM_REQUIRE(pred && geom_dev && area && orig_area && ids && keep, "panoptic_filter: null pointer");
  LDM_REQUIRE(nb > 0 && max_hw > 0 && out_stride >= max_hw, "panoptic_filter: bad geometry");
  int chunks = (max_hw + 256 * 8 - 1) / (256 * 8);
  if (chunks > 4 * num_sms()) chunks = 4 * num_sms();
  launch_kernel(panoptic_filter_kernel, dim3(chunks, nb), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), pred,
                geom_dev, out_stride, area, orig_area, count_th, overlap_th, ignore_label, ids, keep);
  return check_launch("panoptic_filter_kernel");
}
