// Video pre-processing on the device: the rescale + normalise + patchify stage of the HF Qwen2-VL
// video processor that the reference runs on the host CPU for every sample
// (src/time_r1/rl/timer1_trainer.py:547-556 -> HF models/qwen2_vl/video_processing_qwen2_vl.py:236-272; frames come from
// fetch_video_v3, src/utils/vision_process.py:440-475, as [T, 3, H, W] with integer values 0..255).
//
//   out[row, col] = (frame[t, c, y, x] - mean255[c]) / std255[c]          (fp32, IEEE subtract then divide, as tvF.normalize)
//   row = (grid_t, h/2, w/2, merge_h, merge_w),  col = (c, temporal, patch_h, patch_w)   (permute 0,1,4,7,5,8,3,2,6,9)
// with mean255 = fp32(mean) * 255, std255 = fp32(std) * 255 (the fused rescale-and-normalise of the HF processor) and an odd
// frame count padded by repeating the last frame.  Outputs, either or both:
//   out_f32  [N, 1176] in HF row order  == pixel_values_videos, bit-identical to the HF processor;
//   out_bf16 [N, 1176] with rows gathered through `perm_rows` (the ViT window order) == what the patch-embed GEMM reads,
//            so the fp32 [N, 1176] tensor (77 MB at cfg2) and its H2D copy never exist: frames travel as uint8.
// HBM-bound: reads T*3*H*W pixels once, writes N*1176 values once.
#include "t1_common.cuh"

namespace {

template <typename TIn>
__global__ void video_patchify_kernel(const TIn* __restrict__ frames, int T, int H, int W, int patch, int tpatch, int merge,
                                      float m0, float m1, float m2, float s0, float s1, float s2,
                                      const int* __restrict__ perm_rows, float* __restrict__ out_f32,
                                      bf16* __restrict__ out_bf16, long long n_rows, int row_len) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows * row_len) return;
  const long long orow = idx / row_len;
  const int col = (int)(idx - orow * row_len);
  const long long row = perm_rows ? (long long)perm_rows[orow] : orow;
  const int gh = H / patch, gw = W / patch;
  const int hb_n = gh / merge, wb_n = gw / merge;
  // row -> (grid_t, hb, wb, mh, mw)
  long long r = row;
  const int mw = (int)(r % merge); r /= merge;
  const int mh = (int)(r % merge); r /= merge;
  const int wb = (int)(r % wb_n); r /= wb_n;
  const int hb = (int)(r % hb_n); r /= hb_n;
  const int gt = (int)r;
  // col -> (c, tp, ph, pw)
  int cc = col;
  const int pw = cc % patch; cc /= patch;
  const int ph = cc % patch; cc /= patch;
  const int tp = cc % tpatch; cc /= tpatch;
  const int c = cc;
  int t = gt * tpatch + tp;
  if (t >= T) t = T - 1;                                     // odd frame counts repeat the last frame
  const int y = (hb * merge + mh) * patch + ph, x = (wb * merge + mw) * patch + pw;
  const float v = (float)frames[(((long long)t * 3 + c) * H + y) * W + x];
  const float m = c == 0 ? m0 : (c == 1 ? m1 : m2), s = c == 0 ? s0 : (c == 1 ? s1 : s2);
  const float o = __fdiv_rn(__fsub_rn(v, m), s);
  if (out_f32) out_f32[row * row_len + col] = o;             // HF order (each row written exactly once: perm is a permutation)
  if (out_bf16) out_bf16[orow * row_len + col] = __float2bfloat16(o);
}

}  // namespace

// frames: [T, 3, H, W] device tensor, uint8 (in_u8 != 0) or fp32, values 0..255; H, W multiples of patch * merge.
// mean255 / std255: the three per-channel constants already scaled by 255 in fp32 (host).  perm_rows (nullable, device
// [N] int32): output row i of out_bf16 holds HF row perm_rows[i].  N = ceil(T / tpatch) * (H / patch) * (W / patch).
extern "C" int t1_video_patchify(const void* frames, int in_u8, int T, int H, int W, int patch, int tpatch, int merge,
                                 const float* mean255, const float* std255, const int* perm_rows, float* out_f32,
                                 void* out_bf16, cudaStream_t stream) {
  T1_CHECK_ARG(frames && mean255 && std255 && (out_f32 || out_bf16), "t1_video_patchify: null pointer");
  T1_CHECK_ARG(T > 0 && patch > 0 && tpatch > 0 && merge > 0 && H > 0 && W > 0 && H % (patch * merge) == 0 && W % (patch * merge) == 0,
               "t1_video_patchify: H=%d W=%d must be positive multiples of patch*merge=%d", H, W, patch * merge);
  const long long gt = (T + tpatch - 1) / tpatch;
  const long long n_rows = gt * (H / patch) * (W / patch);
  const int row_len = 3 * tpatch * patch * patch;
  const long long total = n_rows * row_len;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (in_u8)
    video_patchify_kernel<unsigned char><<<blocks, 256, 0, stream>>>((const unsigned char*)frames, T, H, W, patch, tpatch, merge,
                                                                    mean255[0], mean255[1], mean255[2], std255[0], std255[1], std255[2],
                                                                    perm_rows, out_f32, (bf16*)out_bf16, n_rows, row_len);
  else
    video_patchify_kernel<float><<<blocks, 256, 0, stream>>>((const float*)frames, T, H, W, patch, tpatch, merge, mean255[0],
                                                            mean255[1], mean255[2], std255[0], std255[1], std255[2], perm_rows,
                                                            out_f32, (bf16*)out_bf16, n_rows, row_len);
  T1_CUDA_LAUNCH_CHECK("t1_video_patchify");
  return T1_OK;
}
