// Fused '1x1conv' adaptation + FGD loss passes on the tensor cores (see adapt_loss_tc.cu).
#pragma once

#include "common.cuh"

namespace dbev {

// Arrays of the FGD loss state the fused kernel reads / writes (all device pointers; [B,HW] maps are indexed b*HW + cell,
// [B,C] vectors b*C + channel, chan_p is [B, C, ntiles] with ntiles = ceil(HW / 128)).
struct AdaptFgdArgs {
  const float* bias;      // [C] adaptation conv bias (nullable)
  const float* teacher;   // [B, C, HW] (NCHW)
  const float* catt;      // [B, C] teacher channel attention
  float* chan_p;          // per-tile channel sums of the adapted student (mode 0) / of d loss / d adapted (mode 1); nullable
  // mode 0 (forward): per-cell sums over the channels of the adapted student s
  float *sa, *sm, *d1, *d2;   // sum |s|, sum s, sum (s-t)^2, sum catt (s-t)^2
  // mode 1 (backward): ds = (s - t) (Wa + catt Wb) + gc + gsp
  const float *fgw, *bgw, *fpw, *gsp;   // [B, HW]
  const float* gc;                      // [B, C]
  const float* grad_losses;             // [5]
  float w_fg, w_bg, w_fp;
  int use_fp, channel_mask;
  float* ds_cl;           // [B, HW, C] channels-last output of mode 1
};

// s = x W^T + bias for every 128-cell tile on tcgen05 (TF32), consumed in the epilogue; the adapted map never
// reaches global memory. x_cl [B, HW, C_in] channels-last, w [C_out, C_in]. C_in % 32 == 0; C_out % 32 == 0 up to
// 256, or % 64 == 0 up to 512; HW % 4 == 0.
int adapt_fgd_fused(int mode, const float* x_cl, const float* w, int batch, int c_in, int c_out, int hw,
                    const AdaptFgdArgs& args, cudaStream_t stream);

bool adapt_fgd_fused_supports(int c_in, int c_out, int hw);

}  // namespace dbev
