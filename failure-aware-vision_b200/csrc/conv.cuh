// conv.cuh -- launcher interface of the tcgen05 implicit-GEMM convolution (conv.cu).
#pragma once
#include "common.cuh"

namespace fav {

struct ConvLayer {          // device-resident, BN folded
  const __nv_bfloat16* w = nullptr;   // [cout_pad][kpad]  (k = (r*S + s)*cin + c), zero padded
  const float* bias = nullptr;        // [cout_pad]
  int cin = 0, cout = 0, r = 0, s = 0, stride = 1, pad = 0;
  int cin_store = 0, s_store = 0;      // stem layout: channels padded to 4, filter-row slots padded to even (0 = as cin / s)
  int s2d = 0;                         // stem as a 4x4/s1 conv over the 2x2 space-to-depth input (16 ch): k = (r'*4 + s')*16 + (dy*2 + dx)*4 + c
  int k = 0, kpad = 0, cout_pad = 0, bn = 0;
  int fold = 0;                           // 4: a 3x3/s1/p1 conv on 2x2 images folded into a dense 1x1 GEMM (cin, cout already x4)
  int k2pad = 0, cin2 = 0, stride2 = 1;   // fused 1x1 downsample branch: extra K columns [kpad, kpad + k2pad) of w, its Cin and stride
  alignas(64) unsigned char tmap_w[128];   // CUtensorMap for the weights (box 64 x bn, SWIZZLE_128B)
  bool tmap_ok = false;
  alignas(64) unsigned char tmap_w64[128];  // same weights with a 64-row box (one CTA's half in the 2-SM variant; narrow N tiles)
  bool tmap64_ok = false;
  alignas(64) unsigned char tmap_w32[128];  // same weights with a 32-row box (narrow N tiles of small launches)
  bool tmap32_ok = false;
};

struct ConvCall {
  const ConvLayer* L = nullptr;
  const void* x = nullptr;        // bf16 NHWC [p, h, w, cin]
  void* y = nullptr;              // bf16 / fp32 NHWC [p*rep, oh, ow, cout]
  const void* res = nullptr;      // bf16 NHWC [p, oh, ow, cout] or null
  const void* x2 = nullptr;       // input of the fused downsample branch, bf16 NHWC [p, h2, w2, cin2]
  int h2 = 0, w2 = 0;
  int p = 0, h = 0, w = 0;
  int relu = 0, out_f32 = 0;
  int force_mt = 0;               // 0 auto, 1 / 2: 128- or 256-pixel CTA tiles (a_mode 0 only)
  int a_mode = -1;                // -1 auto, 0 TMA-tiled, 1 vector gather, 2 scalar gather, 3 channel-padded stem gather,
                                  // 4 flat-padded resident 3x3 (conv_flat.cu), 5 TMA on the space-to-depth stem input
                                  // (x = [p][hp][wp][16], see conv_stem_padded_dims)
  // MC-dropout in the epilogue
  int T = 1;                      // passes; rows of x are pass-images (image = row / T, t = row % T) unless rep > 1
  int rep = 1;                    // rep == T: x holds plain images, each output row is written T times with mask t
  int drop = 0;
  float p_drop = 0.f;
  uint64_t seed = 0, first_image = 0;
  int layer_id = 0;
  int drop2_layer = -1;           // >= 0: a second, independent mask + scale with this layer id on the same elements (the dropout
                                  // before fc when the feature map is 1x1, so the pooled feature IS this output); needs drop, rep == 1
};

int conv_out_dim(int in, int k, int stride, int pad);
int conv_pick_bn(int cout);
// fills k/kpad/cout_pad/bn and encodes the weight tensor map (weights must already be on the device)
int conv_layer_finalize(ConvLayer& L);
int conv_launch(Ctx* ctx, const ConvCall& c, cudaStream_t st);
bool conv_stem_padded_dims(const ConvLayer& L, int h, int w, int* hp, int* wp);
int conv_timing_begin(Ctx* ctx, cudaStream_t st, float gflop, float gbyte, cudaEvent_t* stop, unsigned long long** stats);
// conv_flat.cu: 3x3 / stride 1 / pad 1, Cin = Cout = 64 on small images: activations resident in shared memory as a
// flat zero-padded pixel list, weights resident, filter taps = shifted UMMA descriptors
bool conv_flat_applicable(const ConvCall& c);
int conv_flat_launch(Ctx* ctx, const ConvCall& c, cudaStream_t st);

}  // namespace fav
