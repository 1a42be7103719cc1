// VGG front-end of the VGG-BLSTM encoder for sm_100a: forward and backward.
//
// Replaces VGGBLSTMEncoder.__call__ up to the BLSTM (models/encoders/core/vgg_blstm.py:93-177)
// and conv_layer / max_pool (models/encoders/core/cnn_util.py:52-84, 13-29):
//   [N=B*T, H=num_channels, W=splice*num_stack, 3] -> conv3x3(3->64)+ReLU -> dropout ->
//   conv(64->64)+ReLU -> pool 2x2/2 SAME -> dropout -> conv(64->128)+ReLU -> dropout ->
//   conv(128->128)+ReLU -> pool -> dropout -> flatten -> FC 256 + ReLU -> dropout.
//
// Layout.  Every activation lives in HBM as a zero-bordered NHWC image, flattened to rows of C
// channels: row index c = (n*Hp + hp)*Wp + wp with Hp = H+2, Wp = W+2 (Wp = W when W == 1:
// a 1-wide image needs no horizontal taps, two thirds of TF's multiplies there hit padding).
// With that border the 3x3 convolution is three GEMMs, one per kernel row dh, whose A operand is
// the activation buffer itself read with OVERLAPPING rows (row stride C, K = 3C): the three
// horizontal taps of one kernel row are adjacent in memory, so im2col costs no traffic.  The GEMM
// output ("frame" row r = top-left aligned) and the next layer's input ("centred") are the same
// buffer shifted by lead = Wp + (W>1) rows, so the epilogue (bias, ReLU, dropout, zeroing of the
// border rows) is a single elementwise pass, and in the backward pass one gradient buffer serves
// as the dY operand of the weight-gradient GEMM (frame view) and as the zero-bordered input of
// the data-gradient convolution with flipped filters (centred view).
#include "common.cuh"
#include "lstm_internal.cuh"

namespace b2 {

struct Geo {
  int N, H, W, pw, Hp, Wp, lead;
  int64_t M;           // frame rows = N*Hp*Wp
  int64_t rows;        // allocated rows (M + slack for the shifted reads)
};
static Geo make_geo(int N, int H, int W) {
  Geo g;
  g.N = N; g.H = H; g.W = W; g.pw = W > 1 ? 1 : 0; g.Hp = H + 2; g.Wp = W + 2 * g.pw;
  g.lead = g.Wp + g.pw;
  g.M = (int64_t)N * g.Hp * g.Wp;
  g.rows = g.M + 3 * g.Wp + 8;
  return g;
}

__device__ __forceinline__ bool frame_valid(const Geo& g, int64_t r, int& n, int& h, int& w) {
  if (r < 0 || r >= g.M) return false;
  w = (int)(r % g.Wp);
  const int64_t q = r / g.Wp;
  h = (int)(q % g.Hp);
  n = (int)(q / g.Hp);
  return w < g.W && h < g.H;
}

__device__ __forceinline__ float drop_scale(uint64_t seed, uint64_t idx, float keep) {
  if (keep >= 1.f) return 1.f;
  return dropout_keep(seed, idx, keep) ? 1.f / keep : 0.f;
}

// x [N,H,W,C] dense -> centred zero-bordered buffer
__global__ void __launch_bounds__(256)
vgg_pack_kernel(const float* __restrict__ x, Geo g, int C, float* __restrict__ P) {
  const int64_t total = g.rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i % C);
    int n, h, w;
    float v = 0.f;
    if (frame_valid(g, row - g.lead, n, h, w)) v = x[(((int64_t)n * g.H + h) * g.W + w) * C + c];
    P[i] = v;
  }
}

// non-pool epilogue: P[r + lead] = valid ? drop(relu(Y[r] + b)) : 0
__global__ void __launch_bounds__(256)
vgg_epi_kernel(const float* __restrict__ Y, const float* __restrict__ bias, Geo g, int C, float keep,
               uint64_t seed, float* __restrict__ P) {
  const int64_t total = g.rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i % C);
    const int64_t r = row - g.lead;
    int n, h, w;
    float v = 0.f;
    if (frame_valid(g, r, n, h, w)) {
      v = fmaxf(Y[r * C + c] + bias[c], 0.f);
      v *= drop_scale(seed, (((uint64_t)n * g.H + h) * g.W + w) * C + c, keep);
    }
    P[i] = v;
  }
}

// pooled epilogue: out (centred buffer of geometry g2, or dense [N,H2,W2,C]) = drop(maxpool(relu(Y+b)))
__global__ void __launch_bounds__(256)
vgg_epi_pool_kernel(const float* __restrict__ Y, const float* __restrict__ bias, Geo g, Geo g2, int C,
                    float keep, uint64_t seed, int dense, float* __restrict__ out) {
  const int64_t total = (dense ? (int64_t)g2.N * g2.H * g2.W : g2.rows) * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i % C);
    int n, h2, w2;
    bool ok;
    if (dense) { w2 = (int)(row % g2.W); h2 = (int)((row / g2.W) % g2.H); n = (int)(row / ((int64_t)g2.W * g2.H)); ok = true; }
    else ok = frame_valid(g2, row - g2.lead, n, h2, w2);
    float v = 0.f;
    if (ok) {
      float m = 0.f;                                   // relu output is >= 0
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int h = 2 * h2 + a, w = 2 * w2 + b;
          if (h < g.H && w < g.W)
            m = fmaxf(m, Y[(((int64_t)n * g.Hp + h) * g.Wp + w) * C + c] + bias[c]);
        }
      v = m * drop_scale(seed, (((uint64_t)n * g2.H + h2) * g2.W + w2) * C + c, keep);
    }
    out[i] = v;
  }
}

// backward of the non-pool epilogue: D[r + lead] = valid ? G[r] * (P[r+lead] > 0 ? 1/keep : 0) : 0
__global__ void __launch_bounds__(256)
vgg_epi_bwd_kernel(const float* __restrict__ G, const float* __restrict__ P, Geo g, int C, float keep,
                   float* __restrict__ D) {
  const int64_t total = g.rows * C;
  const float sc = keep < 1.f ? 1.f / keep : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i % C);
    const int64_t r = row - g.lead;
    int n, h, w;
    float v = 0.f;
    if (frame_valid(g, r, n, h, w) && P[i] > 0.f) v = G[r * C + c] * sc;
    D[i] = v;
  }
}

// backward of the pooled epilogue.  G2: gradient wrt the pooled output (frame rows of g2, or dense);
// out2: the pooled output itself (centred buffer of g2, or dense) -- its zeros carry the dropout mask.
// The gradient goes to the first arg-max of each window (tf.nn.max_pool), if ReLU was active there.
__global__ void __launch_bounds__(256)
vgg_epi_pool_bwd_kernel(const float* __restrict__ G2, const float* __restrict__ out2,
                        const float* __restrict__ Y, const float* __restrict__ bias, Geo g, Geo g2,
                        int C, float keep, int dense, float* __restrict__ D) {
  const int64_t total = g.rows * C;
  const float sc = keep < 1.f ? 1.f / keep : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i % C);
    const int64_t r = row - g.lead;
    int n, h, w;
    float v = 0.f;
    if (frame_valid(g, r, n, h, w)) {
      const float mine = Y[r * C + c] + bias[c];
      if (mine > 0.f) {
        const int h2 = h >> 1, w2 = w >> 1;
        bool first = true;                              // am I the first maximum of my window?
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int hh = 2 * h2 + a, ww = 2 * w2 + b;
            if (hh >= g.H || ww >= g.W || (hh == h && ww == w)) continue;
            const float o = Y[(((int64_t)n * g.Hp + hh) * g.Wp + ww) * C + c] + bias[c];
            const bool before = (hh < h) || (hh == h && ww < w);
            if (o > mine || (before && o == mine)) first = false;
          }
        if (first) {
          int64_t j2, o2;
          if (dense) { j2 = (((int64_t)n * g2.H + h2) * g2.W + w2) * C + c; o2 = j2; }
          else {
            const int64_t r2 = ((int64_t)n * g2.Hp + h2) * g2.Wp + w2;
            j2 = r2 * C + c; o2 = (r2 + g2.lead) * C + c;
          }
          if (out2[o2] > 0.f) v = G2[j2] * sc;
        }
      }
    }
    D[i] = v;
  }
}

// Wt[dh][dw][co][ci] = W[2-dh][2-dw][ci][co]
__global__ void __launch_bounds__(256)
vgg_flip_filter_kernel(const float* __restrict__ Wf, int Cin, int Cout, float* __restrict__ Wt) {
  const int total = 9 * Cin * Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % Cin, co = (i / Cin) % Cout, t = i / (Cin * Cout);
    const int dh = t / 3, dw = t % 3;
    Wt[i] = Wf[(((2 - dh) * 3 + (2 - dw)) * Cin + ci) * Cout + co];
  }
}

// FC epilogue forward: out = drop(relu(z)) in place; backward: dz = dout * (out > 0 ? 1/keep : 0)
__global__ void __launch_bounds__(256)
vgg_fc_epi_kernel(float* __restrict__ z, int64_t n, float keep, uint64_t seed) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    z[i] = fmaxf(z[i], 0.f) * drop_scale(seed, (uint64_t)i, keep);
}
__global__ void __launch_bounds__(256)
vgg_fc_epi_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, int64_t n, float keep,
                      float* __restrict__ dz) {
  const float sc = keep < 1.f ? 1.f / keep : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dz[i] = out[i] > 0.f ? dout[i] * sc : 0.f;
}

static int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}


// bf16 tensor-core path (precision == B2_PREC_BF16): every activation / gradient buffer that feeds a GEMM also
// exists as a bf16 shadow with the same row geometry; the 64- and 128-channel convolutions and the bridge FC run on
// gemm_tc_kernel (tcgen05), conv1_1 (3 input channels = 6-byte rows, no TMA) stays on the fp32 CUDA-core GEMM.
struct VggBuf {
  Geo g0, g2, g4;
  __nv_bfloat16* Pb[4];   // shadows of P[1..3] (index 0 unused)
  __nv_bfloat16* Fb;      // shadow of F
  float* P[4];     // P[0] = padded input (3 ch), P[1] = conv1_1 out (64, g0), P[2] = pooled conv1_2 (64, g2),
                   // P[3] = conv2_1 out (128, g2)
  float* Y2;       // conv1_2 GEMM output, frame of g0, 64 ch
  float* Y4;       // conv2_2 GEMM output, frame of g2, 128 ch
  float* F;        // pooled conv2_2, dense [N, H4*W4*128]
  float* fc_out;   // [N,256]
};
static size_t vgg_reserve_layout(const b2_vgg_desc* d, void* base, VggBuf* b) {
  VggBuf t;
  t.g0 = make_geo(d->N, d->H, d->W);
  t.g2 = make_geo(d->N, (d->H + 1) / 2, (d->W + 1) / 2);
  t.g4 = make_geo(d->N, (t.g2.H + 1) / 2, (t.g2.W + 1) / 2);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t o0 = take((size_t)t.g0.rows * 3 * 4), o1 = take((size_t)t.g0.rows * 64 * 4);
  const size_t o2 = take((size_t)t.g2.rows * 64 * 4), o3 = take((size_t)t.g2.rows * 128 * 4);
  const size_t oy2 = take((size_t)t.g0.M * 64 * 4), oy4 = take((size_t)t.g2.M * 128 * 4);
  const size_t of = take((size_t)d->N * t.g4.H * t.g4.W * 128 * 4), oo = take((size_t)d->N * 256 * 4);
  const bool lp = d->precision == B2_PREC_BF16;
  const size_t ob1 = lp ? take((size_t)t.g0.rows * 64 * 2) : 0, ob2 = lp ? take((size_t)t.g2.rows * 64 * 2) : 0;
  const size_t ob3 = lp ? take((size_t)t.g2.rows * 128 * 2) : 0;
  const size_t obf = lp ? take((size_t)d->N * t.g4.H * t.g4.W * 128 * 2) : 0;
  if (b) {
    char* p = (char*)base;
    t.P[0] = (float*)(p + o0); t.P[1] = (float*)(p + o1); t.P[2] = (float*)(p + o2); t.P[3] = (float*)(p + o3);
    t.Y2 = (float*)(p + oy2); t.Y4 = (float*)(p + oy4); t.F = (float*)(p + of); t.fc_out = (float*)(p + oo);
    t.Pb[0] = nullptr;
    t.Pb[1] = lp ? (__nv_bfloat16*)(p + ob1) : nullptr; t.Pb[2] = lp ? (__nv_bfloat16*)(p + ob2) : nullptr;
    t.Pb[3] = lp ? (__nv_bfloat16*)(p + ob3) : nullptr; t.Fb = lp ? (__nv_bfloat16*)(p + obf) : nullptr;
    *b = t;
  }
  return off;
}
// workspace: one frame-sized GEMM output / gradient pair at the largest geometry + flipped filter
struct VggWs { float* Ya; float* Da; float* Ga; float* Wt; float* dz;
               __nv_bfloat16* Dab; __nv_bfloat16* Wb; __nv_bfloat16* dzb; };
static size_t vgg_ws_layout(const b2_vgg_desc* d, void* base, VggWs* w) {
  const Geo g0 = make_geo(d->N, d->H, d->W);
  const Geo g2 = make_geo(d->N, (d->H + 1) / 2, (d->W + 1) / 2);
  size_t big = (size_t)g0.rows * 64 * 4;
  if ((size_t)g2.rows * 128 * 4 > big) big = (size_t)g2.rows * 128 * 4;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t oa = take(big), od = take(big), og = take(big), ow = take((size_t)9 * 128 * 128 * 4);
  const size_t oz = take((size_t)d->N * 256 * 4);
  const bool lp = d->precision == B2_PREC_BF16;
  const Geo g4 = make_geo(d->N, (g2.H + 1) / 2, (g2.W + 1) / 2);
  size_t wb = (size_t)9 * 128 * 128 * 2;                       // packed conv filter or the bridge weights
  if ((size_t)g4.H * g4.W * 128 * 256 * 2 > wb) wb = (size_t)g4.H * g4.W * 128 * 256 * 2;
  const size_t odb = lp ? take(big / 2) : 0, owb = lp ? take(wb) : 0, ozb = lp ? take((size_t)d->N * 256 * 2) : 0;
  if (w) {
    char* p = (char*)base;
    w->Ya = (float*)(p + oa); w->Da = (float*)(p + od); w->Ga = (float*)(p + og); w->Wt = (float*)(p + ow);
    w->dz = (float*)(p + oz);
    w->Dab = lp ? (__nv_bfloat16*)(p + odb) : nullptr; w->Wb = lp ? (__nv_bfloat16*)(p + owb) : nullptr;
    w->dzb = lp ? (__nv_bfloat16*)(p + ozb) : nullptr;
  }
  return off;
}

// Y[M,Cout] = conv3x3(P) as one GEMM per kernel row; P centred buffer with Cin channels
static int conv_gemms(const float* P, const Geo& g, int Cin, int Cout, const float* filt, float* Y,
                      cudaStream_t stream) {
  const int kw = g.pw ? 3 : 1;
  for (int dh = 0; dh < 3; ++dh) {
    const float* A = P + (size_t)dh * g.Wp * Cin;
    const float* Bm = filt + (size_t)(dh * 3 + (g.pw ? 0 : 1)) * Cin * Cout;
    int rc = gemm_simt(0, 0, (int)g.M, Cout, kw * Cin, 1.f, A, Cin, Bm, Cout, dh ? 1.f : 0.f, Y, Cout, nullptr, stream);
    if (rc) return rc;
  }
  return B2_OK;
}
// dfilt[dh] += A_dh^T . dYframe
static int conv_wgrad(const float* P, const Geo& g, int Cin, int Cout, const float* Dbuf, float* dfilt,
                      cudaStream_t stream) {
  const int kw = g.pw ? 3 : 1;
  const float* dY = Dbuf + (size_t)g.lead * Cout;
  for (int dh = 0; dh < 3; ++dh) {
    const float* A = P + (size_t)dh * g.Wp * Cin;
    float* Cm = dfilt + (size_t)(dh * 3 + (g.pw ? 0 : 1)) * Cin * Cout;
    int rc = gemm_simt(1, 0, kw * Cin, Cout, (int)g.M, 1.f, A, Cin, dY, Cout, 1.f, Cm, Cout, nullptr, stream);
    if (rc) return rc;
  }
  return B2_OK;
}

// filt [3][3][Cin][Cout] fp32 -> bf16 [3 kernel rows][kw*Cin][Cout] (only the centre column when the image is 1 wide)
__global__ void __launch_bounds__(256)
vgg_pack_filter_bf16_kernel(const float* __restrict__ filt, int Cin, int Cout, int kw, __nv_bfloat16* __restrict__ out) {
  const int total = 3 * kw * Cin * Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % Cout, r = i / Cout;                 // r = dh*(kw*Cin) + dw*Cin + ci
    const int dh = r / (kw * Cin), q = r % (kw * Cin), dw = q / Cin, ci = q % Cin;
    out[i] = __float2bfloat16(filt[(((size_t)dh * 3 + (kw == 3 ? dw : 1)) * Cin + ci) * Cout + co]);
  }
}

// conv3x3 on tensor cores: ONE tcgen05 GEMM with K = 3*kw*Cin, A = the bf16 activation shadow read with the
// per-kernel-row shift inside the TMA producer (gemm_bf16_tc_conv)
static int conv_gemms_tc(const __nv_bfloat16* Pb, const Geo& g, int Cin, int Cout, const float* filt,
                         __nv_bfloat16* Wb, float* Y, cudaStream_t stream) {
  const int kw = g.pw ? 3 : 1;
  vgg_pack_filter_bf16_kernel<<<ew_blocks((int64_t)3 * kw * Cin * Cout), 256, 0, stream>>>(filt, Cin, Cout, kw, Wb);
  B2_LAUNCH_CHECK();
  return gemm_bf16_tc_conv((int)g.M, Cout, kw * Cin, 3, g.Wp, Pb, g.rows, Cin, Wb, Cout, Y, Cout, stream);
}
// dfilt[dh] += A_dh^T . dYframe on tensor cores (split-K, fp32 atomics); Db = bf16 shadow of the gradient buffer
static int conv_wgrad_tc(const __nv_bfloat16* Pb, const Geo& g, int Cin, int Cout, const __nv_bfloat16* Db,
                         float* dfilt, cudaStream_t stream) {
  const int kw = g.pw ? 3 : 1;
  const __nv_bfloat16* dY = Db + (size_t)g.lead * Cout;
  for (int dh = 0; dh < 3; ++dh) {
    const __nv_bfloat16* A = Pb + (size_t)dh * g.Wp * Cin;
    float* Cm = dfilt + (size_t)(dh * 3 + (g.pw ? 0 : 1)) * Cin * Cout;
    int rc = gemm_bf16_tc(1, 1, kw * Cin, Cout, (int)g.M, 1.f, A, Cin, dY, Cout, Cm, Cout, nullptr, 1 /*atomic*/, 0, stream);
    if (rc) return rc;
  }
  return B2_OK;
}
static int to_bf16(const float* src, int64_t n, __nv_bfloat16* dst, cudaStream_t stream) {
  return cast_f32_bf16(src, 1, (int)n, (int)n, dst, (int)n, stream);
}

}  // namespace b2

using namespace b2;

static int vgg_check(const b2_vgg_desc* d) {
  B2_CHECK_ARG(d != nullptr, "vgg: null descriptor");
  B2_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0, "vgg: bad shape N=%d H=%d W=%d", d->N, d->H, d->W);
  B2_CHECK_ARG(d->keep_prob > 0.f && d->keep_prob <= 1.f, "vgg: keep_prob %f", d->keep_prob);
  B2_CHECK_ARG((int64_t)d->N * (d->H + 2) * (d->W + 2) < (int64_t)2000000000, "vgg: too many frames for one call");
  return B2_OK;
}

extern "C" size_t b2_vgg_reserve_bytes(const b2_vgg_desc* d) { return d ? vgg_reserve_layout(d, nullptr, nullptr) : 0; }
extern "C" size_t b2_vgg_workspace_bytes(const b2_vgg_desc* d) { return d ? vgg_ws_layout(d, nullptr, nullptr) : 0; }
extern "C" int b2_vgg_output_size(const b2_vgg_desc* d) {
  if (!d) return 0;
  const int h2 = (d->H + 1) / 2, w2 = (d->W + 1) / 2;
  return ((h2 + 1) / 2) * ((w2 + 1) / 2) * 128;
}

extern "C" int b2_vgg_frontend_forward(const b2_vgg_desc* d, const float* x, const b2_vgg_params* p,
                                       float* out, void* reserve, void* workspace,
                                       size_t workspace_bytes, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = vgg_check(d);
  if (rc) return rc;
  B2_CHECK_ARG(x && p && out && reserve && workspace, "b2_vgg_frontend_forward: null pointer");
  VggWs w;
  if (workspace_bytes < vgg_ws_layout(d, workspace, &w)) { set_error("b2_vgg_frontend_forward: workspace too small"); return B2_ERR_WORKSPACE; }
  VggBuf b;
  vgg_reserve_layout(d, reserve, &b);
  const float keep = d->keep_prob;
  const uint64_t seed = d->dropout_seed;
  const Geo &g0 = b.g0, &g2 = b.g2, &g4 = b.g4;
  vgg_pack_kernel<<<ew_blocks(g0.rows * 3), 256, 0, stream>>>(x, g0, 3, b.P[0]);
  B2_LAUNCH_CHECK();
  // VGG1
  if ((rc = conv_gemms(b.P[0], g0, 3, 64, p->conv_w[0], w.Ya, stream))) return rc;
  vgg_epi_kernel<<<ew_blocks(g0.rows * 64), 256, 0, stream>>>(w.Ya, p->conv_b[0], g0, 64, keep, seed + 1, b.P[1]);
  B2_LAUNCH_CHECK();
  const bool tc = d->precision == B2_PREC_BF16 && b2_device_is_sm100() == 1;
  if (tc) {
    if ((rc = to_bf16(b.P[1], g0.rows * 64, b.Pb[1], stream))) return rc;
    if ((rc = conv_gemms_tc(b.Pb[1], g0, 64, 64, p->conv_w[1], w.Wb, b.Y2, stream))) return rc;
  } else if ((rc = conv_gemms(b.P[1], g0, 64, 64, p->conv_w[1], b.Y2, stream))) return rc;
  vgg_epi_pool_kernel<<<ew_blocks(g2.rows * 64), 256, 0, stream>>>(b.Y2, p->conv_b[1], g0, g2, 64, keep, seed + 2, 0, b.P[2]);
  B2_LAUNCH_CHECK();
  // VGG2
  if (tc) {
    if ((rc = to_bf16(b.P[2], g2.rows * 64, b.Pb[2], stream))) return rc;
    if ((rc = conv_gemms_tc(b.Pb[2], g2, 64, 128, p->conv_w[2], w.Wb, w.Ya, stream))) return rc;
  } else if ((rc = conv_gemms(b.P[2], g2, 64, 128, p->conv_w[2], w.Ya, stream))) return rc;
  vgg_epi_kernel<<<ew_blocks(g2.rows * 128), 256, 0, stream>>>(w.Ya, p->conv_b[2], g2, 128, keep, seed + 3, b.P[3]);
  B2_LAUNCH_CHECK();
  if (tc) {
    if ((rc = to_bf16(b.P[3], g2.rows * 128, b.Pb[3], stream))) return rc;
    if ((rc = conv_gemms_tc(b.Pb[3], g2, 128, 128, p->conv_w[3], w.Wb, b.Y4, stream))) return rc;
  } else if ((rc = conv_gemms(b.P[3], g2, 128, 128, p->conv_w[3], b.Y4, stream))) return rc;
  const int64_t nf = (int64_t)d->N * g4.H * g4.W * 128;
  vgg_epi_pool_kernel<<<ew_blocks(nf), 256, 0, stream>>>(b.Y4, p->conv_b[3], g2, g4, 128, keep, seed + 4, 1, b.F);
  B2_LAUNCH_CHECK();
  // bridge FC 256 + ReLU + dropout
  const int Kf = g4.H * g4.W * 128;
  if (tc && Kf % 8 == 0) {
    if ((rc = to_bf16(b.F, nf, b.Fb, stream))) return rc;
    if ((rc = to_bf16(p->fc_w, (int64_t)Kf * 256, w.Wb, stream))) return rc;
    if ((rc = gemm_bf16_tc(0, 1, d->N, 256, Kf, 1.f, b.Fb, Kf, w.Wb, 256, b.fc_out, 256, p->fc_b, 0 /*store*/, 0, stream))) return rc;
  } else
  if ((rc = gemm_simt(0, 0, d->N, 256, Kf, 1.f, b.F, Kf, p->fc_w, 256, 0.f, b.fc_out, 256, p->fc_b, stream))) return rc;
  vgg_fc_epi_kernel<<<ew_blocks((int64_t)d->N * 256), 256, 0, stream>>>(b.fc_out, (int64_t)d->N * 256, keep, seed + 5);
  B2_LAUNCH_CHECK();
  B2_CUDA(cudaMemcpyAsync(out, b.fc_out, (size_t)d->N * 256 * 4, cudaMemcpyDeviceToDevice, stream));
  return B2_OK;
}

extern "C" int b2_vgg_frontend_backward(const b2_vgg_desc* d, const b2_vgg_params* p, const float* d_out,
                                        const void* reserve, const b2_vgg_grads* gr, void* workspace,
                                        size_t workspace_bytes, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = vgg_check(d);
  if (rc) return rc;
  B2_CHECK_ARG(p && d_out && reserve && gr && workspace, "b2_vgg_frontend_backward: null pointer");
  VggWs w;
  if (workspace_bytes < vgg_ws_layout(d, workspace, &w)) { set_error("b2_vgg_frontend_backward: workspace too small"); return B2_ERR_WORKSPACE; }
  VggBuf b;
  vgg_reserve_layout(d, (void*)reserve, &b);
  const float keep = d->keep_prob;
  const Geo &g0 = b.g0, &g2 = b.g2, &g4 = b.g4;
  const int N = d->N, Kf = g4.H * g4.W * 128;
  // FC
  vgg_fc_epi_bwd_kernel<<<ew_blocks((int64_t)N * 256), 256, 0, stream>>>(d_out, b.fc_out, (int64_t)N * 256, keep, w.dz);
  B2_LAUNCH_CHECK();
  const bool tc = d->precision == B2_PREC_BF16 && b2_device_is_sm100() == 1;
  float* dF = w.Ga;                                           // dense [N, Kf]
  if (tc && Kf % 8 == 0) {
    if ((rc = to_bf16(w.dz, (int64_t)N * 256, w.dzb, stream))) return rc;
    if ((rc = gemm_bf16_tc(1, 1, Kf, 256, N, 1.f, b.Fb, Kf, w.dzb, 256, gr->fc_w, 256, nullptr, 1 /*atomic*/, 0, stream))) return rc;
    if ((rc = to_bf16(p->fc_w, (int64_t)Kf * 256, w.Wb, stream))) return rc;
    if ((rc = gemm_bf16_tc(0, 0, N, Kf, 256, 1.f, w.dzb, 256, w.Wb, 256, dF, Kf, nullptr, 0 /*store*/, 0, stream))) return rc;
  } else {
    if ((rc = gemm_simt(1, 0, Kf, 256, N, 1.f, b.F, Kf, w.dz, 256, 1.f, gr->fc_w, 256, nullptr, stream))) return rc;
    if ((rc = gemm_simt(0, 1, N, Kf, 256, 1.f, w.dz, 256, p->fc_w, 256, 0.f, dF, Kf, nullptr, stream))) return rc;
  }
  if ((rc = b2_colsum(w.dz, N, 256, 256, gr->fc_b, 1, stream_))) return rc;
  // conv2_2 (pooled, dense output)
  vgg_epi_pool_bwd_kernel<<<ew_blocks(g2.rows * 128), 256, 0, stream>>>(dF, b.F, b.Y4, p->conv_b[3], g2, g4, 128, keep, 1, w.Da);
  B2_LAUNCH_CHECK();
  if ((rc = b2_colsum(w.Da + (size_t)g2.lead * 128, g2.M, 128, 128, gr->conv_b[3], 1, stream_))) return rc;
  if (tc) { if ((rc = to_bf16(w.Da, g2.rows * 128, w.Dab, stream))) return rc; }
  if (tc) rc = conv_wgrad_tc(b.Pb[3], g2, 128, 128, w.Dab, gr->conv_w[3], stream);
  else rc = conv_wgrad(b.P[3], g2, 128, 128, w.Da, gr->conv_w[3], stream);
  if (rc) return rc;
  vgg_flip_filter_kernel<<<ew_blocks(9 * 128 * 128), 256, 0, stream>>>(p->conv_w[3], 128, 128, w.Wt);
  B2_LAUNCH_CHECK();
  if (tc) rc = conv_gemms_tc(w.Dab, g2, 128, 128, w.Wt, w.Wb, w.Ga, stream);
  else rc = conv_gemms(w.Da, g2, 128, 128, w.Wt, w.Ga, stream);                   // d(P3 valid), frame of g2
  if (rc) return rc;
  // conv2_1
  vgg_epi_bwd_kernel<<<ew_blocks(g2.rows * 128), 256, 0, stream>>>(w.Ga, b.P[3], g2, 128, keep, w.Da);
  B2_LAUNCH_CHECK();
  if ((rc = b2_colsum(w.Da + (size_t)g2.lead * 128, g2.M, 128, 128, gr->conv_b[2], 1, stream_))) return rc;
  if (tc) { if ((rc = to_bf16(w.Da, g2.rows * 128, w.Dab, stream))) return rc; }
  if (tc) rc = conv_wgrad_tc(b.Pb[2], g2, 64, 128, w.Dab, gr->conv_w[2], stream);
  else rc = conv_wgrad(b.P[2], g2, 64, 128, w.Da, gr->conv_w[2], stream);
  if (rc) return rc;
  vgg_flip_filter_kernel<<<ew_blocks(9 * 64 * 128), 256, 0, stream>>>(p->conv_w[2], 64, 128, w.Wt);
  B2_LAUNCH_CHECK();
  if (tc) rc = conv_gemms_tc(w.Dab, g2, 128, 64, w.Wt, w.Wb, w.Ga, stream);
  else rc = conv_gemms(w.Da, g2, 128, 64, w.Wt, w.Ga, stream);                    // d(P2 valid), frame of g2
  if (rc) return rc;
  // conv1_2 (pooled into g2)
  vgg_epi_pool_bwd_kernel<<<ew_blocks(g0.rows * 64), 256, 0, stream>>>(w.Ga, b.P[2], b.Y2, p->conv_b[1], g0, g2, 64, keep, 0, w.Da);
  B2_LAUNCH_CHECK();
  if ((rc = b2_colsum(w.Da + (size_t)g0.lead * 64, g0.M, 64, 64, gr->conv_b[1], 1, stream_))) return rc;
  if (tc) { if ((rc = to_bf16(w.Da, g0.rows * 64, w.Dab, stream))) return rc; }
  if (tc) rc = conv_wgrad_tc(b.Pb[1], g0, 64, 64, w.Dab, gr->conv_w[1], stream);
  else rc = conv_wgrad(b.P[1], g0, 64, 64, w.Da, gr->conv_w[1], stream);
  if (rc) return rc;
  vgg_flip_filter_kernel<<<ew_blocks(9 * 64 * 64), 256, 0, stream>>>(p->conv_w[1], 64, 64, w.Wt);
  B2_LAUNCH_CHECK();
  if (tc) rc = conv_gemms_tc(w.Dab, g0, 64, 64, w.Wt, w.Wb, w.Ga, stream);
  else rc = conv_gemms(w.Da, g0, 64, 64, w.Wt, w.Ga, stream);                     // d(P1 valid), frame of g0
  if (rc) return rc;
  // conv1_1 (no data gradient: the input is the feature matrix)
  vgg_epi_bwd_kernel<<<ew_blocks(g0.rows * 64), 256, 0, stream>>>(w.Ga, b.P[1], g0, 64, keep, w.Da);
  B2_LAUNCH_CHECK();
  if ((rc = b2_colsum(w.Da + (size_t)g0.lead * 64, g0.M, 64, 64, gr->conv_b[0], 1, stream_))) return rc;
  if ((rc = conv_wgrad(b.P[0], g0, 3, 64, w.Da, gr->conv_w[0], stream))) return rc;
  return B2_OK;
}

// ReLU + tf.nn.dropout as one in-place pass / its backward; also the epilogue of the CTC model's
// bottleneck layer (models/ctc/ctc.py:200-213).
extern "C" int b2_relu_dropout_forward(float* x, int64_t n, float keep_prob, uint64_t seed, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(x && n > 0 && keep_prob > 0.f && keep_prob <= 1.f, "b2_relu_dropout_forward: bad argument");
  vgg_fc_epi_kernel<<<ew_blocks(n), 256, 0, stream>>>(x, n, keep_prob, seed);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
extern "C" int b2_relu_dropout_backward(const float* d_out, const float* out, int64_t n, float keep_prob,
                                        float* d_in, b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(d_out && out && d_in && n > 0 && keep_prob > 0.f && keep_prob <= 1.f, "b2_relu_dropout_backward: bad argument");
  vgg_fc_epi_bwd_kernel<<<ew_blocks(n), 256, 0, stream>>>(d_out, out, n, keep_prob, d_in);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
