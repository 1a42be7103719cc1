// Argument blocks of the tcgen05 recurrence kernels (lstm_rec_tc.cu), shared with the
// bf16 layer orchestration (lstm_tc.cu).
#pragma once
#include "lstm_internal.cuh"

namespace b2 {

constexpr int RU = 32;        // hidden units per CTA
constexpr int RN = 16;        // batch columns per chain (MMA N)
constexpr int kDefaultGateWarps = 4;   // gate-math warps per chain (4 or 8; B2_REC_GW overrides)

struct RecFwdArgs {
  int T, B, H, NG;            // NG = number of 16-wide batch groups per direction (filled by launcher)
  const int* seq_len;
  const uint16_t* wpack;      // [2][CS][128][H] bf16, row r = unit_local*4 + gate
  const float* wi[2]; const float* wf[2]; const float* wo[2];
  int use_peephole; float forget_bias, cell_clip, keep_prob; unsigned long long seed;
  float* y;                   // [T,B,2H] fp32
  __nv_bfloat16* hs_lp;       // [T*B, 2H] bf16 undropped h (or null)
  __nv_bfloat16* y_lp;        // [T*B, 2H] bf16 output after dropout (or null; may alias hs_lp)
  float* gates; float* cs;    // reserve ([T,B,2,H,4], [T,B,2,H]) or null
  float* final_state;         // [4,B,H] or null
  long long* dbg;             // optional phase timers (clock64 sums), cluster 0 / CTA 0 only
  unsigned* progress;         // optional [ceil(T / chunk_T)] counters: +1 per (CTA, chain) once its outputs of every
  int chunk_T;                //   frame of the chunk [k*chunk_T, (k+1)*chunk_T) are stored (consumer: the next layer's
                              //   gate GEMM, launched chunk by chunk while this kernel is still running)
};

struct RecBwdArgs {
  int T, B, H, NG;
  const int* seq_len;
  const uint16_t* wpackT;     // [2][CS][4][128][128] bf16: [m-tile][unit-in row][gate col r]
  const float* wi[2]; const float* wf[2]; const float* wo[2];
  int use_peephole; float cell_clip, keep_prob; unsigned long long seed;
  const float* gates; const float* cs;   // reserve
  __nv_bfloat16* dG;          // [T*B, 8H] bf16, column = dir*4H + u*4 + gate
  const float* dfinal;        // [4,B,H] or null
  float* dbias;               // [8H] packed bias gradient (+=), or null
  float* dwi[2]; float* dwf[2]; float* dwo[2];   // peephole gradients (+=)
  int wait_mode;              // how the gate warps wait for the peers' partials (see the kernel; B2_REC_WAIT)
  unsigned* resident;         // optional device counter: +1 per cluster once all of its CTAs are running
  long long* dbg;
  unsigned* progress;         // optional per-chunk counters (see RecFwdArgs): +1 per (CTA, chain) once dG of every frame
  int chunk_T;                //   of the chunk is stored -> the dX GEMM runs chunk by chunk beside this kernel
};

bool rec_tc_supported(int H);
int rec_tc_forward(RecFwdArgs a, const float* G, int nchain, int gate_warps, cudaStream_t stream);
// dy: [T,B,2H] fp32 gradient of the layer output (after dropout)
int rec_tc_backward(RecBwdArgs a, const float* dy, int nchain, int gate_warps, cudaStream_t stream);

}  // namespace b2
