// CTC prefix beam search for sm_100a.
//
// Replaces BeamSearchDecoder.__call__ of the reference's numpy decoder
// (models/ctc/decoders/beam_search_decoder.py:53-152), the decoder
// examples/librispeech/evaluation/eval_ctc.py:126-131 selects (beam width 20), restated in
// oracle/decode.py.  Semantics kept: every (beam entry) x (class) candidate is evaluated, no
// pruning, blank keeps the prefix, a repeated last character updates both the extended prefix
// (from p_b only) and the unchanged prefix (from p_nb), an extension that equals a prefix
// already in the beam merges with it, ranking key = logsumexp(p_b, p_nb) in float64, ties
// broken by the dict insertion order of the reference (class-major, beam-minor).
//
// One CTA per utterance.  Prefixes live in a parent-pointer tree in the workspace (exact
// equality test by walking the chains, no hashing).  Per frame: (1) every beam entry finds
// the beam entry that is its own prefix minus the last character, (2) stay candidates and
// new-prefix candidates are scored in fp64, (3) W rounds of block-wide arg-max pick the next
// beam.  Integer/ordering work, latency-bound; labels are bit-exact against the oracle.
#include "common.cuh"
#include <math_constants.h>

namespace b2 {

constexpr int kBeamThreads = 256;
constexpr int kMaxBeam = 128;

__device__ __forceinline__ double dlse2(double a, double b) {
  if (a == -CUDART_INF && b == -CUDART_INF) return -CUDART_INF;
  const double m = a > b ? a : b;
  return m + log(exp(a - m) + exp(b - m));
}

struct Cand { double s; int key; int idx; };    // idx < 0: none

__device__ __forceinline__ bool cand_better(const Cand& a, const Cand& b) {
  if (a.idx < 0) return false;
  if (b.idx < 0) return true;
  if (a.s > b.s) return true;
  if (a.s < b.s) return false;
  return a.key < b.key;
}
__device__ __forceinline__ Cand cand_shfl(const Cand& c, int off) {
  Cand r;
  r.s = __shfl_xor_sync(0xffffffffu, c.s, off);
  r.key = __shfl_xor_sync(0xffffffffu, c.key, off);
  r.idx = __shfl_xor_sync(0xffffffffu, c.idx, off);
  return r;
}

// scores buffer: [W*C] doubles (NaN = not a candidate / already taken)
__global__ void __launch_bounds__(kBeamThreads)
ctc_beam_kernel(const float* __restrict__ log_probs, const int* __restrict__ seq_len, int T, int B,
                int C, int blank, int W, int* __restrict__ node_parent_all,
                int* __restrict__ node_char_all, double* __restrict__ scores_all,
                int* __restrict__ out_labels, int* __restrict__ out_len, float* __restrict__ out_score) {
  __shared__ double pb[kMaxBeam], pnb[kMaxBeam], spb[kMaxBeam], spnb[kMaxBeam], stot[kMaxBeam];
  __shared__ double npb[kMaxBeam], npnb[kMaxBeam];
  __shared__ int node[kMaxBeam], endc[kMaxBeam], blen[kMaxBeam], pi[kMaxBeam], skey[kMaxBeam];
  __shared__ int nnode[kMaxBeam], nendc[kMaxBeam], nblen[kMaxBeam];
  __shared__ int sel_idx[kMaxBeam];
  __shared__ Cand wbest[kBeamThreads / 32];
  __shared__ int s_wc, s_nodes;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tb = min(seq_len[b], T);
  const size_t node_cap = (size_t)T * W + 2;
  int* nparent = node_parent_all + (size_t)b * node_cap;
  int* nchar = node_char_all + (size_t)b * node_cap;
  double* S = scores_all + (size_t)b * W * C;
  const double NINF = -CUDART_INF;

  if (tid == 0) {
    nparent[0] = -1; nchar[0] = -1;
    node[0] = 0; endc[0] = -1; blen[0] = 0; pb[0] = 0.0; pnb[0] = NINF;
    s_wc = 1; s_nodes = 1;
  }
  __syncthreads();

  for (int t = 0; t < Tb; ++t) {
    const float* lp = log_probs + ((size_t)b * T + t) * C;
    const int Wc = s_wc;
    // (1) pi[j] = index of the beam entry equal to prefix_j minus its last character
    if (tid < Wc) {
      const int j = tid;
      int found = -1;
      if (blen[j] > 0) {
        const int pj = nparent[node[j]];
        for (int i = 0; i < Wc && found < 0; ++i) {
          if (blen[i] != blen[j] - 1) continue;
          int a = node[i], c2 = pj;
          bool eq = true;
          while (a != c2) {
            if (a < 0 || c2 < 0 || nchar[a] != nchar[c2]) { eq = false; break; }
            a = nparent[a]; c2 = nparent[c2];
          }
          if (eq) found = i;
        }
      }
      pi[j] = found;
    }
    __syncthreads();
    // (2a) stay candidates
    if (tid < Wc) {
      const int j = tid;
      const double lpb = (double)lp[blank];
      const double q_b = dlse2(pb[j] + lpb, pnb[j] + lpb);
      double q_nb = NINF;
      int key = (blank * W + j) * 2;
      if (blen[j] > 0) {
        const int e = endc[j];
        q_nb = pnb[j] + (double)lp[e];                       // repeat of the last character
        key = min(key, (e * W + j) * 2 + 1);
        const int i = pi[j];
        if (i >= 0) {                                        // prefix_i + e == prefix_j
          const double l = (double)lp[e];
          const double contrib = (blen[i] == 0 || endc[i] != e) ? dlse2(pb[i] + l, pnb[i] + l)
                                                                : pb[i] + l;
          q_nb = dlse2(q_nb, contrib);
          key = min(key, (e * W + i) * 2);
        }
      }
      spb[j] = q_b; spnb[j] = q_nb; stot[j] = dlse2(q_b, q_nb); skey[j] = key;
    }
    // (2b) new-prefix candidates (i, c)
    const int N = Wc * C;
    for (int n = tid; n < N; n += kBeamThreads) {
      const int i = n / C, c = n % C;
      double sc = CUDART_NAN;
      if (c != blank) {
        bool merged = false;
        for (int j = 0; j < Wc; ++j)
          if (pi[j] == i && endc[j] == c) { merged = true; break; }
        if (!merged) {
          const double l = (double)lp[c];
          sc = (blen[i] == 0 || endc[i] != c) ? dlse2(pb[i] + l, pnb[i] + l) : pb[i] + l;
        }
      }
      S[n] = sc;
    }
    __syncthreads();
    // (3) W rounds of arg-max over stay entries (idx = N + j) and new entries (idx = n)
    int nsel = 0;
    for (int r = 0; r < W; ++r) {
      Cand best; best.idx = -1; best.s = 0.0; best.key = 0;
      for (int n = tid; n < N + Wc; n += kBeamThreads) {
        Cand c;
        if (n < N) {
          c.s = S[n];
          if (c.s != c.s) continue;                          // NaN: not a candidate
          const int i = n / C, cc = n % C;
          c.key = (cc * W + i) * 2; c.idx = n;
        } else {
          const int j = n - N;
          c.s = stot[j];
          if (c.s != c.s) continue;
          c.key = skey[j]; c.idx = n;
        }
        if (cand_better(c, best)) best = c;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const Cand other = cand_shfl(best, o);
        if (cand_better(other, best)) best = other;
      }
      if (lane == 0) wbest[warp] = best;
      __syncthreads();
      if (tid == 0) {
        Cand bb = wbest[0];
        for (int w = 1; w < kBeamThreads / 32; ++w)
          if (cand_better(wbest[w], bb)) bb = wbest[w];
        sel_idx[r] = bb.idx;
        if (bb.idx >= 0) {                                   // mark taken
          if (bb.idx < N) S[bb.idx] = CUDART_NAN; else stot[bb.idx - N] = CUDART_NAN;
        }
      }
      __syncthreads();
      if (sel_idx[r] < 0) break;
      nsel = r + 1;
    }
    // (4) materialise the next beam (thread 0 allocates tree nodes in selection order)
    if (tid == 0) {
      int nodes = s_nodes;
      for (int r = 0; r < nsel; ++r) {
        const int idx = sel_idx[r];
        if (idx >= N) {
          const int j = idx - N;
          nnode[r] = node[j]; nendc[r] = endc[j]; nblen[r] = blen[j];
          npb[r] = spb[j]; npnb[r] = spnb[j];
        } else {
          const int i = idx / C, c = idx % C;
          const double l = (double)lp[c];
          nparent[nodes] = node[i]; nchar[nodes] = c;
          nnode[r] = nodes++; nendc[r] = c; nblen[r] = blen[i] + 1;
          npb[r] = NINF;
          npnb[r] = (blen[i] == 0 || endc[i] != c) ? dlse2(pb[i] + l, pnb[i] + l) : pb[i] + l;
        }
      }
      s_nodes = nodes;
      s_wc = nsel;
    }
    __syncthreads();
    if (tid < nsel) {
      node[tid] = nnode[tid]; endc[tid] = nendc[tid]; blen[tid] = nblen[tid];
      pb[tid] = npb[tid]; pnb[tid] = npnb[tid];
    }
    __syncthreads();
  }
  // best hypothesis = beam[0]
  if (tid == 0) {
    const int L = blen[0];
    out_len[b] = L;
    out_score[b] = (float)(-dlse2(pb[0], pnb[0]));
    int a = node[0];
    int* out = out_labels + (size_t)b * T;
    for (int k = L - 1; k >= 0; --k) { out[k] = nchar[a]; a = nparent[a]; }
    for (int k = L; k < T; ++k) out[k] = -1;
  }
}

struct BeamWs { int* parent; int* chr; double* scores; };
static size_t beam_ws_layout(int T, int B, int C, int W, void* base, BeamWs* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t cap = (size_t)T * W + 2;
  const size_t op = take((size_t)B * cap * 4), oc = take((size_t)B * cap * 4);
  const size_t os = take((size_t)B * W * C * 8);
  if (w) {
    char* p = (char*)base;
    w->parent = (int*)(p + op); w->chr = (int*)(p + oc); w->scores = (double*)(p + os);
  }
  return off;
}

}  // namespace b2

using namespace b2;

extern "C" size_t b2_ctc_beam_workspace_bytes(int T, int B, int C, int beam_width) {
  return beam_ws_layout(T, B, C, beam_width, nullptr, nullptr);
}

extern "C" int b2_ctc_beam_decode(const float* log_probs, const int32_t* seq_len, int T, int B, int C,
                                  int blank, int beam_width, int32_t* out_labels, int32_t* out_len,
                                  float* out_score, void* workspace, size_t workspace_bytes,
                                  b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(log_probs && seq_len && out_labels && out_len && out_score && workspace,
               "b2_ctc_beam_decode: null pointer");
  B2_CHECK_ARG(T > 0 && B > 0 && C > 1 && blank >= 0 && blank < C, "b2_ctc_beam_decode: bad shape");
  B2_CHECK_ARG(beam_width >= 1 && beam_width <= kMaxBeam, "b2_ctc_beam_decode: beam width %d not in [1,%d]",
               beam_width, kMaxBeam);
  BeamWs w;
  const size_t need = beam_ws_layout(T, B, C, beam_width, workspace, &w);
  if (workspace_bytes < need) { set_error("b2_ctc_beam_decode: workspace %zu < %zu", workspace_bytes, need); return B2_ERR_WORKSPACE; }
  ctc_beam_kernel<<<B, kBeamThreads, 0, stream>>>(log_probs, seq_len, T, B, C, blank, beam_width,
                                                 w.parent, w.chr, w.scores, out_labels, out_len,
                                                 out_score);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
