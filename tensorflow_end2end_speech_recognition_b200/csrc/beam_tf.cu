// CTC beam search with the semantics of tf.nn.ctc_beam_search_decoder, for sm_100a.
//
// Replaces the op the reference's model calls at models/ctc/ctc.py:344-346 and
// models/ctc/multitask_ctc.py:344-349 (beam_width from the config -- 100 in
// examples/librispeech/config/ctc/blstm_ctc_960h_char.yml:44 --, top_paths=1, merge_repeated=True).
// TensorFlow's algorithm (tensorflow/core/util/ctc/ctc_beam_search.h, restated in oracle/decode.py::
// tf_ctc_beam_search_single, which documents what is reproduced and that it is unpinned against TF itself):
// prefix tree, per leaf (P_total, P_blank, P_label) in the log domain; per frame the existing leaves are updated
// (the label part is fed from the parent only while the parent is still in the beam), then every (leaf, label)
// child not yet in the beam is a candidate, and the beam keeps the `beam_width` best totals of leaves + candidates;
// the emitted path optionally drops a label that repeats its successor (merge_repeated).
//
// One CTA per utterance.  The sequential insert-if-better-than-the-bottom loop of TF equals an exact top-W
// selection (a candidate's total never exceeds its parent's old total, so TF's early-outs prune nothing that could
// enter), which is what this kernel computes: per frame
//   1. x = logits row - max (shared memory), the W+1 labels with the largest x (a candidate outside them has at
//      least W better siblings), 2. leaf update in fp64, 3. candidate totals, fp64, [leaf][top label | own label],
//   4. W rounds of block-wide arg-max over leaves + candidates (ties: lowest index), 5. new leaves get tree nodes.
// Integer / ordering work, latency-bound; label sequences are bit-exact against the oracle.
#include "common.cuh"
#include <math_constants.h>

namespace b2 {

constexpr int kTfBeamThreads = 256;
constexpr int kTfMaxBeam = 128;

__device__ __forceinline__ double tf_lse2(double a, double b) {
  if (a == -CUDART_INF && b == -CUDART_INF) return -CUDART_INF;
  const double m = a > b ? a : b;
  return m + log(exp(a - m) + exp(b - m));
}

struct TfBest { double s; int idx; };     // idx < 0: none
__device__ __forceinline__ bool tf_better(const TfBest& a, const TfBest& b) {
  if (a.idx < 0) return false;
  if (b.idx < 0) return true;
  if (a.s > b.s) return true;
  if (a.s < b.s) return false;
  return a.idx < b.idx;
}
// block-wide arg-max over vals[0..n) (entries equal to -inf are not candidates); every thread gets the result
__device__ TfBest tf_block_argmax(const double* vals, int n, TfBest* wbest) {
  TfBest b; b.s = -CUDART_INF; b.idx = -1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = vals[i];
    if (v == -CUDART_INF) continue;
    TfBest c; c.s = v; c.idx = i;
    if (tf_better(c, b)) b = c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    TfBest c;
    c.s = __shfl_xor_sync(0xffffffffu, b.s, o);
    c.idx = __shfl_xor_sync(0xffffffffu, b.idx, o);
    if (tf_better(c, b)) b = c;
  }
  if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = b;
  __syncthreads();
  b = wbest[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
    if (tf_better(wbest[w], b)) b = wbest[w];
  __syncthreads();
  return b;
}

// dynamic shared memory: [cand: W*(W+2) doubles][xs: C doubles][xsel: C doubles][rank: C ints]
__global__ void __launch_bounds__(kTfBeamThreads)
ctc_beam_tf_kernel(const float* __restrict__ logits, const int* __restrict__ seq_len, int T, int B, int C,
                   int blank, int W, int merge_repeated, int* __restrict__ node_parent_all,
                   int* __restrict__ node_label_all, int* __restrict__ out_labels, int* __restrict__ out_len,
                   float* __restrict__ out_score) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int WL = min(W + 1, C - 1);              // labels considered per leaf (+ the leaf's own label)
  const int CS = WL + 1;                         // candidate slots per leaf
  double* cand = (double*)smem_raw;              // [W][CS]
  double* xs = cand + (size_t)W * CS;            // [C] row - max
  double* xsel = xs + C;                         // [C] scratch for the label selection
  int* rank = (int*)(xsel + C);                  // [C] position in topl, or -1
  __shared__ double tot[kTfMaxBeam], pbl[kTfMaxBeam], plb[kTfMaxBeam];      // new (t)
  __shared__ double otot[kTfMaxBeam], opbl[kTfMaxBeam];                      // old (t-1)
  __shared__ double ntot[kTfMaxBeam], npbl[kTfMaxBeam], nplb[kTfMaxBeam];
  __shared__ int node[kTfMaxBeam], label[kTfMaxBeam], pslot[kTfMaxBeam];
  __shared__ int nnode[kTfMaxBeam], nlabel[kTfMaxBeam], npar[kTfMaxBeam];
  __shared__ int topl[kTfMaxBeam + 1];
  __shared__ unsigned int active_bits[kTfMaxBeam * 5];                        // [W][CS <= 130 bits -> 5 words]
  __shared__ TfBest wbest[kTfBeamThreads / 32];
  __shared__ double lvals[kTfMaxBeam];
  __shared__ int s_n, s_nodes;
  __shared__ double s_max;

  const int b = blockIdx.x, tid = threadIdx.x;
  const int Tb = min(seq_len[b], T);
  const size_t node_cap = (size_t)T * W + 1;
  int* node_parent = node_parent_all + (size_t)b * node_cap;
  int* node_label = node_label_all + (size_t)b * node_cap;

  for (int c = tid; c < C; c += blockDim.x) rank[c] = -1;
  if (tid == 0) {
    node[0] = 0; label[0] = -1; pslot[0] = -1;
    tot[0] = 0.0; pbl[0] = 0.0; plb[0] = -CUDART_INF;
    node_parent[0] = -1; node_label[0] = -1;
    s_n = 1; s_nodes = 1;
  }
  __syncthreads();

  for (int t = 0; t < Tb; ++t) {
    const int n = s_n;
    // ---- 1. x = row - max
    const float* row = logits + ((size_t)t * B + b) * C;
    float m = -INFINITY;
    for (int c = tid; c < C; c += blockDim.x) m = fmaxf(m, row[c]);
    m = warp_max(m);
    if ((tid & 31) == 0) lvals[tid >> 5] = (double)m;
    __syncthreads();
    if (tid == 0) {
      double mm = lvals[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mm = fmax(mm, lvals[w]);
      s_max = mm;
    }
    __syncthreads();
    const double mx = s_max;
    for (int c = tid; c < C; c += blockDim.x) {
      const double v = (double)row[c] - mx;
      xs[c] = v;
      xsel[c] = (c == blank) ? -CUDART_INF : v;
    }
    __syncthreads();
    // the WL labels with the largest x (all of them when the vocabulary is small)
    if (C - 1 <= WL) {
      for (int c = tid; c < C; c += blockDim.x)
        if (c != blank) { const int j = c < blank ? c : c - 1; topl[j] = c; rank[c] = j; }
      __syncthreads();
    } else {
      for (int j = 0; j < WL; ++j) {
        const TfBest bst = tf_block_argmax(xsel, C, wbest);
        if (tid == 0) { topl[j] = bst.idx; rank[bst.idx] = j; xsel[bst.idx] = -CUDART_INF; }
        __syncthreads();
      }
    }
    // ---- 2. existing leaves: old <- new, then the update of CTCBeamSearchDecoder::Step's first loop
    if (tid < n) { otot[tid] = tot[tid]; opbl[tid] = pbl[tid]; }
    for (int i = tid; i < W * 5; i += blockDim.x) active_bits[i] = 0u;
    __syncthreads();
    if (tid < n) {
      const int i = tid;
      double nl = plb[i];
      if (label[i] >= 0) {
        const int p = pslot[i];
        if (p >= 0) {
          const double prev = (label[i] == label[p]) ? opbl[p] : otot[p];
          nl = tf_lse2(nl, prev);
          // this leaf IS the child (p, label[i]): it is not a candidate again
          const int j = (label[i] == label[p]) ? WL : rank[label[i]];
          if (j >= 0) atomicOr(&active_bits[p * 5 + (j >> 5)], 1u << (j & 31));
        }
        nl += xs[label[i]];
      }
      const double nb = otot[i] + xs[blank];
      plb[i] = nl; pbl[i] = nb; tot[i] = tf_lse2(nb, nl);
    }
    __syncthreads();
    // ---- 3. candidates [leaf i][slot j]: j < WL -> label topl[j] (skipped when it is the leaf's own label),
    //         j == WL -> the leaf's own label, fed from P_blank only
    for (int k = tid; k < n * CS; k += blockDim.x) {
      const int i = k / CS, j = k - i * CS;
      double v = -CUDART_INF;
      if (otot[i] != -CUDART_INF && !((active_bits[i * 5 + (j >> 5)] >> (j & 31)) & 1u)) {
        if (j < WL) {
          const int c = topl[j];
          if (c != label[i]) v = xs[c] + otot[i];
        } else if (label[i] >= 0) {
          v = xs[label[i]] + opbl[i];
        }
      }
      cand[k] = v;
    }
    if (tid < n) lvals[tid] = tot[tid];
    __syncthreads();
    // ---- 4. W rounds: best of (leaves, candidates); a leaf wins ties (it was pushed first)
    int nsel = 0;
    for (int r = 0; r < W; ++r) {
      const TfBest bl = tf_block_argmax(lvals, n, wbest);
      const TfBest bc = tf_block_argmax(cand, n * CS, wbest);
      if (bl.idx < 0 && bc.idx < 0) break;
      const bool take_leaf = bl.idx >= 0 && (bc.idx < 0 || bl.s >= bc.s);
      if (tid == 0) {
        if (take_leaf) {
          const int i = bl.idx;
          nnode[nsel] = node[i]; nlabel[nsel] = label[i]; npar[nsel] = node_parent[node[i]];
          ntot[nsel] = tot[i]; npbl[nsel] = pbl[i]; nplb[nsel] = plb[i];
          lvals[i] = -CUDART_INF;
        } else {
          const int i = bc.idx / CS, j = bc.idx - i * CS;
          const int c = j < WL ? topl[j] : label[i];
          const int id = s_nodes++;
          node_parent[id] = node[i]; node_label[id] = c;
          nnode[nsel] = id; nlabel[nsel] = c; npar[nsel] = node[i];
          ntot[nsel] = bc.s; npbl[nsel] = -CUDART_INF; nplb[nsel] = bc.s;
          cand[bc.idx] = -CUDART_INF;
        }
      }
      ++nsel;
      __syncthreads();
    }
    // ---- 5. the new beam; parent slot = position of the parent node in it (or -1: parent left the beam)
    if (tid < nsel) {
      node[tid] = nnode[tid]; label[tid] = nlabel[tid];
      tot[tid] = ntot[tid]; pbl[tid] = npbl[tid]; plb[tid] = nplb[tid];
      int ps = -1;
      for (int k = 0; k < nsel; ++k)
        if (nnode[k] == npar[tid]) { ps = k; break; }
      pslot[tid] = npar[tid] >= 0 ? ps : -1;
    }
    for (int j = tid; j < WL; j += blockDim.x) rank[topl[j]] = -1;
    if (tid == 0) s_n = nsel;
    __syncthreads();
  }

  // ---- TopPaths(1): best total; LabelSeq(merge_repeated) walks leaf -> root
  {
    const int n = s_n;
    if (tid < n) lvals[tid] = tot[tid];
    __syncthreads();
    const TfBest best = tf_block_argmax(lvals, n, wbest);
    if (tid == 0) {
      int* out = out_labels + (size_t)b * T;
      int len = 0;
      if (best.idx >= 0) {
        int cur = node[best.idx], prev_label = -1;
        while (node_parent[cur] >= 0 || node_label[cur] >= 0) {
          const int l = node_label[cur];
          if (l < 0) break;
          if (!merge_repeated || l != prev_label) out[len++] = l;
          prev_label = l;
          cur = node_parent[cur];
          if (cur < 0) break;
        }
        for (int i = 0; i < len / 2; ++i) { const int tmp = out[i]; out[i] = out[len - 1 - i]; out[len - 1 - i] = tmp; }
      }
      for (int i = len; i < T; ++i) out[i] = -1;
      out_len[b] = len;
      out_score[b] = best.idx >= 0 ? (float)best.s : -INFINITY;
    }
  }
}

static size_t tf_beam_ws_layout(int T, int B, int W, void* base, int** parent, int** lab) {
  const size_t cap = ((size_t)T * W + 1) * B;
  const size_t o0 = 0, o1 = align_up(cap * 4, 256);
  if (base) { *parent = (int*)((char*)base + o0); *lab = (int*)((char*)base + o1); }
  return o1 + align_up(cap * 4, 256);
}

}  // namespace b2

using namespace b2;

extern "C" size_t b2_ctc_beam_tf_workspace_bytes(int T, int B, int C, int beam_width) {
  (void)C;
  return tf_beam_ws_layout(T, B, beam_width, nullptr, nullptr, nullptr);
}

extern "C" int b2_ctc_beam_decode_tf(const float* logits, const int32_t* seq_len, int T, int B, int C, int blank,
                                     int beam_width, int merge_repeated, int32_t* out_labels, int32_t* out_len,
                                     float* out_score, void* workspace, size_t workspace_bytes,
                                     b2_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2_CHECK_ARG(logits && seq_len && out_labels && out_len && out_score && workspace,
               "b2_ctc_beam_decode_tf: null pointer");
  B2_CHECK_ARG(T > 0 && B > 0 && C > 1, "b2_ctc_beam_decode_tf: bad shape");
  B2_CHECK_ARG(blank == C - 1, "b2_ctc_beam_decode_tf: the blank must be the last class (TF convention), got %d of %d",
               blank, C);
  B2_CHECK_ARG(beam_width >= 1 && beam_width <= kTfMaxBeam, "b2_ctc_beam_decode_tf: beam width %d not in [1,%d]",
               beam_width, kTfMaxBeam);
  int* parent = nullptr; int* lab = nullptr;
  const size_t need = tf_beam_ws_layout(T, B, beam_width, workspace, &parent, &lab);
  if (workspace_bytes < need) { set_error("b2_ctc_beam_decode_tf: workspace %zu < %zu", workspace_bytes, need); return B2_ERR_WORKSPACE; }
  const int WL = (beam_width + 1 < C - 1) ? beam_width + 1 : C - 1;
  const size_t smem = ((size_t)beam_width * (WL + 1) + 2 * (size_t)C) * sizeof(double) + (size_t)C * sizeof(int);
  B2_CHECK_ARG(smem <= 200 * 1024, "b2_ctc_beam_decode_tf: beam %d x vocabulary %d needs %zu bytes of shared memory",
               beam_width, C, smem);
  B2_CUDA(cudaFuncSetAttribute(ctc_beam_tf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_beam_tf_kernel<<<B, kTfBeamThreads, smem, stream>>>(logits, seq_len, T, B, C, blank, beam_width,
                                                           merge_repeated ? 1 : 0, parent, lab, out_labels, out_len,
                                                           out_score);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
