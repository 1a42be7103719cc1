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
// One CTA per utterance.  TF's Step() is a SEQUENTIAL procedure whose outcome depends on its order: leaves grow in
// descending old-probability order, children in ascending label order, each accepted child evicts the current bottom
// of the beam, and a child that is REJECTED has both its old and new probabilities reset -- when that child is itself
// a leaf evicted earlier in the same frame, it thereby loses its own turn to grow (ctc_beam_search.h, "Deactivate
// child").  An exact top-W selection is therefore NOT equivalent; this kernel reproduces the procedure:
//   1. x = logits row - max (shared memory);
//   2. every leaf is updated in parallel (fp64), and the growth order (descending old total) is ranked in parallel;
//   3. per growing leaf b: all threads compute the candidate totals of its children, then ONE thread replays TF's
//      insert / evict / reset sequence over them (rejections are a compare against the cached bottom);
//   4. the surviving members become the next frame's leaves; new children get tree nodes.
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

// dynamic shared memory: [xs: C doubles][score: C doubles][child_slot: C ints]
__global__ void __launch_bounds__(kTfBeamThreads)
ctc_beam_tf_kernel(const float* __restrict__ logits, const int* __restrict__ seq_len, int T, int B, int C,
                   int blank, int W, int merge_repeated, int* __restrict__ node_parent_all,
                   int* __restrict__ node_label_all, int* __restrict__ out_labels, int* __restrict__ out_len,
                   float* __restrict__ out_score) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* xs = (double*)smem_raw;                // [C] row - max
  double* score = xs + C;                        // [C] candidate totals of the growing leaf's children
  int* child_slot = (int*)(score + C);           // [C] existing leaf that IS child (b, c), or -1
  // leaves of the previous frame ("branches")
  __shared__ double tot[kTfMaxBeam], pbl[kTfMaxBeam], plb[kTfMaxBeam];      // new (t)
  __shared__ double otot[kTfMaxBeam], opbl[kTfMaxBeam];                      // old (t-1)
  __shared__ int node[kTfMaxBeam], label[kTfMaxBeam], pslot[kTfMaxBeam], order[kTfMaxBeam];
  __shared__ unsigned char alive[kTfMaxBeam], oreset[kTfMaxBeam];
  // members of the beam while it grows: kind 0 = branch `idx` with its updated probabilities, 1 = branch `idx`
  // re-entered as a fresh child (it keeps its node), 2 = new child (parent branch `par`, label `lab`)
  __shared__ int m_kind[kTfMaxBeam], m_idx[kTfMaxBeam], m_par[kTfMaxBeam], m_lab[kTfMaxBeam];
  __shared__ double m_tot[kTfMaxBeam];
  __shared__ double nt[kTfMaxBeam], nb[kTfMaxBeam], nl[kTfMaxBeam];
  __shared__ int nnode[kTfMaxBeam], nlabel[kTfMaxBeam], npar[kTfMaxBeam];
  __shared__ double red[kTfBeamThreads / 32];
  __shared__ int s_n, s_nodes, s_members, s_bottom;
  __shared__ double s_max, s_bottom_val;

  const int b = blockIdx.x, tid = threadIdx.x;
  const int Tb = min(seq_len[b], T);
  const size_t node_cap = (size_t)T * W + 1;
  int* node_parent = node_parent_all + (size_t)b * node_cap;
  int* node_label = node_label_all + (size_t)b * node_cap;

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
    if ((tid & 31) == 0) red[tid >> 5] = (double)m;
    __syncthreads();
    if (tid == 0) {
      double mm = red[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mm = fmax(mm, red[w]);
      s_max = mm;
    }
    __syncthreads();
    const double mx = s_max;
    for (int c = tid; c < C; c += blockDim.x) xs[c] = (double)row[c] - mx;
    if (tid < n) { otot[tid] = tot[tid]; opbl[tid] = pbl[tid]; }
    __syncthreads();
    // ---- 2. existing leaves (first loop of Step) + growth order = descending old total
    if (tid < n) {
      const int i = tid;
      double l = plb[i];
      if (label[i] >= 0) {
        const int p = pslot[i];
        if (p >= 0) l = tf_lse2(l, (label[i] == label[p]) ? opbl[p] : otot[p]);
        l += xs[label[i]];
      }
      const double bk = otot[i] + xs[blank];
      plb[i] = l; pbl[i] = bk; tot[i] = tf_lse2(bk, l);
      int r = 0;
      for (int j = 0; j < n; ++j) r += (otot[j] > otot[i]) || (otot[j] == otot[i] && j < i);
      order[r] = i;
      alive[i] = 1; oreset[i] = 0;
      m_kind[i] = 0; m_idx[i] = i; m_tot[i] = tot[i];
    }
    __syncthreads();
    if (tid == 0) {
      s_members = n;
      int bi = -1; double bv = CUDART_INF;
      if (n == W) for (int k = 0; k < n; ++k) if (m_tot[k] < bv) { bv = m_tot[k]; bi = k; }
      s_bottom = bi; s_bottom_val = (n == W) ? bv : -CUDART_INF;
    }
    __syncthreads();
    // ---- 3. growth, leaf by leaf in TF's order
    for (int r = 0; r < n; ++r) {
      const int g = order[r];
      const bool full = s_members == W;
      // is_candidate(b->oldp): total nonzero and (beam not full or better than the bottom); a leaf whose
      // probabilities were reset as a rejected child earlier in this frame has lost its turn
      const bool go = !oreset[g] && otot[g] != -CUDART_INF && (!full || otot[g] > s_bottom_val);
      if (!go) continue;                                   // uniform: decided from shared state
      for (int c = tid; c < C; c += blockDim.x) {
        child_slot[c] = -1;
        score[c] = (c == blank) ? -CUDART_INF : xs[c] + ((c == label[g]) ? opbl[g] : otot[g]);
      }
      __syncthreads();
      if (tid < n && pslot[tid] == g) child_slot[label[tid]] = tid;
      __syncthreads();
      if (tid == 0) {
        int members = s_members, bi = s_bottom;
        double bv = s_bottom_val;
        for (int c = 0; c < C; ++c) {
          if (c == blank) continue;
          const int e = child_slot[c];
          const double sc = score[c];
          if (e >= 0 && alive[e]) continue;                // c.Active(): the child is in the beam already
          const bool cand = sc != -CUDART_INF && (members < W || sc > bv);
          if (!cand) {
            if (e >= 0) oreset[e] = 1;                     // "Deactivate child": oldp and newp reset
            continue;
          }
          int slot;
          if (members == W) {                              // the bottom leaves the beam
            slot = bi;
            if (m_kind[slot] != 2) alive[m_idx[slot]] = 0;
          } else {
            slot = members++;
          }
          if (e >= 0) { m_kind[slot] = 1; m_idx[slot] = e; alive[e] = 1; }
          else { m_kind[slot] = 2; m_idx[slot] = -1; }
          m_par[slot] = g; m_lab[slot] = c; m_tot[slot] = sc;
          if (members == W) {
            bi = 0; bv = m_tot[0];
            for (int k = 1; k < W; ++k) if (m_tot[k] < bv) { bv = m_tot[k]; bi = k; }
          }
        }
        s_members = members; s_bottom = bi; s_bottom_val = (members == W) ? bv : -CUDART_INF;
      }
      __syncthreads();
    }
    // ---- 4. members -> leaves of the next frame
    const int nm = s_members;
    if (tid == 0) {
      for (int k = 0; k < nm; ++k) {
        if (m_kind[k] == 0) {
          const int i = m_idx[k];
          nnode[k] = node[i]; nlabel[k] = label[i]; npar[k] = node_parent[node[i]];
          nt[k] = tot[i]; nb[k] = pbl[i]; nl[k] = plb[i];
        } else {
          int id;
          if (m_kind[k] == 1) id = node[m_idx[k]];
          else { id = s_nodes++; node_parent[id] = node[m_par[k]]; node_label[id] = m_lab[k]; }
          nnode[k] = id; nlabel[k] = m_lab[k]; npar[k] = node[m_par[k]];
          nt[k] = m_tot[k]; nb[k] = -CUDART_INF; nl[k] = m_tot[k];
        }
      }
    }
    __syncthreads();
    if (tid < nm) {
      int ps = -1;
      if (npar[tid] >= 0)
        for (int k = 0; k < nm; ++k)
          if (nnode[k] == npar[tid]) { ps = k; break; }
      m_par[tid] = ps;
    }
    __syncthreads();
    if (tid < nm) {
      node[tid] = nnode[tid]; label[tid] = nlabel[tid]; pslot[tid] = m_par[tid];
      tot[tid] = nt[tid]; pbl[tid] = nb[tid]; plb[tid] = nl[tid];
    }
    if (tid == 0) s_n = nm;
    __syncthreads();
  }

  // ---- TopPaths(1): best total; LabelSeq(merge_repeated) walks leaf -> root
  if (tid == 0) {
    const int n = s_n;
    int best = -1; double bv = -CUDART_INF;
    for (int k = 0; k < n; ++k) if (best < 0 || tot[k] > bv) { bv = tot[k]; best = k; }
    int* out = out_labels + (size_t)b * T;
    int len = 0;
    if (best >= 0) {
      int cur = node[best], prev_label = -1;
      while (cur >= 0 && node_label[cur] >= 0) {
        const int l = node_label[cur];
        if (!merge_repeated || l != prev_label) out[len++] = l;
        prev_label = l;
        cur = node_parent[cur];
      }
      for (int i = 0; i < len / 2; ++i) { const int tmp = out[i]; out[i] = out[len - 1 - i]; out[len - 1 - i] = tmp; }
    }
    for (int i = len; i < T; ++i) out[i] = -1;
    out_len[b] = len;
    out_score[b] = best >= 0 ? (float)bv : -INFINITY;
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
  const size_t smem = 2 * (size_t)C * sizeof(double) + (size_t)C * sizeof(int);
  B2_CHECK_ARG(smem <= 200 * 1024, "b2_ctc_beam_decode_tf: beam %d x vocabulary %d needs %zu bytes of shared memory",
               beam_width, C, smem);
  B2_CUDA(cudaFuncSetAttribute(ctc_beam_tf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_beam_tf_kernel<<<B, kTfBeamThreads, smem, stream>>>(logits, seq_len, T, B, C, blank, beam_width,
                                                           merge_repeated ? 1 : 0, parent, lab, out_labels, out_len,
                                                           out_score);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
